set -x
for hw in 1 2 3 4 8; do
  export CW_FUSED_HASH_WARPS=$hw
  python tools/prof_kernels.py fused 262144 8192 3 2>&1 | sed "s/^/hash_warps=$hw /" | tee -a gpurun_out/fused_hashwarps_r01.txt
done
for hw in 1 2 4; do
  export CW_FUSED_HASH_WARPS=$hw
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cw_fused_kernel -s 1 -c 1 --csv --log-file gpurun_out/fused_dram_hw$hw.csv python tools/prof_kernels.py fused 262144 8192 1 > /dev/null 2>&1
  tail -3 gpurun_out/fused_dram_hw$hw.csv | cut -d, -f13-
done
