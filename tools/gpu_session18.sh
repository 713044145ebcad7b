for hw in 1 2 3; do for em in 2 3; do
  export CW_FUSED_HASH_WARPS=$hw CW_LZ4_EMITTERS=$em
  python tools/prof_kernels.py fused 262144 8192 3 2>&1 | sed "s/^/hw=$hw em=$em compat /"
done; done
for hw in 1 2; do
  export CW_FUSED_HASH_WARPS=$hw CW_LZ4_EMITTERS=3
  ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:cw_fused_kernel -s 1 -c 1 --csv --log-file gpurun_out/fd_hw$hw.csv python tools/prof_kernels.py fused 262144 8192 1 > /dev/null 2>&1
  tail -2 gpurun_out/fd_hw$hw.csv | cut -d, -f13- | sed "s/^/hw=$hw /"
done
