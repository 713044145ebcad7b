# one ncu --set full capture of the stand-alone LZ4 kernel on the synthetic mix (after the same command ran clean)
set -x
mkdir -p gpurun_out
python tools/prof_one.py lz4 mixed 32768 > gpurun_out/plainL.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cw_lz4_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_lz4_mixed python tools/prof_one.py lz4 mixed 32768 > gpurun_out/ncu_L.log 2>&1; tail -2 gpurun_out/ncu_L.log
