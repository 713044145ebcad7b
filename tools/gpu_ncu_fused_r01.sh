# one ncu --set full capture of the dominant kernel (after the same command ran clean without ncu)
set -x
mkdir -p gpurun_out
python tools/prof_kernels.py fused 262144 8192 1 > gpurun_out/plainF.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cw_fused_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_fused python tools/prof_kernels.py fused 262144 8192 1 > gpurun_out/ncu_F.log 2>&1; tail -2 gpurun_out/ncu_F.log
