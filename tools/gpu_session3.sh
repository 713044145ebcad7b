set -x
mkdir -p gpurun_out
python tools/prof_kernels.py lz4 16384 8192 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cw_lz4_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_lz4 python tools/prof_kernels.py lz4 16384 8192 1 > gpurun_out/ncu_lz4.log 2>&1; tail -3 gpurun_out/ncu_lz4.log
python tools/prof_kernels.py skein512 262144 8192 1 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cw_hash_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_skein512 python tools/prof_kernels.py skein512 262144 8192 1 > gpurun_out/ncu_sk.log 2>&1; tail -3 gpurun_out/ncu_sk.log
