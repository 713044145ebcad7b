# one ncu --set full capture of the stand-alone Skein-512 kernel
set -x
mkdir -p gpurun_out
python tools/prof_kernels.py skein512,sha256 262144 8192 1 > gpurun_out/plainH.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cw_hash_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_skein512 python tools/prof_kernels.py skein512 262144 8192 1 > gpurun_out/ncu_H.log 2>&1; tail -2 gpurun_out/ncu_H.log
