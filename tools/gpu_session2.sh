set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=line 2>&1 | grep -v "operation not supported" | tail -12
python tools/prof_kernels.py all 65536 8192 > gpurun_out/kernels_r01.txt 2>&1; cat gpurun_out/kernels_r01.txt
python tools/prof_kernels.py all 65536 4096 >> gpurun_out/kernels_r01.txt 2>&1; tail -6 gpurun_out/kernels_r01.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -5 gpurun_out/bench2.err; cat gpurun_out/bench2.json
python tools/prof_kernels.py lz4,skein512 16384 8192 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'cw_lz4_kernel|cw_hash_kernel' -s 2 -c 2 -f -o gpurun_out/prof_r01a python tools/prof_kernels.py lz4,skein512 16384 8192 1 > gpurun_out/ncu_a.log 2>&1; tail -3 gpurun_out/ncu_a.log
