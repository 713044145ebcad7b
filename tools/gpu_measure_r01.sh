# Round-1 measurement session: bench (both arms), per-kernel and per-class tables, ncu launch list + full captures.
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref_r01.json
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -3 gpurun_out/bench_r01.err; cat gpurun_out/bench_r01.json
python tools/prof_kernels.py all 262144 8192 > gpurun_out/kernels_r01.txt 2>&1; cat gpurun_out/kernels_r01.txt
python tools/prof_kernels.py all 262144 4096 >> gpurun_out/kernels_r01.txt 2>&1
python tools/prof_kernels.py lz4,lzf,fused 16384 65536 >> gpurun_out/kernels_r01.txt 2>&1; tail -9 gpurun_out/kernels_r01.txt
for mode in 0 1; do CW_MODE=$mode python tools/prof_classes.py fused 8192 2>&1 | sed "s/^/mode=$mode /"; done > gpurun_out/classes_fused_r01.txt
CW_MODE=0 python tools/prof_classes.py lzf 8192 > gpurun_out/classes_lzf_r01.txt 2>&1
CW_MODE=0 python tools/prof_classes.py lz4 8192 > gpurun_out/classes_lz4_r01.txt 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log
