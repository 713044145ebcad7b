set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
nproc; free -g | head -2
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu && /tmp/ubench > gpurun_out/ubench.json; cat gpurun_out/ubench.json
timeout 900 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -40
timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -5 gpurun_out/bench1.err; cat gpurun_out/bench1.json
