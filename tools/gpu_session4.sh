set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=line 2>&1 | grep -v "operation not supported" | tail -8
python tools/prof_kernels.py lz4,fused 65536 8192 > gpurun_out/kernels_r01b.txt 2>&1; cat gpurun_out/kernels_r01b.txt
python tools/prof_kernels.py lz4 65536 4096 2>&1 | tee -a gpurun_out/kernels_r01b.txt
python tools/prof_kernels.py lz4 4096 65536 2>&1 | tee -a gpurun_out/kernels_r01b.txt
