set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -25
python tools/prof_kernels.py lz4,lzf,fused 65536 8192 2>&1 | tee gpurun_out/kernels_r01c.txt
python tools/prof_classes.py fused 8192 2>&1 | tee gpurun_out/classes_r01_fused.txt
python tools/prof_classes.py lzf 8192 2>&1 | tee gpurun_out/classes_r01_lzf.txt
