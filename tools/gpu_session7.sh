set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15
CW_MODE=0 python tools/prof_classes.py lz4 8192 2>&1 | tee gpurun_out/classes_r01_v3_compat.txt
CW_MODE=1 python tools/prof_classes.py lz4 8192 2>&1 | tee gpurun_out/classes_r01_v3_fast.txt
CW_MODE=1 python tools/prof_classes.py fused 8192 2>&1 | tee gpurun_out/classes_r01_v3_fused_fast.txt
