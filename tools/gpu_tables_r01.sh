# kernel / data-class tables only (subset of gpu_measure_r01.sh)
mkdir -p gpurun_out
python tools/prof_kernels.py all 262144 8192 > gpurun_out/kernels_r01.txt 2>&1
python tools/prof_kernels.py all 262144 4096 >> gpurun_out/kernels_r01.txt 2>&1
python tools/prof_kernels.py lz4,lzf,fused 16384 65536 >> gpurun_out/kernels_r01.txt 2>&1; cat gpurun_out/kernels_r01.txt
CW_MODE=0 python tools/prof_classes.py lzf 8192 > gpurun_out/classes_lzf_r01.txt 2>&1
CW_MODE=0 python tools/prof_classes.py lzf 4096 >> gpurun_out/classes_lzf_r01.txt 2>&1
