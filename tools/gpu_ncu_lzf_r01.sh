# one ncu --set full capture of the stand-alone LZF kernel on text (after the same command ran clean)
set -x
mkdir -p gpurun_out
python tools/prof_one.py lzf file:bible.txt 16384 > gpurun_out/plainZ.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cw_lzf_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_lzf_bible python tools/prof_one.py lzf file:bible.txt 16384 > gpurun_out/ncu_Z.log 2>&1; tail -2 gpurun_out/ncu_Z.log
