python tools/prof_one.py lz4 file:bible.txt 16384 > gpurun_out/plainT.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cw_lz4_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_lz4v4_bible python tools/prof_one.py lz4 file:bible.txt 16384 > gpurun_out/ncu_T.log 2>&1; tail -2 gpurun_out/ncu_T.log
