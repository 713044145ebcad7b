#!/bin/bash
# round 2 evidence: default bench line, ncu launch list of the same command, ncu --set full captures of the fused kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_ours.json 2> gpurun_out/r02_bench_ours.err || exit 1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench_ours.err
timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/plain_launches.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
# fused fast / compat on the synthetic mix, 8 KiB; fused fast / compat at 64 KiB (single read: dram__bytes_read vs input)
for spec in "1 fused mixed 32768 8192 fusedfast_mixed_8k" "0 fused mixed 32768 8192 fusedcompat_mixed_8k" "1 fused mixed 4096 65536 fusedfast_mixed_64k" "0 fused mixed 4096 65536 fusedcompat_mixed_64k" "1 lz4 file:bible.txt 16384 8192 lz4fast_bible_8k" "1 lzf file:bible.txt 16384 8192 lzffast_bible_8k" "1 fusedlzf mixed 32768 8192 fusedlzffast_mixed_8k"; do
  set -- $spec
  CW_MODE=$1 timeout 120 python tools/prof_one.py $2 $3 $4 $5 > gpurun_out/plain_$6.log 2>&1 &&
  CW_MODE=$1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cw_ -s 1 -c 1 -f -o gpurun_out/prof_r02_$6 python tools/prof_one.py $2 $3 $4 $5 > gpurun_out/ncu_$6.log 2>&1
  tail -1 gpurun_out/ncu_$6.log
  # the reports are 25 MB each and gpurun_out/ carries 64 MiB back: keep the text summary, drop the report
  python tools/ncu_summary.py gpurun_out/prof_r02_$6.ncu-rep > gpurun_out/r02_ncu_$6.txt 2>&1
  rm -f gpurun_out/prof_r02_$6.ncu-rep
done
ls -la gpurun_out | tail -20
