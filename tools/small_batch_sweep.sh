#!/bin/bash
# small-batch front-end: the reference's calling pattern (14 worker threads, one blocking call per work item) at
# several work-item sizes, against one whole-file call; 1 GiB of the synthetic stream, 4 KiB chunks, pageable buffers
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0,'tests'); import oracle_lib as O
O.synth(4096, 0, 262144).tofile('/tmp/sweep.bin')
PY
T=./compute_war_b200/hashandcompress
{
echo "== one call per file (pinned / pageable)"
$T -H skein512 -C lz4 -k 4096 -i 3 /tmp/sweep.bin 2>&1 | tail -2
$T -H skein512 -C lz4 -k 4096 -i 3 --no-pin /tmp/sweep.bin 2>&1 | tail -2
for r in 8 64 256 1024 2048 4096; do
  echo "== work items of $r chunks ($((r*4)) KiB), 14 threads, pageable"
  $T -H skein512 -C lz4 -k 4096 -i 3 --no-pin --work-items -c 14 -r $r /tmp/sweep.bin 2>&1 | tail -3
done
for r in 8 64; do
  echo "== work items of $r chunks ($((r*4)) KiB), 14 threads, HashOffload style (64 tickets in flight per thread), pageable"
  $T -H skein512 -C lz4 -k 4096 -i 3 --no-pin --work-items --gpu-offload 1 -c 14 -r $r /tmp/sweep.bin 2>&1 | tail -3
done
echo "== work items of 4096 chunks (16 MiB), 1 thread, pageable"
$T -H skein512 -C lz4 -k 4096 -i 3 --no-pin --work-items -c 1 -r 4096 /tmp/sweep.bin 2>&1 | tail -3
echo "== work items of 8 chunks, batcher off (every call its own launch)"
CW_NO_BATCHER=1 timeout 300 $T -H skein512 -C lz4 -k 4096 -i 1 --no-pin --work-items -c 14 -r 8 /tmp/sweep.bin 2>&1 | tail -3
echo "== four cw_submit batches vs the blocking call (pinned)"
$T -H skein512 -C lz4 -k 4096 -i 3 --gpu-offload 1 -c 4 /tmp/sweep.bin 2>&1 | tail -2
} > gpurun_out/small_batch_sweep.log 2>&1
cat gpurun_out/small_batch_sweep.log
