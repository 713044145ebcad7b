#!/bin/bash
# host tool on a 1 GiB synthetic file: pinned vs pageable buffers, 1 GPU (and 2 if present)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys; sys.path.insert(0, "tests")
import oracle_lib as O
O.synth(8192, 0, 131072).tofile("gpurun_out/synth_1g.bin")
PY
T=compute_war_b200/hashandcompress
for args in "" "--no-pin" "--gpu-offload 1 --c-threads 4" "-m fast"; do
  echo "== $args"; $T -H skein512 -C lz4 -k 8192 -i 3 $args gpurun_out/synth_1g.bin 2>&1 | tail -2
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then echo "== --devices 2"; $T -H skein512 -C lz4 -k 8192 -i 3 --devices 2 gpurun_out/synth_1g.bin 2>&1 | tail -2; fi
rm -f gpurun_out/synth_1g.bin
