timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for mode in 0 1; do
  CW_MODE=$mode python tools/prof_classes.py lz4 65536 2>&1 | grep "synthetic mixed\|bible\|tiled" | sed "s/^/mode=$mode /"
  CW_MODE=$mode python tools/prof_classes.py lz4 8192 2>&1 | grep "synthetic mixed\|bible" | sed "s/^/mode=$mode /"
  CW_MODE=$mode python tools/prof_classes.py fused 8192 2>&1 | grep "synthetic mixed\|bible" | sed "s/^/mode=$mode /"
done
