for mb in 16 32 64 128; do for sl in 3 4; do for mode in 0 1; do
  CW_SLICE_MB=$mb CW_SLOTS=$sl CW_MODE=$mode python tools/prof_e2e.py 2>&1 | grep "e2e\|alone"
done; done; done
