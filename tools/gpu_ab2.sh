timeout 600 python -m pytest tests -m gpu -q -x -k "hash or skein or fused" 2>&1 | tail -3
for v in "" _mix0 _mix1; do
  export CW_LIB=$PWD/compute_war_b200/libcomputewar$v.so
  python tools/prof_kernels.py skein512,skein256,fused 262144 8192 2>&1 | sed "s/^/v=[$v] /"
  CW_MODE=1 python tools/prof_classes.py fused 8192 2>&1 | grep "synthetic mixed\|random" | sed "s/^/v=[$v] fast /"
done
