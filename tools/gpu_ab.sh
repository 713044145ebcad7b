for v in "" _static; do for mode in 0 1; do
  export CW_LIB=$PWD/compute_war_b200/libcomputewar$v.so CW_MODE=$mode
  python tools/prof_classes.py fused 8192 2>&1 | grep "synthetic\|bible\|kennedy" | sed "s/^/v=[$v] mode=$mode /"
done; done
