for hw in 1 2; do for mode in 0 1; do
  export CW_FUSED_HASH_WARPS=$hw CW_MODE=$mode
  python tools/prof_classes.py fused 8192 2>&1 | grep "synthetic" | sed "s/^/hw=$hw mode=$mode /" | tee -a gpurun_out/fused_hw_classes_r01.txt
done; done
