for em in 1 2 3; do for hw in 1 2; do for mode in 0 1; do
  export CW_LZ4_EMITTERS=$em CW_FUSED_HASH_WARPS=$hw CW_MODE=$mode
  python tools/prof_classes.py fused 8192 2>&1 | grep "synthetic mixed\|bible\|(random)" | sed "s/^/em=$em hw=$hw mode=$mode /"
done; done; done
for em in 1 2 3 4; do for mode in 0 1; do
  export CW_LZ4_EMITTERS=$em CW_MODE=$mode
  python tools/prof_classes.py lz4 8192 2>&1 | grep "synthetic mixed\|bible" | sed "s/^/LZ4ONLY em=$em mode=$mode /"
done; done
