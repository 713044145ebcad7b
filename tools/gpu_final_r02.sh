#!/bin/bash
# end-of-round evidence: whole GPU suite, smoke(), then the bench / launch list / ncu captures of tools/gpu_profiles_r02.sh
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_final_tests.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' >> gpurun_out/r02_final_tests.log 2>&1
cat gpurun_out/r02_final_tests.log
bash tools/gpu_profiles_r02.sh
