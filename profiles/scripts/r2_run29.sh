#!/bin/bash
# GPU run 29: full GPU test suite + smoke on the final build of the round (after the last attention / LayerNorm variants)
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -rP -p no:cacheprovider --durations=8 2>&1 | grep -v "^$" > gpurun_out/r2_pytest_gpu_full.log
tail -4 gpurun_out/r2_pytest_gpu_full.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
