#!/bin/bash
# Full GPU suite with the failure summary (the final-evidence run reported 8 failures but its files were not copied back).
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -rf --tb=short 2>&1 | grep -v "^$" > gpurun_out/r2_pytest20.log
grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest20.log | cut -c1-300
