#!/bin/bash
# GPU run 6: full GPU suite after the epilogue TMEM prefetch / GN grid / temporal offset tables, bench, launch list
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2_pytest6.log
cat gpurun_out/r2_pytest6.log | tail -10
timeout 300 python profiles/run_ops.py --time > gpurun_out/r2_ops_time6.txt 2>&1; cat gpurun_out/r2_ops_time6.txt
timeout 600 python bench.py --steps 5 --warmup 3 --ops-out gpurun_out/r2_ops_step6.txt > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
tail -40 gpurun_out/r2_bench6.err | cut -c1-150; cat gpurun_out/r2_bench6.json
