#!/bin/bash
# GPU run 17: TMA stores as a compile-time epilogue variant -- tests, K sweep A/B, op timings A/B, bench A/B
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "specialised" 2>&1 | tail -5 > gpurun_out/r2_pytest17a.log
cat gpurun_out/r2_pytest17a.log
timeout 300 python profiles/k_sweep.py 2>&1 | grep -E "K=   64|K=  320|K= 1280" > gpurun_out/r2_k_sweep17.txt; cat gpurun_out/r2_k_sweep17.txt
timeout 300 python profiles/run_ops.py --time > gpurun_out/r2_ops_time17.txt 2>&1; grep -E "gemm|conv" gpurun_out/r2_ops_time17.txt
timeout 300 python profiles/run_ops.py --time --lane-stores > gpurun_out/r2_ops_time17_lane.txt 2>&1; grep -E "gemm|conv" gpurun_out/r2_ops_time17_lane.txt
timeout 400 python bench.py --quick --steps 4 --warmup 3 2>/dev/null | tail -1 | cut -c1-100
timeout 400 python bench.py --quick --steps 4 --warmup 3 --tma-store 0 2>/dev/null | tail -1 | cut -c1-100
timeout 400 python bench.py --quick --steps 4 --warmup 3 2>/dev/null | tail -1 | cut -c1-100
timeout 400 python bench.py --quick --steps 4 --warmup 3 --tma-store 0 2>/dev/null | tail -1 | cut -c1-100
