#!/bin/bash
# GPU run 16: TMA stores in the specialised epilogues -- tests, K sweep, op timings A/B, bench
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "specialised" 2>&1 | tail -15 > gpurun_out/r2_pytest16a.log
cat gpurun_out/r2_pytest16a.log
timeout 300 python profiles/k_sweep.py 2>&1 | grep -E "K=   64|K=  320|K= 1280" > gpurun_out/r2_k_sweep16.txt; cat gpurun_out/r2_k_sweep16.txt
timeout 300 python profiles/run_ops.py --time > gpurun_out/r2_ops_time16.txt 2>&1; grep -E "gemm|conv" gpurun_out/r2_ops_time16.txt
timeout 300 python profiles/run_ops.py --time --lane-stores > gpurun_out/r2_ops_time16_lane.txt 2>&1; grep -E "gemm|conv" gpurun_out/r2_ops_time16_lane.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_config2_gpu.py -m gpu -q --timeout 400 -p no:cacheprovider -k "conv or gemm or unet or config2 or mmhaa" 2>&1 | tail -12 > gpurun_out/r2_pytest16.log
cat gpurun_out/r2_pytest16.log | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ops-out gpurun_out/r2_ops_step16.txt > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err
tail -5 gpurun_out/r2_bench16.err | cut -c1-150; cat gpurun_out/r2_bench16.json | cut -c1-400
timeout 400 python bench.py --quick --steps 3 --warmup 3 --tma-store 0 2>/dev/null | tail -1 | cut -c1-300
