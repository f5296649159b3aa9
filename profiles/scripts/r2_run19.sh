#!/bin/bash
# GPU run 19: residual through the tensor cores + no BN=128 resident plan -- tests, sweeps, op timings A/B, bench A/B
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "specialised or gemm_full or weight_stationary" 2>&1 | tail -8 > gpurun_out/r2_pytest19a.log
cat gpurun_out/r2_pytest19a.log
timeout 300 python profiles/bres_sweep.py > gpurun_out/r2_bres_sweep19.txt 2>&1; cat gpurun_out/r2_bres_sweep19.txt
timeout 300 python profiles/k_sweep.py 2>&1 | grep -E "^res" > gpurun_out/r2_k_sweep19.txt; cat gpurun_out/r2_k_sweep19.txt
timeout 300 python profiles/run_ops.py --time > gpurun_out/r2_ops_time19.txt 2>&1; grep -E "gemm|conv" gpurun_out/r2_ops_time19.txt
timeout 400 python bench.py --quick --steps 4 --warmup 3 2>/dev/null | tail -1 | cut -c1-100
timeout 400 python bench.py --quick --steps 4 --warmup 3 --residual-mma 0 2>/dev/null | tail -1 | cut -c1-100
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_config2_gpu.py -m gpu -q --timeout 400 -p no:cacheprovider -k "conv or gemm or unet or config2 or mmhaa" 2>&1 | tail -12 > gpurun_out/r2_pytest19.log
cat gpurun_out/r2_pytest19.log | tail -8
