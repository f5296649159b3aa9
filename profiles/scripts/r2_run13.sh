#!/bin/bash
# GPU run 13: shared-memory address space fix in gemm_tc (LDS instead of generic LD for the epilogue vectors)
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_config2_gpu.py -m gpu -q --timeout 400 -p no:cacheprovider -k "conv or gemm or unet or config2 or mmhaa" 2>&1 | tail -12 > gpurun_out/r2_pytest13.log
cat gpurun_out/r2_pytest13.log | tail -8
timeout 300 python profiles/run_ops.py --time > gpurun_out/r2_ops_time13.txt 2>&1; grep -E "gemm|conv" gpurun_out/r2_ops_time13.txt
timeout 300 python profiles/run_ops.py --time --general-epilogue > gpurun_out/r2_ops_time13_general.txt 2>&1; grep -E "gemm|conv" gpurun_out/r2_ops_time13_general.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ops-out gpurun_out/r2_ops_step13.txt > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err
tail -5 gpurun_out/r2_bench13.err | cut -c1-150; cat gpurun_out/r2_bench13.json | cut -c1-400
