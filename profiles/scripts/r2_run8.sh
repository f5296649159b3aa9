#!/bin/bash
# GPU run 8: tightened TMA producer loop -- conv / gemm / unet tests, op timings, bench
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_config2_gpu.py -m gpu -q --timeout 400 -p no:cacheprovider -k "conv or gemm or unet or config2 or mmhaa" 2>&1 | tail -6 > gpurun_out/r2_pytest8.log
cat gpurun_out/r2_pytest8.log | tail -4
timeout 300 python profiles/run_ops.py --time > gpurun_out/r2_ops_time8.txt 2>&1; grep -E "gemm|conv" gpurun_out/r2_ops_time8.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ops-out gpurun_out/r2_ops_step8.txt > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
tail -30 gpurun_out/r2_bench8.err | cut -c1-150; cat gpurun_out/r2_bench8.json | cut -c1-400
