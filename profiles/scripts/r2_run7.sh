#!/bin/bash
# GPU run 7: transposed epilogue stores -- kernel / model tests, op timings, bench
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_frame_shard_gpu.py tests/test_config2_gpu.py tests/test_conditioning_gpu.py -m gpu -q --timeout 400 -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2_pytest7.log
cat gpurun_out/r2_pytest7.log | tail -10
timeout 300 python profiles/run_ops.py --time > gpurun_out/r2_ops_time7.txt 2>&1; cat gpurun_out/r2_ops_time7.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ops-out gpurun_out/r2_ops_step7.txt > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
tail -40 gpurun_out/r2_bench7.err | cut -c1-150; cat gpurun_out/r2_bench7.json
