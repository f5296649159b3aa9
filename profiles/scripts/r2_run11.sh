#!/bin/bash
# GPU run 11: specialised straight-line GEMM epilogues -- tests, op timings A/B, bench A/B
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "specialised" 2>&1 | tail -15 > gpurun_out/r2_pytest11a.log
cat gpurun_out/r2_pytest11a.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_config2_gpu.py -m gpu -q --timeout 400 -p no:cacheprovider -k "conv or gemm or unet or config2 or mmhaa" 2>&1 | tail -12 > gpurun_out/r2_pytest11.log
cat gpurun_out/r2_pytest11.log | tail -8
timeout 300 python profiles/run_ops.py --time > gpurun_out/r2_ops_time11.txt 2>&1; grep -E "gemm|conv" gpurun_out/r2_ops_time11.txt
timeout 300 python profiles/run_ops.py --time --general-epilogue > gpurun_out/r2_ops_time11_general.txt 2>&1; grep -E "gemm|conv" gpurun_out/r2_ops_time11_general.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ops-out gpurun_out/r2_ops_step11.txt > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
tail -5 gpurun_out/r2_bench11.err | cut -c1-150; cat gpurun_out/r2_bench11.json | cut -c1-400
timeout 400 python bench.py --quick --steps 3 --warmup 3 --lean-epilogue 0 2>/dev/null | tail -1 | cut -c1-300
