#!/bin/bash
# GPU run: attention kernel tests, op-level timings (round-2 kernels vs the round-1 ones behind their A/B flags), one bench line
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "attention" -p no:cacheprovider > gpurun_out/r2_pytest_attn.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_attn.log; tail -5 gpurun_out/r2_pytest_attn.log
timeout 300 python profiles/run_ops.py --time > gpurun_out/r2_ops_time.txt 2>&1
timeout 120 python profiles/run_ops.py --time --attn-v1 --gn-fused --conv-im2col --geglu-exact attn_d40 attn_d40_self groupnorm_320_silu groupnorm_1280_silu groupnorm_640_silu_64x64 conv3x3_320_320_stride2 conv3x3_640_640_upsample2x gemm_geglu_2560x320 > gpurun_out/r2_ops_time_round1_kernels.txt 2>&1
cat gpurun_out/r2_ops_time.txt gpurun_out/r2_ops_time_round1_kernels.txt
timeout 600 python bench.py --steps 3 --warmup 3 --ops-out gpurun_out/r2_ops_step1.txt > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -45 gpurun_out/r2_bench1.err; cat gpurun_out/r2_bench1.json
