#!/bin/bash
# GPU run 5: tests (new kernels + f1 ReferenceNet), bench, A/B, ncu of the open questions
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_conditioning_gpu.py -m gpu -q --timeout 300 -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2_pytest5.log
cat gpurun_out/r2_pytest5.log | tail -12
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --ops-out gpurun_out/r2_ops_step5.txt > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
tail -42 gpurun_out/r2_bench5.err | cut -c1-150; cat gpurun_out/r2_bench5.json
for f in "--temporal-rows 0" "--gn-split 0"; do
  echo "== $f"; timeout 300 python bench.py --quick --steps 3 --warmup 2 $f 2> gpurun_out/r2_ab5.err | tee -a gpurun_out/r2_ab5.jsonl; tail -2 gpurun_out/r2_ab5.err
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_ops5 python profiles/run_ops.py attn_d40 gemm_960x320 gemm_320x320_res gemm_geglu_2560x320 groupnorm_320_silu temporal_attn_d40 conv3x3_320_320 > gpurun_out/r2_ncu_ops5.log 2>&1
ls -la gpurun_out/*.ncu-rep
