#!/bin/bash
# GPU run 3: new tests, bench line, whole-step A/B of the round-2 switches, ncu launch list + op captures
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_conditioning_gpu.py tests/test_config2_gpu.py "tests/test_unet_gpu.py::test_thirty_step_denoise_latent_psnr" -m gpu -q --timeout 300 -rP -p no:cacheprovider 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_pytest3.log
cat gpurun_out/r2_pytest3.log | tail -25
timeout 600 python bench.py --steps 3 --warmup 3 --ops-out gpurun_out/r2_ops_step2.txt > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
tail -42 gpurun_out/r2_bench2.err; cat gpurun_out/r2_bench2.json
for f in "--fuse-ln 0" "--attn-v2 0" "--gn-split 0" "--geglu-exact 1" "--conv-implicit 0"; do
  echo "== $f"; timeout 300 python bench.py --quick --steps 3 --warmup 2 $f 2>/dev/null | tee -a gpurun_out/r2_ab.jsonl
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --ncu-step --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_ncu_step.err
tail -3 gpurun_out/r2_ncu_step.err; wc -l gpurun_out/r2_launches_step.csv
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_ops python profiles/run_ops.py attn_d40 groupnorm_320_silu groupnorm_1280_silu row_stats_320 gemm_960x320 gemm_960x320_lnfused gemm_geglu_2560x320 conv3x3_320_320_stride2 conv3x3_640_640_upsample2x gemm_320x320_res > gpurun_out/r2_ncu_ops.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_ops_attn_v1 python profiles/run_ops.py --attn-v1 attn_d40 > gpurun_out/r2_ncu_ops_v1.log 2>&1
ls -la gpurun_out/*.ncu-rep
