#!/bin/bash
# Final 1-GPU evidence of round 2: full GPU test suite, smoke, default bench line (+ CPU baseline), reference arm,
# strict tensor-core mode, ncu launch list of one step, ncu --set full of the hot kernels.  Reports are summarised on the
# box (gpurun_out/ is capped at 64 MiB: only the attention report, with sources, comes back whole).
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -rP -p no:cacheprovider --durations=8 2>&1 | grep -v "^$" > gpurun_out/r2_pytest_gpu_full.log
tail -14 gpurun_out/r2_pytest_gpu_full.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
timeout 900 python bench.py --ops-out gpurun_out/r2_ops_step_final.txt > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
tail -3 gpurun_out/r2_bench_n1_final.err | cut -c1-200; cat gpurun_out/r2_bench_n1_final.json
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; cat gpurun_out/r2_bench_reference_arm.json
timeout 300 python bench.py --quick --strict --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2_bench_strict.json; cut -c1-300 gpurun_out/r2_bench_strict.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --ncu-step --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_ncu_step.err
tail -2 gpurun_out/r2_ncu_step.err; wc -l gpurun_out/r2_launches_step.csv
python profiles/summarize_launches.py gpurun_out/r2_launches_step.csv gpurun_out/r2_launches_step --all > /dev/null 2>&1; head -24 gpurun_out/r2_launches_step.md; rm -f gpurun_out/r2_launches_step.csv
OPS="attn_d40 attn_d40_self attn_d80 gemm_960x320 gemm_320x320_res gemm_geglu_2560x320 gemm_320x1280_res gemm_320x968_res conv3x3_320_320 conv3x3_1280_1280 conv3x3_320_320_stride2 conv3x3_640_640_upsample2x groupnorm_320_silu groupnorm_1280_silu layernorm_320 temporal_attn_d40 audio_attention_fused_d40"
timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r2_ops_final python profiles/run_ops.py $OPS > gpurun_out/r2_ncu_ops_final.log 2>&1
python profiles/summarize_ncu.py /tmp/r2_ops_final.ncu-rep gpurun_out/r2_ops_ncu.md ${OPS//_silu/_silu*2} > /dev/null 2>&1; cat gpurun_out/r2_ops_ncu.md | cut -c1-260
ncu -i /tmp/r2_ops_final.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2_ops_ncu_raw.csv.gz
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_attn_d40_final python profiles/run_ops.py attn_d40 > /dev/null 2>&1
ls -la gpurun_out/ | tail -20; du -sh gpurun_out
