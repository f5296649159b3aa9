#!/bin/bash
# GPU run 4: kernel + model tests after the MMA-issue (elect.sync) and epilogue-vector changes, bench, whole-step A/B
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_frame_shard_gpu.py tests/test_config2_gpu.py -m gpu -q --timeout 300 -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2_pytest4.log
cat gpurun_out/r2_pytest4.log | tail -12
timeout 300 python profiles/run_ops.py --time > gpurun_out/r2_ops_time4.txt 2>&1; cat gpurun_out/r2_ops_time4.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --ops-out gpurun_out/r2_ops_step4.txt > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
tail -42 gpurun_out/r2_bench4.err; cat gpurun_out/r2_bench4.json
for f in "--fuse-ln 0" "--attn-v2 0" "--gn-split 0" "--deep-batch 1" "--fuse-ln 0 --gn-split 0 --attn-v2 0 --deep-batch 1"; do
  echo "== $f"; timeout 300 python bench.py --quick --steps 3 --warmup 2 $f 2> gpurun_out/r2_ab4.err | tee -a gpurun_out/r2_ab4.jsonl; tail -2 gpurun_out/r2_ab4.err
done
