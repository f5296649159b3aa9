#!/bin/bash
# Multi-GPU run on N GPUs of one box:  bash profiles/scripts/r2_run_multi.sh N
# (a) every DenoiseLoop schedule at N ranks vs the oracle loop (tests/multigpu_frame_shard.py under torchrun, collected by
#     pytest), (b) bench.py config 2 at N ranks with the in-run latents parity, (c) at N = 8 also BASELINE configs 4 and 5.
N=$1
cd $GRAFT_REPO_ROOT
export NCCL_DEBUG=WARN
timeout 900 python -m pytest tests/test_frame_shard_gpu.py -m gpu -q --timeout 800 -rP -p no:cacheprovider -k "multi_gpu and ${N}]" 2>&1 | grep -E "rel-L2|PARITY|passed|failed|error|Error" | tail -40 > gpurun_out/r2_multigpu_parity_n${N}.log
tail -25 gpurun_out/r2_multigpu_parity_n${N}.log
run() { # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus $N --steps 3 --warmup 3 $2 > gpurun_out/r2_bench_n${N}$1.json 2> gpurun_out/r2_bench_n${N}$1.err
  grep -E "first video|falling back|Error|error" gpurun_out/r2_bench_n${N}$1.err | tail -4; cat gpurun_out/r2_bench_n${N}$1.json
}
run "" ""
if [ "$N" = "8" ]; then
  run "_config4" "--config 4"
  run "_config5" "--config 5"
fi
