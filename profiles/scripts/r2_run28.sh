#!/bin/bash
# GPU run 28 (2 GPUs): bench.py config 2 on 2 ranks with the 256-query attention default (in-run latents parity vs rank 0 alone)
cd $GRAFT_REPO_ROOT
export NCCL_DEBUG=WARN
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_q256.json 2> gpurun_out/r2_bench_n2_q256.err
grep -E "first video|falling back|Error|error|parity" gpurun_out/r2_bench_n2_q256.err | tail -4; cat gpurun_out/r2_bench_n2_q256.json
