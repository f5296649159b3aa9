#!/bin/bash
# GPU run 22: persistent attention kernel (flag 14) -- parity vs the one-item kernel, op timing A/B, whole-step A/B
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_frame_shard_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "attention_two_segments or running_max or row_exchange" 2>&1 | tail -15 | cut -c1-300
timeout 200 python profiles/run_ops.py --time attn_d40 attn_d40_self 2>&1 | tail -2 | tee gpurun_out/r2_ops_time22.txt
timeout 200 python profiles/run_ops.py --time --attn-persist attn_d40 attn_d40_self 2>&1 | tail -2 | tee -a gpurun_out/r2_ops_time22.txt
q() { timeout 400 python bench.py --quick --steps 4 --warmup 3 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-60s %.1f ms  %d MHz %s' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" "$@" | tee -a gpurun_out/r2_ab22.txt; }
q
q --attn-persist 0
q
q --attn-persist 0
