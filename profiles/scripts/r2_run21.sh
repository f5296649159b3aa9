#!/bin/bash
# GPU run 21: exchange test after the fix; attention with long frames first; per-level batching A/B of the whole step
# (level_batch = forwards per batch at the 64x64 / 32x32 / 16x16 / 8x8 level; split = one unit per CFG branch).
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_frame_shard_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider -k "row_exchange" 2>&1 | tail -3
timeout 300 python profiles/run_ops.py --time attn_d40 attn_d40_self attn_d80 2>&1 | tail -3 | tee gpurun_out/r2_ops_time21.txt
q() { timeout 400 python bench.py --quick --steps 4 --warmup 3 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-60s %.1f ms  %d MHz %s' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" "$@" | tee -a gpurun_out/r2_ab21.txt; }
q
q --deep-from 1
q --split-branches 1 --level-batch 1,1,99,99
q --split-branches 1 --level-batch 1,2,99,99
q --split-branches 1 --level-batch 1,4,99,99
q --split-branches 1 --level-batch 2,4,99,99
q --level-batch 1,2,99,99
q --level-batch 1,5,99,99
q
