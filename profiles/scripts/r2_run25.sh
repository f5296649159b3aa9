#!/bin/bash
# GPU run 25: step A/B of the 256-query attention kernel with 1 of 4 score pairs through the FMA-pipe exp2 (flag 15 = 5),
# source-level ncu profile of that variant
cd $GRAFT_REPO_ROOT
rm -f gpurun_out/r2_ab25.txt
q() { timeout 400 python bench.py --quick --steps 4 --warmup 3 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-60s %.1f ms  %d MHz %s' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" "$@" | tee -a gpurun_out/r2_ab25.txt; }
q --attn-q256 5
q --attn-q256 1
q --attn-q256 5
q --attn-q256 1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_attn_d40_q256v5 python profiles/run_ops.py --attn-q256 5 attn_d40 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
