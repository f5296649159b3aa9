#!/bin/bash
# GPU run 24: 256-query attention kernel variants (flag 15 = 1..5: early S hand-back, FMA-pipe exp2 for 1/4 or 1/2 of the
# scores): parity, op timing A/B, step A/B
cd $GRAFT_REPO_ROOT
timeout 420 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "attention_kernel_variants or running_max or attention_two_segments" 2>&1 | tail -15 | cut -c1-300
rm -f gpurun_out/r2_ops_time24.txt gpurun_out/r2_ab24.txt
for f in 1 2 3 4 5; do
  echo "--attn-q256 $f" | tee -a gpurun_out/r2_ops_time24.txt
  timeout 200 python profiles/run_ops.py --time --attn-q256 $f attn_d40 attn_d40_self 2>&1 | tail -2 | tee -a gpurun_out/r2_ops_time24.txt
done
q() { timeout 400 python bench.py --quick --steps 4 --warmup 3 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-60s %.1f ms  %d MHz %s' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" "$@" | tee -a gpurun_out/r2_ab24.txt; }
q --attn-q256 1
q --attn-q256 2
q --attn-q256 3
q --attn-q256 4
q --attn-q256 1
q --attn-q256 3
