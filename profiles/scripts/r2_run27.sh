#!/bin/bash
# GPU run 27: softmax row sums from the tensor cores in the 256-query attention kernel (flag 15 = 7 / 8 / 9): parity, op
# timing A/B, step A/B
cd $GRAFT_REPO_ROOT
timeout 420 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "attention_kernel_variants or running_max" 2>&1 | tail -8 | cut -c1-300
rm -f gpurun_out/r2_ops_time27.txt gpurun_out/r2_ab27.txt
for f in 5 7 8 9; do
  echo "--attn-q256 $f" | tee -a gpurun_out/r2_ops_time27.txt
  timeout 200 python profiles/run_ops.py --time --attn-q256 $f attn_d40 attn_d40_self 2>&1 | tail -2 | tee -a gpurun_out/r2_ops_time27.txt
done
q() { timeout 400 python bench.py --quick --steps 4 --warmup 3 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-60s %.1f ms  %d MHz %s' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" "$@" | tee -a gpurun_out/r2_ab27.txt; }
q
q --attn-q256 7
q --attn-q256 8
q --attn-q256 9
q
q --attn-q256 7
