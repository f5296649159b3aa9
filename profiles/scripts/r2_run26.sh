#!/bin/bash
# GPU run 26: LayerNorm grid variants (flag 17) and suspend-time hints in the 256-query attention kernel (flag 15 = 6):
# parity, op timing A/B, step A/B
cd $GRAFT_REPO_ROOT
timeout 420 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "attention_kernel_variants or running_max or layernorm" 2>&1 | tail -8 | cut -c1-300
rm -f gpurun_out/r2_ops_time26.txt gpurun_out/r2_ab26.txt
for f in "--attn-q256 5" "--attn-q256 6" "--attn-q256 1" "--attn-q256 2"; do
  echo "$f" | tee -a gpurun_out/r2_ops_time26.txt
  timeout 200 python profiles/run_ops.py --time $f attn_d40 attn_d40_self 2>&1 | tail -2 | tee -a gpurun_out/r2_ops_time26.txt
done
for f in 0 1 2; do
  echo "--ln-persist $f" | tee -a gpurun_out/r2_ops_time26.txt
  timeout 200 python profiles/run_ops.py --time --ln-persist $f layernorm_320 2>&1 | tail -1 | tee -a gpurun_out/r2_ops_time26.txt
done
q() { timeout 400 python bench.py --quick --steps 4 --warmup 3 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-60s %.1f ms  %d MHz %s' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" "$@" | tee -a gpurun_out/r2_ab26.txt; }
q
q --attn-q256 6
q --ln-persist 1
q --ln-persist 2
q --attn-q256 6 --ln-persist 2
q
