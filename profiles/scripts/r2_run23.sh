#!/bin/bash
# GPU run 23: 256-query attention kernel (flag 15) and packed fp32 softmax math (flag 16): parity, op timing A/B, step A/B
cd $GRAFT_REPO_ROOT
timeout 420 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "attention_kernel_variants or running_max" 2>&1 | tail -15 | cut -c1-300
rm -f gpurun_out/r2_ops_time23.txt gpurun_out/r2_ab23.txt
for f in "0 0" "0 1" "1 0" "1 1"; do set -- $f
  echo "--attn-q256 $1 --attn-packed $2" | tee -a gpurun_out/r2_ops_time23.txt
  timeout 200 python profiles/run_ops.py --time --attn-q256 $1 --attn-packed $2 attn_d40 attn_d40_self attn_d80 2>&1 | tail -3 | tee -a gpurun_out/r2_ops_time23.txt
done
q() { timeout 400 python bench.py --quick --steps 4 --warmup 3 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-60s %.1f ms  %d MHz %s' % (' '.join(sys.argv[1:]), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" "$@" | tee -a gpurun_out/r2_ab23.txt; }
q
q --attn-q256 1 --attn-packed 1
q --attn-packed 1
q --attn-q256 1
q
q --attn-q256 1 --attn-packed 1
