#!/bin/bash
# GPU run 15 (experiment): K sweep with the epilogue's stores / TMEM loads disabled
cd $GRAFT_REPO_ROOT
for d in 0 1 2 3; do echo "== MMGT_GEMM_DBG=$d"; MMGT_GEMM_DBG=$d timeout 300 python profiles/k_sweep.py 2>&1 | grep -E "K=   64|K=  320|K= 1280"; done > gpurun_out/r2_k_sweep_dbg.txt 2>&1; cat gpurun_out/r2_k_sweep_dbg.txt
