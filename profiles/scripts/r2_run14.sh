#!/bin/bash
# GPU run 14: K sweep of the small-K GEMMs (is the epilogue overlapped with the next tile's main loop?)
cd $GRAFT_REPO_ROOT
timeout 300 python profiles/k_sweep.py > gpurun_out/r2_k_sweep.txt 2>&1; cat gpurun_out/r2_k_sweep.txt
