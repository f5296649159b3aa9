#!/bin/bash
# GPU run 18: resident-weight vs streaming GEMM kernel on the level-1 / level-2 shapes
cd $GRAFT_REPO_ROOT
timeout 300 python profiles/bres_sweep.py > gpurun_out/r2_bres_sweep.txt 2>&1; cat gpurun_out/r2_bres_sweep.txt
