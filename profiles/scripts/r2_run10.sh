#!/bin/bash
# GPU run 10: ncu source-level capture of the small-K GEMMs, TMA-store and lane-store epilogues
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_ops10_tma python profiles/run_ops.py gemm_960x320 gemm_320x320_res > gpurun_out/r2_ncu_ops10.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_ops10_lane python profiles/run_ops.py --lane-stores gemm_960x320 gemm_320x320_res >> gpurun_out/r2_ncu_ops10.log 2>&1
ls -la gpurun_out/*.ncu-rep
