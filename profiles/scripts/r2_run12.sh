#!/bin/bash
# GPU run 12: ncu source-level capture of the small-K GEMMs with the specialised epilogues
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_ops12_lean python profiles/run_ops.py gemm_960x320 gemm_320x320_res gemm_geglu_2560x320 gemm_320x1280_res > gpurun_out/r2_ncu_ops12.log 2>&1
ls -la gpurun_out/*.ncu-rep
