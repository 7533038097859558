#!/bin/bash
# round 2, third part: --set full capture of the TMA-fed block-scaled GEMM (k_gemm_mx8_tma) at 4096^3 and 8192^3
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_mx8_tma -c 6 -f -o gpurun_out/r02c_prof_mx8_tma2 python scripts/prof_bigm.py > gpurun_out/r02c_ncu_mx8_2.log 2>&1
tail -2 gpurun_out/r02c_ncu_mx8_2.log
