#!/bin/bash
# round 2, last part: --set full captures of k_pack_tma<16, CVT> and of the two-CTA kind::mxf4 GEMM
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pack_tma|k_gemm_mx8_tma2' -c 6 -f -o gpurun_out/r02d_prof python scripts/prof_r02d.py > gpurun_out/r02d_ncu.log 2>&1
tail -2 gpurun_out/r02d_ncu.log
