#!/bin/bash
# ncu --set full of the TMA-staged quantizer kernels (one 11008x4096 weight)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_pack_tma|k_strided_tma' -s 2 -c 2 -f -o gpurun_out/r02_prof_tma python scripts/prof_tma.py > gpurun_out/r02_ncu_tma.log 2>&1
tail -3 gpurun_out/r02_ncu_tma.log
