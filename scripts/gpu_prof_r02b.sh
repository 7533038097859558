#!/bin/bash
# round 2, second half: launch list of the bench command + --set full capture of the bench kernel (k_strided_tma, one launch = 7 linears)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02b_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_strided_tma -s 3 -c 1 -f -o gpurun_out/r02b_prof_strided_tma python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02b_ncu_full.log 2>&1
tail -2 gpurun_out/r02b_ncu_full.log
