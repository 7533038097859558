cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# launch list of the bench command (timed region only: no extras, no CPU baseline), device time per launch
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_list.log 2>&1
# the bench kernel, full set, one 4096x4096 launch and one 11008x4096 launch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_strided -s 7 -c 5 -f -o gpurun_out/prof_strided_r01 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_full.log 2>&1
timeout 600 python scripts/gpu_sweep.py > gpurun_out/sweep.txt 2>&1
ls -la gpurun_out | tail -5
