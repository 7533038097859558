cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_strided -s 7 -c 1 -o gpurun_out/prof_strided_cur python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
