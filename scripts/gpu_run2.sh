set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_strided -s 7 -c 1 -o gpurun_out/prof_strided_r01b python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
