cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1
grep -v "^$" gpurun_out/pytest_gpu.txt | grep -E "^(FAILED|ERROR)|passed|failed" | tail -60
