cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1
grep -v "^$" gpurun_out/pytest_gpu.txt | grep -E "^(FAILED|ERROR)|passed|failed" | tail -30
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -3 gpurun_out/bench_n1.err
python scripts/show_bench.py gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
