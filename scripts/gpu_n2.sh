cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t_dist.txt 2>&1; grep -v "^$" gpurun_out/t_dist.txt | grep -E "^(FAILED|ERROR)|passed|failed|Error" | tail -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -5 gpurun_out/bench_n2.err
python scripts/show_bench.py gpurun_out/bench_n2.json | cut -c1-1500
