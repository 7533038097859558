cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?
tail -c 1500 gpurun_out/bench_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 scripts/llama7b_e2e.py > gpurun_out/llama7b_e2e_n8.json 2> gpurun_out/llama7b_e2e_n8.err; echo rc=$?
tail -2 gpurun_out/llama7b_e2e_n8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 scripts/llama7b_e2e.py --nearest > gpurun_out/llama7b_rtn_n8.json 2> gpurun_out/llama7b_rtn_n8.err; echo rc=$?
tail -2 gpurun_out/llama7b_rtn_n8.json
