cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
  tail -c 6000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 1500 gpurun_out/bench_ref.json
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?
  tail -c 3000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
fi
