cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-extras > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'frac/gpu', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
PY
