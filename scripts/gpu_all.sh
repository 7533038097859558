cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('VALUE', d['value'], 'frac', d['roofline']['frac'], 'clocks', d['clocks'])
print('extra', json.dumps(d.get('extra'), indent=0)[:1500])
print('cpu', d.get('cpu_baseline'), 'e2e', d['e2e'])
PY
tail -3 gpurun_out/bench_n1.err
