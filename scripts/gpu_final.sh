cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('VALUE', d['value'], 'frac', d['roofline']['frac'], 'steps', d['steps'], 'clocks', d['clocks'])
print('extra', json.dumps(d.get('extra'))[:1200])
print('cpu', d.get('cpu_baseline'), 'e2e', d['e2e'])
r = json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1]); print('REF', r['value'], r['cpu_baseline'])
PY
