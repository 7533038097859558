cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r02d_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; echo rc=$?
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02d_bench_ref.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02d_bench_under_ncu.log 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02d_bench_n1.json').read().strip().splitlines()[-1])
print('VALUE', d['value'], 'frac', d['roofline']['frac'], 'steps', d['steps'], 'clocks', d['clocks'])
print('pack', d['extra'].get('pack'))
print('mx4', d['extra'].get('blockscaled_mxfp4_gemm_w4a4'))
print('act', d['extra'].get('activation_quant'))
print('cpu', d.get('cpu_baseline'), 'e2e', d['e2e'])
r = json.loads(open('gpurun_out/r02d_bench_ref.json').read().strip().splitlines()[-1]); print('REF', r['value'], r['cpu_baseline'])
PY
