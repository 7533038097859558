cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/mx4_bn_sweep.py > gpurun_out/r02d_mx4_bn_sweep.txt 2>&1; tail -5 gpurun_out/r02d_mx4_bn_sweep.txt | cut -c1-200
timeout 300 python scripts/mx4_test.py > gpurun_out/r02d_mx4_test.txt 2>&1; tail -6 gpurun_out/r02d_mx4_test.txt | cut -c1-60,100-260
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r02d_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; echo rc=$?
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02d_bench_n1.json').read().strip().splitlines()[-1])
print('VALUE', d['value'], 'frac', d['roofline']['frac'], 'steps', d['steps'], 'clocks', d['clocks'])
print('pack', d['extra'].get('pack'))
print('mx4', d['extra'].get('blockscaled_mxfp4_gemm_w4a4'))
print('mx8', json.dumps(d['extra'].get('blockscaled_mxfp8_gemm_k4096_n4096'))[:600])
print('cpu', d.get('cpu_baseline'), 'e2e', d['e2e'])
PY
