cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_host.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('VALUE', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])
    else: print(l, end='')
"
