cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('VALUE', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'])
"
