cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_quant.py tests/test_gpu_pack.py tests/test_gpu_gptq.py tests/test_ref_native.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('VALUE', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
    else: print(l, end='')
"
timeout 600 ncu --clock-control none -k regex:k_strided -s 7 -c 1 --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>&1 | grep -E "k_strided|duration|inst_executed|issue_active|warps_active|dram_throughput|registers"
