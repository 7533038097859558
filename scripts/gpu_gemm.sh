cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -15
timeout 300 python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from msq_b200 import _tc
for (T,K,dt) in [(2048,4096,torch.float16),(2048,11008,torch.float16),(16384,4096,torch.float16)]:
    X = torch.randn(T,K,device='cuda').to(dt); H = torch.zeros(K,K,device='cuda')
    for sym in (False, True):
        for _ in range(3): _tc.hessian_accumulate(H, X, 1.0, 0.5, symmetric=sym)
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): _tc.hessian_accumulate(H, X, 1.0, 0.5, symmetric=sym)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/10
        print(f'hessian T={T} K={K} sym={sym}: {ms:.3f} ms  {2*T*K*K/ms/1e9:.1f} TFLOP/s (full-product flops)')
    A = X.t().contiguous()
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): ref = A.float() @ A.float().t() if False else torch.matmul(A, A.t())
    e1.record(); torch.cuda.synchronize(); print('  cublas fp16 same product:', e0.elapsed_time(e1)/10, 'ms')
PY
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('VALUE', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'])
"
