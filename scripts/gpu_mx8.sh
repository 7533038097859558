cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_mx8.py -x -q 2>&1 | tail -15
timeout 200 python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
import msq_b200
from msq_b200 import _tc
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
for K in (4096, 8192):
    W = torch.randn(K, K, device='cuda')*0.02
    pw = msq_b200.pack_mx_outlier(W, 8,8,'fp4_e2m1','fp4_e2m1',5,[1],32, variant=1)
    print('mx8 ok:', _tc.mx8_weight_ok(pw))
    for Mx in (16, 256, 1024, 4096, 8192):
        x = torch.randn(Mx, K, device='cuda')
        xq = msq_b200.pack_mx(x, 8, 'fp8_e4m3', axes=[-1], block_size=32)
        ms = timeit(lambda: _tc.packed_linear_mx8(xq, pw))
        ms2 = timeit(lambda: _tc.packed_linear(x.to(torch.bfloat16), pw))
        print(f'K=N={K} M={Mx}: mxf8 block-scaled {ms:.3f} ms {2*Mx*K*K/ms/1e9:.1f} TFLOP/s | bf16 fused/dispatch {ms2:.3f} ms {2*Mx*K*K/ms2/1e9:.1f} TFLOP/s')
PY
