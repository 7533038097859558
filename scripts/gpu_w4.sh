cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_packed_gemm.py -x -q 2>&1 | tail -15
timeout 300 python - <<'PY'
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
    N = K
    W = torch.randn(N, K, device='cuda')*0.02
    pw = msq_b200.pack_mx_outlier(W, 8,8,'fp4_e2m1','fp4_e2m1',5,[1],32, variant=1)
    wb = pw.dequantize(torch.bfloat16)
    for Mx in (1, 16, 64, 256, 511):
        x = torch.randn(Mx, K, device='cuda').to(torch.bfloat16)
        ms = timeit(lambda: _tc.packed_linear(x, pw))
        ms2 = timeit(lambda: torch.matmul(x, wb.t()))
        print(f'K=N={K} M={Mx}: fused packed {ms:.3f} ms {2*Mx*N*K/ms/1e9:.1f} TFLOP/s  | weight bytes/s {pw.nbytes/ms/1e6:.0f} GB/s | cuBLAS bf16 dense {ms2:.3f} ms {2*Mx*N*K/ms2/1e9:.1f} TFLOP/s')
PY
