"""W4A16 below the compute-bound sizes: split-activation block-scaled path vs the producer-dequantizing bf16 kernel vs cuBLAS"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
dev = 'cuda'
def tm(fn):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 100 * 1e3
for KN in (4096, 8192):
    W = torch.randn(KN, KN, device=dev) * 0.02
    pw = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
    wb = pw.dequantize(torch.bfloat16)
    for M in (1, 8, 16, 64, 256, 511):
        x = torch.randn(M, KN, device=dev).to(torch.bfloat16)
        t_c = tm(lambda: torch.matmul(x, wb.t()))
        _tc.SPLIT_A16 = False
        t_old = tm(lambda: _tc.packed_linear(x, pw))
        _tc.SPLIT_A16 = True
        t_new = tm(lambda: _tc.packed_linear(x, pw))
        y = _tc.packed_linear(x, pw); ref = x.float() @ pw.dequantize().t()
        print("KN %d M %3d: cuBLAS bf16 %.1f us | k_gemm_w4 %.1f | split-activation block-scaled %.1f | rel err %.1e" % (
            KN, M, t_c, t_old, t_new, float((y - ref).abs().max() / ref.abs().max())), flush=True)
