"""W4A16 at compute-bound sizes: split activations + accumulated terms on the packed codes vs bf16 copy + kind::f16 GEMM"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
def tm(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for KN, M in ((4096, 4096), (4096, 8192), (8192, 4096), (8192, 8192)):
    W = torch.randn(KN, KN, device='cuda') * 0.02
    pw = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
    x = torch.randn(M, KN, device='cuda').to(torch.bfloat16)
    fl = 2.0 * M * KN * KN
    _tc.LARGE_M_WITHOUT_BF16_COPY = True
    t1 = tm(lambda: _tc.packed_linear(x, pw))
    y1 = _tc.packed_linear(x, pw)
    assert getattr(pw, "_dense_bf16", None) is None
    _tc.LARGE_M_WITHOUT_BF16_COPY = False
    t2 = tm(lambda: _tc.packed_linear(x, pw))
    y2 = _tc.packed_linear(x, pw)
    wb = pw.dequantize(torch.bfloat16)
    t3 = tm(lambda: torch.matmul(x, wb.t()))
    print("KN %d M %d: packed codes + split activations %.3f ms %.0f TF | bf16 copy + kind::f16 %.3f ms %.0f TF | cuBLAS bf16 %.3f ms %.0f TF | max rel diff %.1e" % (
        KN, M, t1, fl / t1 / 1e9, t2, fl / t2 / 1e9, t3, fl / t3 / 1e9, float((y1 - y2).abs().max() / y2.abs().max())), flush=True)
