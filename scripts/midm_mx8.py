"""mid-size token counts (32 < M < 512): TMA-fed kernel with the packed FP4 codes as the operand vs the producer-widen kernel"""
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
    pwp = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
    pwo = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
    wb = pwp.dequantize(torch.bfloat16)
    for M in (33, 64, 128, 256, 384, 512):
        x = torch.randn(M, KN, device=dev)
        xq = msq_b200.pack_mx(x, 8, 'fp8_e4m3', axes=[-1], block_size=32)
        xb = x.to(torch.bfloat16)
        t_c = tm(lambda: torch.matmul(xb, wb.t()))
        _tc.MX8_DENSE_MIN_M = 100000
        t_old = tm(lambda: _tc.packed_linear_mx8(xq, pwp))
        t_oldo = tm(lambda: _tc.packed_linear_mx8(xq, pwo))
        _tc.MX8_DENSE_MIN_M = 1
        t_new = tm(lambda: _tc.packed_linear_mx8(xq, pwp))
        t_newo = tm(lambda: _tc.packed_linear_mx8(xq, pwo))
        print("KN %d M %3d: cuBLAS bf16 %.1f us | producer-widen plain %.1f outl %.1f | TMA-fed (bn %d) FP4 codes %.1f E4M3 copy %.1f" % (
            KN, M, t_c, t_old, t_oldo, _tc.N.lib().msq_gemm_mxfp8_tma_tile_tokens(M, KN, KN), t_new, t_newo), flush=True)
