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
    W2 = W.clone(); W2.view(-1)[torch.randperm(W2.numel(), device=dev)[:KN]] *= 30
    pwo2 = msq_b200.pack_mx_outlier(W2, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
    wb = pwp.dequantize(torch.bfloat16)
    for M in (64, 256, 512, 4096):
        x = torch.randn(M, KN, device=dev)
        xq = msq_b200.pack_mx(x, 8, 'fp8_e4m3', axes=[-1], block_size=32)
        xb = x.to(torch.bfloat16)
        t_c = tm(lambda: torch.matmul(xb, wb.t()))
        _tc.MX8_TMA = False; _tc.MX8_DENSE_MIN_M = 100000
        t_old = tm(lambda: _tc.packed_linear_mx8(xq, pwp))
        t_oldo = tm(lambda: _tc.packed_linear_mx8(xq, pwo))
        _tc.MX8_TMA = True
        _tc.MX8_DENSE_MIN_M = 512
        t_new = tm(lambda: _tc.packed_linear_mx8(xq, pwp))
        _tc.MX8_TMA_SPARSE_BUDGET = _tc.MX8_TMA_SPARSE_BUDGET_LARGE_M = 1e18
        t_sp = tm(lambda: _tc.packed_linear_mx8(xq, pwo))            # FP4 codes + sparse outlier epilogue (no copy)
        t_sp2 = tm(lambda: _tc.packed_linear_mx8(xq, pwo2))          # the same with ~1 outlier per row
        _tc.MX8_TMA_FUSED_SPARSE = True
        t_fu2 = tm(lambda: _tc.packed_linear_mx8(xq, pwo2))          # ... applied inside the GEMM's epilogue
        _tc.MX8_TMA_FUSED_SPARSE = False
        assert getattr(pwo, "_mx8_dense", None) is None
        t_newo = tm(lambda: _tc.packed_linear_mx8(xq, pwo, cache_dense=True))
        pwo._mx8_dense = None
        print("KN %d M %3d: cuBLAS bf16 %.1f us | producer-widen plain %.1f outl %.1f | TMA-fed (bn %d) FP4 codes %.1f, + sparse outliers (%.3f / row) %.1f (%.2f / row) %.1f [fused in the epilogue: %.1f], E4M3 copy %.1f" % (
            KN, M, t_c, t_old, t_oldo, _tc.N.lib().msq_gemm_mxfp8_tma_tile_tokens(M, KN, KN), t_new, float(pwo.outlier_mask().sum()) / KN, t_sp, float(pwo2.outlier_mask().sum()) / KN, t_sp2, t_fu2, t_newo), flush=True)
