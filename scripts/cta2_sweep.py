"""where the two-CTA form wins: token counts x weight sizes, plain FP4 weight (graph-timed)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
def tm(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 50 * 1e3
for KN in (4096, 8192):
    W = torch.randn(KN, KN, device='cuda') * 0.02
    pw = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
    for M in (64, 128, 256, 512, 1024, 2048, 4096):
        xq = msq_b200.pack_mx(torch.randn(M, KN, device='cuda'), 8, 'fp8_e4m3', axes=[-1], block_size=32)
        os.environ.pop("MSQ_MX8T_CTA2", None)
        b1 = _tc.N.lib().msq_gemm_mxfp8_tma_tile_tokens(M, KN, KN)
        t1 = tm(lambda: _tc.packed_linear_mx8(xq, pw))
        os.environ["MSQ_MX8T_CTA2"] = "1"
        b2 = _tc.N.lib().msq_gemm_mxfp8_tma_tile_tokens(M, KN, KN)
        t2 = tm(lambda: _tc.packed_linear_mx8(xq, pw))
        os.environ.pop("MSQ_MX8T_CTA2", None)
        print("KN %d M %4d: 1-CTA (bn %3d) %.1f us | 2-CTA (bn %3d) %.1f us" % (KN, M, b1, t1, b2, t2), flush=True)
