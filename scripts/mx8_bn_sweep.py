"""token-tile width of the FP8-activation kind::mxf8f6f4 GEMM (two-CTA form): one accumulator of 256 columns against two of <= 240"""
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
    return e0.elapsed_time(e1) / 50
for (M, Nn, K) in ((4096, 4096, 4096), (8192, 8192, 8192), (4096, 11008, 4096), (4096, 4096, 11008), (16384, 4096, 4096)):
    W = torch.randn(Nn, K, device='cuda') * 0.02
    pw = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
    xq = msq_b200.pack_mx(torch.randn(M, K, device='cuda'), 8, 'fp8_e4m3', axes=[-1], block_size=32)
    line = "M %5d N %5d K %5d:" % (M, Nn, K)
    for bn in ("", "256", "240", "224", "208", "192", "128"):
        if bn: os.environ["MSQ_MX8T_BN"] = bn
        else: os.environ.pop("MSQ_MX8T_BN", None)
        t = tm(lambda: _tc.packed_linear_mx8(xq, pw))
        line += "  bn %s %.0f TF" % (bn or "auto(%d)" % _tc.N.lib().msq_gemm_mxfp8_tma_tile_tokens(M, Nn, K), 2.0 * M * Nn * K / t / 1e9)
    os.environ.pop("MSQ_MX8T_BN", None)
    print(line, flush=True)
    del W, pw, xq
