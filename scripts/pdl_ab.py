"""programmatic dependent launch A/B: decode-size split-K GEMM (zero-fill kernel -> GEMM) and the TMA-fed kernel
(scale tiling -> GEMM), device time from CUDA-graph replays"""
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
        for _ in range(20): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 200 * 1e3
for KN in (4096, 8192):
    W = torch.randn(KN, KN, device=dev) * 0.02
    pwo = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
    pwp = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
    wb = pwp.dequantize(torch.bfloat16)
    for M in (1, 16, 32, 64, 256):
        x = torch.randn(M, KN, device=dev)
        xq = msq_b200.pack_mx(x, 8, 'fp8_e4m3', axes=[-1], block_size=32)
        xb = x.to(torch.bfloat16)
        t_c = tm(lambda: torch.matmul(xb, wb.t()))
        res = []
        for pdl in (0, 1):
            if pdl: os.environ.pop("MSQ_MX8_NO_PDL", None)
            else: os.environ["MSQ_MX8_NO_PDL"] = "1"
            res.append((tm(lambda: _tc.packed_linear_mx8(xq, pwp)), tm(lambda: _tc.packed_linear_mx8(xq, pwo))))
        os.environ.pop("MSQ_MX8_NO_PDL", None)
        y = _tc.packed_linear_mx8(xq, pwo)
        ref = xq.dequantize() @ pwo.dequantize().t()
        err = float((y - ref).abs().max() / ref.abs().max())
        print("KN %d M %3d: cuBLAS bf16 %.1f us | no PDL: plain %.1f outl %.1f | PDL: plain %.1f outl %.1f | rel err %.1e" % (
            KN, M, t_c, res[0][0], res[0][1], res[1][0], res[1][1], err), flush=True)
