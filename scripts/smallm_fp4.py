"""decode-size block-scaled GEMM (FP4-by-TMA mode): time vs split-K, against cuBLAS bf16 on the dequantized weight"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
dev = 'cuda'
def tm(fn, it=200):
    """device time per call: the calls are captured into one CUDA graph (20 per graph) so that Python / launch overhead
    does not hide a kernel that is shorter than the host side of its own launch"""
    for _ in range(10): fn()
    torch.cuda.synchronize()
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 200 * 1e3
    except Exception as ex:
        print("graph capture failed:", type(ex).__name__, str(ex)[:100])
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
for KN in (4096, 8192):
    W = torch.randn(KN, KN, device=dev) * 0.02
    pw = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
    wb = pw.dequantize(torch.bfloat16)
    try:
        pwp = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
        _tc.packed_linear_mx8(msq_b200.pack_mx(torch.randn(16, KN, device=dev), 8, 'fp8_e4m3', axes=[-1], block_size=32), pwp)
    except Exception as ex:
        print("plain weight path:", type(ex).__name__, str(ex)[:200]); pwp = None
    nout = sum(bin(int(v) & 0xffffffff).count("1") for v in pw.bitmap[:4096].tolist()) / (4096 * 32)
    print("outlier fraction (first 128K elements): %.4f" % nout)
    for M in (1, 16, 32):
        x = torch.randn(M, KN, device=dev)
        xq = msq_b200.pack_mx(x, 8, 'fp8_e4m3', axes=[-1], block_size=32)
        xb = x.to(torch.bfloat16)
        t_cublas = tm(lambda: torch.matmul(xb, wb.t()))
        res = []
        for ks in (0,):
            if ks: os.environ["MSQ_MX8_KSPLIT"] = str(ks)
            else: os.environ.pop("MSQ_MX8_KSPLIT", None)
            res.append("ks%d %.1f" % (ks, tm(lambda: _tc.packed_linear_mx8(xq, pw))))
        os.environ.pop("MSQ_MX8_KSPLIT", None)
        print("KN %d M %d: cuBLAS bf16 %.1f us | fp4-tma %s" % (KN, M, t_cublas, "  ".join(res)))
        if pwp is not None:
            print("      plain FP4 weight (no outliers): %.1f us; outliers per row %.1f" % (tm(lambda: _tc.packed_linear_mx8(xq, pwp)), float((pw.bitmap_bits().sum() if hasattr(pw, 'bitmap_bits') else -KN) ) / KN))
