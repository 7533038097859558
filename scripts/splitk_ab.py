"""split-K of the TMA-fed block-scaled GEMMs between the decode sizes and the compute-bound ones: results against the
fp64 matmul of the dequantized operands, device time with (default) and without (MSQ_MX8T_KSPLIT=0) the K slices"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
torch.backends.cuda.matmul.allow_tf32 = False
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
ok = True
for KN in (4096, 8192):
    W = torch.randn(KN, KN, device='cuda') * 0.02
    pw = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
    wd = pw.dequantize().double()
    for M in (48, 64, 128, 256, 512, 1024):
        x = torch.randn(M, KN, device='cuda')
        line = "KN %d M %4d:" % (KN, M)
        for name, fmt, fn in (("mxf8", "fp8_e4m3", _tc.packed_linear_mx8), ("mxf4", "fp4_e2m1", _tc.packed_linear_mx4)):
            xq = msq_b200.pack_mx(x, 8, fmt, axes=[-1], block_size=32)
            ref = xq.dequantize().double() @ wd.t()
            os.environ.pop("MSQ_MX8T_KSPLIT", None)
            y = fn(xq, pw); torch.cuda.synchronize()
            err = float((y.double() - ref).abs().max() / ref.abs().max())
            t1 = tm(lambda: fn(xq, pw))
            os.environ["MSQ_MX8T_KSPLIT"] = "0"
            y0 = fn(xq, pw); torch.cuda.synchronize()
            err0 = float((y0.double() - ref).abs().max() / ref.abs().max())
            t0 = tm(lambda: fn(xq, pw))
            os.environ.pop("MSQ_MX8T_KSPLIT", None)
            good = err < 2e-6 and err0 < 2e-6
            ok = ok and good
            line += "  %s %.1f us (no split %.1f us) err %.1e/%.1e%s" % (name, t1, t0, err, err0, "" if good else " WRONG")
        xb = x.to(torch.bfloat16); wb = pw.dequantize(torch.bfloat16)
        line += "  | cuBLAS bf16 %.1f us" % tm(lambda: torch.matmul(xb, wb.t()))
        print(line, flush=True)
print("all ok:", ok)
sys.exit(0 if ok else 1)
