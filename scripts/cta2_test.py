"""two-CTA (cta_group::2) form of the TMA-fed block-scaled GEMM against the one-CTA kernel: results and time"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
def tm(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
shapes = [(512, 256, 128), (520, 384, 256), (1000, 200, 1024), (2048, 1024, 512), (4096, 4096, 4096), (8192, 8192, 8192), (4096, 11008, 4096)]
for (M, Nn, K) in shapes:
    g = torch.Generator(device='cuda').manual_seed(M + Nn)
    W = torch.randn(Nn, K, device='cuda', generator=g) * 0.02
    x = torch.randn(M, K, device='cuda', generator=g)
    xq = msq_b200.pack_mx(x, 8, 'fp8_e4m3', axes=[-1], block_size=32)
    for name, pw in (("fp4 codes", msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)),
                     ("e4m3 copy", msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1))):
        if name == "e4m3 copy":
            pw._mx8_dense = _tc.widen_weight_mx8(pw)
        os.environ.pop("MSQ_MX8T_CTA2", None)
        y1 = _tc.packed_linear_mx8(xq, pw)
        t1 = tm(lambda: _tc.packed_linear_mx8(xq, pw)) if M >= 4096 else 0.0
        os.environ["MSQ_MX8T_CTA2"] = "1"
        y2 = _tc.packed_linear_mx8(xq, pw)
        torch.cuda.synchronize()
        t2 = tm(lambda: _tc.packed_linear_mx8(xq, pw)) if M >= 4096 else 0.0
        os.environ.pop("MSQ_MX8T_CTA2", None)
        ref = xq.dequantize() @ pw.dequantize().t()
        fl = 2.0 * M * Nn * K
        print("M %d N %d K %d %s: identical %s, max rel err 1-CTA %.1e 2-CTA %.1e | 1-CTA %.3f ms %.0f TF | 2-CTA %.3f ms %.0f TF" % (
            M, Nn, K, name, bool(torch.equal(y1, y2)), float((y1 - ref).abs().max() / ref.abs().max()),
            float((y2 - ref).abs().max() / ref.abs().max()), t1, fl / t1 / 1e9 if t1 else 0, t2, fl / t2 / 1e9 if t2 else 0), flush=True)
