"""W4A4 on kind::mxf4 (msq_gemm_mxfp4_tma) against the fp32 matmul of the dequantized operands, and against the
kind::mxf8f6f4 kernel fed the same FP4 weight with MX-FP8 activations (time only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
torch.backends.cuda.matmul.allow_tf32 = False
def tm(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
ok = True
shapes = [(16, 128, 256), (100, 200, 512), (256, 256, 256), (300, 384, 1024), (512, 1000, 2048), (2048, 1024, 512),
          (4096, 4096, 4096), (8192, 8192, 8192), (4096, 11008, 4096), (4096, 4096, 11008)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (M, Nn, K) in shapes:
    g = torch.Generator(device='cuda').manual_seed(M + Nn + K)
    W = torch.randn(Nn, K, device='cuda', generator=g) * 0.02
    x = torch.randn(M, K, device='cuda', generator=g)
    x *= torch.exp2(torch.randint(-3, 4, (M, 1), device='cuda', generator=g).float())
    pw = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
    x4 = msq_b200.pack_mx(x, 8, 'fp4_e2m1', axes=[-1], block_size=32)
    os.environ["MSQ_MX8T_CTA2"] = "0"
    y1 = _tc.packed_linear_mx4(x4, pw)
    os.environ["MSQ_MX8T_CTA2"] = "1"
    y2 = _tc.packed_linear_mx4(x4, pw) if M >= 32 else y1
    os.environ.pop("MSQ_MX8T_CTA2")
    y = _tc.packed_linear_mx4(x4, pw)
    torch.cuda.synchronize()
    if not (torch.equal(y1, y2) and (torch.equal(y, y1))):
        ok = False
        print("   one-CTA / two-CTA results differ: max abs %g" % float((y1 - y2).abs().max()))
    ref = x4.dequantize().double() @ pw.dequantize().double().t()
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    good = err < 2e-6
    ok = ok and good
    line = "M %5d N %5d K %5d: max rel err vs fp64 matmul of the dequantized operands %.2e %s" % (M, Nn, K, err, "ok" if good else "WRONG")
    if M >= 2048:
        os.environ["MSQ_MX8T_CTA2"] = "0"
        t41 = tm(lambda: _tc.packed_linear_mx4(x4, pw))
        os.environ["MSQ_MX8T_CTA2"] = "1"
        t4 = tm(lambda: _tc.packed_linear_mx4(x4, pw))
        os.environ.pop("MSQ_MX8T_CTA2")
        line += " | mxf4 one CTA %.3f ms %.0f TF" % (t41, 2.0 * M * Nn * K / t41 / 1e9)
        x8 = msq_b200.pack_mx(x, 8, 'fp8_e4m3', axes=[-1], block_size=32)
        t8 = tm(lambda: _tc.packed_linear_mx8(x8, pw))
        fl = 2.0 * M * Nn * K
        line += " | two CTAs %.3f ms %.0f TF | mxf8f6f4 (A8) %.3f ms %.0f TF" % (t4, fl / t4 / 1e9, t8, fl / t8 / 1e9)
    print(line, flush=True)
print("all ok:", ok)
sys.exit(0 if ok else 1)
