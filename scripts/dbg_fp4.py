import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
import bench
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
for (N_, K, M, sd) in [(128, 128, 16, 5), (256, 512, 16, 2), (4096, 4096, 16, 5), (4096, 4096, 1, 2), (1024, 2048, 300, 3)]:
    W = torch.randn(N_, K, generator=g) * 0.05
    W.view(-1)[torch.randperm(W.numel(), generator=g)[: W.numel() // 500]] *= 30
    pw = msq_b200.pack_mx_outlier(W.cuda(), 8, 8, "fp4_e2m1", "fp4_e2m1", sd, [1], 32, variant=1)
    x = torch.randn(M, K, generator=g)
    xq = msq_b200.pack_mx(x.cuda(), 8, "fp8_e4m3", axes=[-1], block_size=32)
    assert _tc.mx4_weight_ok(pw)
    y = _tc.packed_linear_mx8(xq, pw).double().cpu()
    xd, wd = xq.dequantize().double().cpu(), pw.dequantize().double().cpu()
    ref = xd @ wd.t()
    bound = 4 * K * 2.0 ** -24 * (xd.abs() @ wd.abs().t()) + 1e-30
    nout = int(pw.outlier_mask().sum())
    print(N_, K, M, "outliers", nout, "max err/bound", float(((y - ref).abs() / bound).max()), "rel", float((y - ref).abs().max() / ref.abs().max()), flush=True)
W = torch.randn(4096, 4096, generator=g) * 0.02
pw = msq_b200.pack_mx_outlier(W.cuda(), 8, 8, "fp4_e2m1", "fp4_e2m1", 5, [1], 32, variant=1)
for M in (1, 16, 64, 256):
    xq = msq_b200.pack_mx(torch.randn(M, 4096).cuda(), 8, "fp8_e4m3", axes=[-1], block_size=32)
    _tc.FP4_TMA = True
    t1 = bench._time_ms(lambda: _tc.packed_linear_mx8(xq, pw), 20)
    _tc.FP4_TMA = False
    t0 = bench._time_ms(lambda: _tc.packed_linear_mx8(xq, pw), 20) if _tc.mx8_weight_ok(pw) else float("nan")
    xb = xq.dequantize().to(torch.bfloat16); wb = pw.dequantize(torch.bfloat16)
    t2 = bench._time_ms(lambda: xb @ wb.t(), 20)
    print("M=%d  fp4-tma %.2f us   producer-widen %.2f us   cuBLAS bf16 %.2f us" % (M, t1 * 1e3, t0 * 1e3, t2 * 1e3), flush=True)
