"""A/B of the persistent TMA-staged fake-quant kernel (k_strided_tma) against the direct-load kernel
(k_strided_multi, MSQ_QUANT_TMA=0): bit-identity on regular and ragged shapes, then GB/s on bench.py's layer."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msq_b200
from msq_b200 import _ops

dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(Ws, bs, fi, fo, tma):
    os.environ["MSQ_QUANT_TMA"] = "1" if tma else "0"
    return _ops.mx_outlier_multi(Ws, 8, 8, fi, fo, 2, bs)


ok = True
shapes = [[(64, 256)], [(16, 4)], [(500, 388), (33, 1028), (4096, 4096)], [(17, 260), (1, 8), (257, 12), (48, 516)],
          [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]]
for bs in (16, 32, 8):
    for fi, fo in (("fp4_e2m1", "fp8_e4m3"), ("int2", "fp4_e2m1"), ("fp6_e3m2", "fp6_e3m2")):
        for shp in shapes:
            Ws = [torch.randn(s, device=dev) * (0.02 if i % 2 == 0 else 3.0) for i, s in enumerate(shp)]
            Ws[0].view(-1)[::7] *= 40.0
            a = run(Ws, bs, fi, fo, True)
            b = run(Ws, bs, fi, fo, False)
            for x, y, s in zip(a, b, shp):
                same = torch.equal(x.view(torch.int32), y.view(torch.int32))
                if not same:
                    ok = False
                    print("MISMATCH bs", bs, fi, fo, s, int((x.view(torch.int32) != y.view(torch.int32)).sum()))
print("bit-identical:", ok)

# timing on the bench layer
shp = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
Ws = [torch.randn(s, device=dev) * 0.02 for s in shp]
outs = [torch.empty_like(w) for w in Ws]
nel = sum(w.numel() for w in Ws)
res = {}
for bs, fi, fo in ((16, "fp4_e2m1", "fp8_e4m3"), (32, "fp6_e3m2", "fp6_e3m2"), (8, "fp4_e2m1", "fp8_e4m3")):
    for tma in (0, 1, 0, 1):
        os.environ["MSQ_QUANT_TMA"] = str(tma)
        plan = _ops.MultiQuantPlan(Ws, outs, 8, 8, fi, fo, 2, bs)
        for _ in range(5):
            plan.launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            plan.launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        print("bs %d %s tma=%d: %.4f ms  %.0f GB/s  frac %.3f" % (bs, fi, tma, ms, nel * 8 / ms / 1e6, nel * 8 / ms / 1e6 / 6543.1))
print("status ok" if ok else "FAILED")
sys.exit(0 if ok else 1)
