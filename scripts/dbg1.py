import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import torch, msq_b200
from msq_b200 import _tc, _ops
import msq_oracle as O
torch.manual_seed(0)
# --- rowblock debug: x = identity
for (N, K, fi, fo, bs) in [(128, 64, "fp4_e2m1", "fp4_e2m1", 32), (128, 64, "fp4_e2m1", "fp8_e4m3", 16)]:
    g = torch.Generator().manual_seed(1)
    W = (torch.randn(N, K, generator=g) * 0.02).cuda()
    pw = msq_b200.pack_mx_outlier(W, 8, 8, fi, fo, 2, [0], bs)
    x = torch.eye(K, dtype=torch.bfloat16).cuda()
    y = _tc.packed_linear(x, pw)          # [K, N] = W^T
    Wd = pw.dequantize()
    d = (y.t() - Wd)
    bad = (d != 0)
    print(fi, fo, bs, "bad", int(bad.sum()), "of", bad.numel())
    idx = bad.nonzero()[:10]
    for r, c in idx.tolist():
        print("  n", r, "k", c, "got", float(y[c, r]), "want", float(Wd[r, c]), "mask", int(pw.outlier_mask()[r, c]))
    print("  rows with errors:", sorted(set(bad.nonzero()[:, 0].tolist()))[:20], "cols:", sorted(set(bad.nonzero()[:, 1].tolist()))[:20])
# --- deviation debug on down_proj
dev_fix = json.load(open("/root/repo/tests/golden/deviation_bench.json"))
g = torch.Generator().manual_seed(1000)
Ws = [torch.randn(m["rows"], m["cols"], generator=g) * 0.02 for m in dev_fix["matrices"]]
W = Ws[6]
y_multi = _ops.mx_outlier_multi([W.cuda()], 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 16)[0].cpu()
y_single = msq_b200.quantize_mx_outlier_v1(W.cuda(), 8, 8, "fp4_e2m1", "fp8_e4m3", "max", 2, [0], 16).cpu()
torch.set_num_threads(os.cpu_count())
orc = O.quantize_mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 0, 16)
print("multi vs single", int((y_multi != y_single).sum()), "single vs oracle", int((y_single != orc).sum()))
bad = (y_single != orc).nonzero()
for r, c in bad.tolist():
    b = r // 16
    blk = W[b * 16:(b + 1) * 16, c]
    print("pos", r, c, "x", float(W[r, c]), "kernel", float(y_single[r, c]), "oracle", float(orc[r, c]))
    a = blk.abs()
    print(" block", [float(v) for v in blk])
    import numpy as np
    print(" mean|a| fp64", float(a.double().mean()), "std fp64", float(a.double().std(unbiased=False)))
    print(" kernel blk", [float(v) for v in y_single[b*16:(b+1)*16, c]])
    print(" oracle blk", [float(v) for v in orc[b*16:(b+1)*16, c]])
