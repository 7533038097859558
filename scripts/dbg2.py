import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import torch, msq_b200
from msq_b200 import _ops
blk = torch.tensor([-0.012943864800035954, 0.007606775965541601, -0.006049501709640026, -0.003153566736727953, 0.0262801144272089, -0.02341817505657673, -0.028241610154509544, 0.025167010724544525, 0.007884628139436245, 0.06395717710256577, -0.030391091480851173, 0.02073526754975319, -0.0056556290946900845, 0.0020561297424137592, 0.0028521972708404064, -0.03499250486493111])
for cols in (32, 256, 11008):
    W = blk.reshape(16, 1).repeat(1, cols).contiguous().cuda()
    r = _ops.mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 0, 16, want=("deq",))
    print("cols", cols, "path", r.path, "deq col0", [float(v) for v in r.deq[:3, 0]], "distinct", int((r.deq[0] != r.deq[0, 0]).sum()))
    r = _ops.mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 0, 16, want=("deq", "blk_cnt"))
    print("   with cnt: path", r.path, float(r.deq[0, 0]), int(r.blk_cnt[0]))
    p = msq_b200.pack_mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16)
    print("   pack: path", p.path, float(p.dequantize()[0, 0]), int(p.outlier_mask()[0, 0]))
