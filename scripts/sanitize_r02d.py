"""small invocations of the kernels added in the last part of round 2 (for compute-sanitizer)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
torch.manual_seed(0)
W = torch.randn(1040, 512, device='cuda') * 0.02
p = msq_b200.pack_mx_outlier(W[:1024].contiguous(), 8, 8, 'fp4_e2m1', 'fp8_e4m3', 2, [0], 16, check=False)
assert p.path == 5
y = p.dequantize(); yb = p.dequantize(torch.bfloat16)
ref = msq_b200.quantize_mx_outlier_v1(W[:1024].contiguous(), 8, 8, 'fp4_e2m1', 'fp8_e4m3', 'max', 2, [0], 16)
assert torch.equal(y, ref)
for dt in (torch.float32, torch.bfloat16, torch.float16):
    x = torch.randn(33, 7, 96, device='cuda').to(dt)
    for fmt in ('fp4_e2m1', 'fp8_e4m3'):
        a = msq_b200.pack_mx(x, 8, fmt, axes=[-1], block_size=32)
        assert a.path == 7
Wk = torch.randn(384, 512, device='cuda') * 0.02
pw = msq_b200.pack_mx(Wk, 8, 'fp4_e2m1', axes=[1], block_size=32)
for M in (5, 300):
    x4 = msq_b200.pack_mx(torch.randn(M, 512, device='cuda'), 8, 'fp4_e2m1', axes=[-1], block_size=32)
    y = _tc.packed_linear_mx4(x4, pw)
Wk = torch.randn(640, 512, device='cuda') * 0.02
pw = msq_b200.pack_mx(Wk, 8, 'fp4_e2m1', axes=[1], block_size=32)
x4 = msq_b200.pack_mx(torch.randn(300, 512, device='cuda'), 8, 'fp4_e2m1', axes=[-1], block_size=32)
y = _tc.packed_linear_mx4(x4, pw)                       # two-CTA form
x8 = msq_b200.pack_mx(torch.randn(300, 512, device='cuda'), 8, 'fp8_e4m3', axes=[-1], block_size=32)
y = _tc.packed_linear_mx8(x8, pw)
torch.cuda.synchronize()
print("done")
