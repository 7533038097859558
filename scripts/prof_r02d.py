"""one call each of the kernels added in the last part of round 2 (for ncu): quantize+pack with the conversion-
instruction core on an 11008 x 4096 weight, W4A4 kind::mxf4 GEMM (two-CTA form) at 8192^3"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
W = torch.randn(11008, 4096, device='cuda') * 0.02
for _ in range(3):
    p = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp8_e4m3', 2, [0], 16, check=False)
torch.cuda.synchronize()
del W, p
n = 8192
W = torch.randn(n, n, device='cuda') * 0.02
pw = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
xq = msq_b200.pack_mx(torch.randn(n, n, device='cuda'), 8, 'fp4_e2m1', axes=[-1], block_size=32)
for _ in range(3):
    y = _tc.packed_linear_mx4(xq, pw)
torch.cuda.synchronize()
