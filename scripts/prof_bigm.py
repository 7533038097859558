"""one 8192^3 and one 4096^3 call of the TMA-fed block-scaled GEMM (for ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
for n in (4096, 8192):
    W = torch.randn(n, n, device='cuda') * 0.02
    pw = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
    xq = msq_b200.pack_mx(torch.randn(n, n, device='cuda'), 8, 'fp8_e4m3', axes=[-1], block_size=32)
    for _ in range(3):
        y = _tc.packed_linear_mx8(xq, pw)
    torch.cuda.synchronize()
