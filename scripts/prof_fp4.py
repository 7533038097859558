import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
W = torch.randn(4096, 4096, device='cuda') * 0.02
pw = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
xq = msq_b200.pack_mx(torch.randn(256, 4096, device='cuda'), 8, 'fp8_e4m3', axes=[-1], block_size=32)
_tc.FP4_TMA = True
for _ in range(2): _tc.packed_linear_mx8(xq, pw)
torch.cuda.synchronize()
_tc.FP4_TMA = False
for _ in range(2): _tc.packed_linear_mx8(xq, pw, cache_dense=True)
torch.cuda.synchronize()
