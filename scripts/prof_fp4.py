import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
KN = int(os.environ.get("KN", "8192")); M = int(os.environ.get("M", "16"))
W = torch.randn(KN, KN, device='cuda') * 0.02
pw = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
xq = msq_b200.pack_mx(torch.randn(M, KN, device='cuda'), 8, 'fp8_e4m3', axes=[-1], block_size=32)
for _ in range(3): _tc.packed_linear_mx8(xq, pw)
torch.cuda.synchronize()
