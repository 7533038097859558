import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _ops
W = torch.randn(11008, 4096, device='cuda') * 0.02
for _ in range(2):
    p = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp8_e4m3', 2, [0], 16)
    r = _ops.mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp8_e4m3', 2, 0, 16)
torch.cuda.synchronize()
