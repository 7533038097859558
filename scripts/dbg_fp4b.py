import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, msq_b200
from msq_b200 import _tc
from gpu_smallm import graph_time
W = torch.randn(4096, 4096, device='cuda') * 0.02
pw = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
for Mx in (16, 64, 256):
    xq = msq_b200.pack_mx(torch.randn(Mx, 4096, device='cuda'), 8, 'fp8_e4m3', axes=[-1], block_size=32)
    pw._mx8_dense = None
    _tc.FP4_TMA = True
    t4 = graph_time(lambda: _tc.packed_linear_mx8(xq, pw))
    _tc.FP4_TMA = False
    _tc.packed_linear_mx8(xq, pw, cache_dense=True)
    td = graph_time(lambda: _tc.packed_linear_mx8(xq, pw))
    print("M=%d fp4 %.1f us  cached-dense E4M3 %.1f us" % (Mx, t4 * 1e3, td * 1e3), flush=True)
