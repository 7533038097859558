"""block 64 / 128 along the rows of a weight: the shared-memory staged kernel (path 6) vs the generic kernel"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _ops
def tm(fn):
    """device time: 10 calls per CUDA graph (the Python side of a call is ~90 us, longer than the kernels)"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 50
for shape in ((4096, 4096), (11008, 4096), (4096, 11008)):
    W = torch.randn(*shape, device='cuda') * 0.02
    for bs in (16, 64, 128):
        r = _ops.mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 0, bs)
        st = torch.zeros(1, dtype=torch.int32, device='cuda')
        t = tm(lambda: _ops.mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 0, bs, check=False, status=st))
        print("%s bs %3d path %d: %.3f ms  %.0f GB/s (%.2f of 6543)" % (shape, bs, r.path, t, W.numel() * 8 / t / 1e6, W.numel() * 8 / t / 1e6 / 6543.1), flush=True)
