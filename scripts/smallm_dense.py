import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
dev = 'cuda'
def tm(fn):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 200 * 1e3
for KN in (4096, 8192):
    W = torch.randn(KN, KN, device=dev) * 0.02
    pw = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
    pw2 = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
    xq = msq_b200.pack_mx(torch.randn(16, KN, device=dev), 8, 'fp8_e4m3', axes=[-1], block_size=32)
    t4 = tm(lambda: _tc.packed_linear_mx8(xq, pw))
    _tc.packed_linear_mx8(xq, pw2, cache_dense=True)
    t8 = tm(lambda: _tc.packed_linear_mx8(xq, pw2, cache_dense=True))
    xb = torch.randn(16, KN, device=dev).bfloat16(); wb = W.bfloat16()
    tc = tm(lambda: torch.matmul(xb, wb.t()))
    print("KN %d M 16: fp4 tiled %.1f us | dense E4M3 operand (1 B/weight) %.1f us | cuBLAS bf16 %.1f us" % (KN, t4, t8, tc))
