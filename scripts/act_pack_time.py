"""activation side of the block-scaled linears: quantize + pack of [M, K] activations along K (block 32, plain MX) --
device time (graph replays) next to the GEMM it feeds"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
def tm(fn):
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
for (M, K) in ((16, 4096), (256, 4096), (4096, 4096), (8192, 8192)):
    W = torch.randn(K, K, device='cuda') * 0.02
    pw = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
    line = "M %5d K %5d:" % (M, K)
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(M, K, device='cuda').to(dt)
        for fmt, w in (("fp4_e2m1", 0.5), ("fp8_e4m3", 1.0)):
            t = tm(lambda: msq_b200.pack_mx(x, 8, fmt, axes=[-1], block_size=32))
            nb = M * K * (x.element_size() + w + 1.0 / 32)
            line += "  %s->%s %.1f us (%.0f GB/s)" % (str(dt).split('.')[-1], fmt[:3], t * 1e3, nb / t / 1e6)
    x4 = msq_b200.pack_mx(torch.randn(M, K, device='cuda'), 8, 'fp4_e2m1', axes=[-1], block_size=32)
    x8 = msq_b200.pack_mx(torch.randn(M, K, device='cuda'), 8, 'fp8_e4m3', axes=[-1], block_size=32)
    t4 = tm(lambda: _tc.packed_linear_mx4(x4, pw))
    t8 = tm(lambda: _tc.packed_linear_mx8(x8, pw))
    line += " | GEMM mxf4 %.1f us, mxf8 %.1f us" % (t4 * 1e3, t8 * 1e3)
    print(line, flush=True)
