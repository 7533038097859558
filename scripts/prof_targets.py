"""small driver for ncu captures of the round-2 kernels: one multi-tensor quantize step, one GPTQ solve of a 4096x4096
linear (native factorization + two-warp column kernel + tensor-core trailing update), one row-block packed GEMM"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import msq_b200
from msq_b200 import _ops, _tc

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
if what in ("all", "quant"):
    shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
    Ws = [torch.randn(r, c, device=dev, generator=g) * 0.02 for r, c in shapes]
    outs = [torch.empty_like(w) for w in Ws]
    plan = _ops.MultiQuantPlan(Ws, outs, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 16)
    for _ in range(3):
        plan.launch()
    torch.cuda.synchronize()
    p = msq_b200.pack_mx_outlier(Ws[4], 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16)
    torch.cuda.synchronize()
    del Ws, outs, p
if what in ("all", "gptq"):
    msq_b200.GPTQ.verbose = False
    layer = torch.nn.Linear(4096, 4096, bias=False, device=dev, dtype=torch.float16)
    layer.weight.data = (torch.randn(4096, 4096, device=dev, generator=g) * 0.02).half()
    X = torch.randn(1, 8192, 4096, device=dev, generator=g).half()
    for it in range(2):
        gq = msq_b200.GPTQ(layer)
        gq.release_cache_on_free = False
        gq.quantizer = msq_b200.MXQuantizer()
        gq.quantizer.configure(8, 8, "fp4_e2m1", "fp8_e4m3", axes=[0], block_size=16)
        for _ in range(4):
            gq.add_batch(X, None)
        gq.fasterquant()
        gq.free()
    torch.cuda.synchronize()
if what in ("all", "gemm"):
    W = torch.randn(4096, 4096, device=dev, generator=g) * 0.02
    pw = msq_b200.pack_mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16)
    pw.hi_index()
    for M in (16, 256):
        x = torch.randn(M, 4096, device=dev, generator=g).to(torch.bfloat16)
        for _ in range(2):
            _tc.packed_linear(x, pw)
    torch.cuda.synchronize()
    pk = msq_b200.pack_mx_outlier(W, 8, 8, "fp4_e2m1", "fp4_e2m1", 5, [1], 32, variant=1)
    for M in (16, 256):
        xq = msq_b200.pack_mx(torch.randn(M, 4096, device=dev, generator=g), 8, "fp8_e4m3", axes=[-1], block_size=32)
        _tc.packed_linear_mx8(xq, pk)
    torch.cuda.synchronize()
print("done", what)
