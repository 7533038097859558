"""one unpack of the quantizers' artefact (for ncu) + timing of the unpack kernels"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
W = torch.randn(11008, 4096, device='cuda') * 0.02
p = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp8_e4m3', 2, [0], 16, check=False)
def tm(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
out = torch.empty_like(W)
for dt in (torch.float32, torch.bfloat16):
    t = tm(lambda: p.dequantize(dt))
    nb = p.nbytes + W.numel() * (4 if dt == torch.float32 else 2)
    print("unpack strided fp4/fp8 bs16 -> %s: %.1f us  %.0f GB/s  frac %.3f" % (dt, t * 1e3, nb / t / 1e6, nb / t / 1e6 / 6543.1), flush=True)
p2 = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
t = tm(lambda: p2.dequantize(torch.float32))
nb = p2.nbytes + W.numel() * 4
print("unpack inner fp4/fp4 bs32 -> fp32: %.1f us  %.0f GB/s  frac %.3f" % (t * 1e3, nb / t / 1e6, nb / t / 1e6 / 6543.1), flush=True)
