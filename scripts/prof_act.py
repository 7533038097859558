"""one call each of the activation quantizer (fp32 / bf16 -> FP4) on 8192 x 8192 and of the strided unpack (for ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
x = torch.randn(8192, 8192, device='cuda')
xb = x.to(torch.bfloat16)
for _ in range(2):
    a = msq_b200.pack_mx(x, 8, 'fp4_e2m1', axes=[-1], block_size=32)
    b = msq_b200.pack_mx(xb, 8, 'fp4_e2m1', axes=[-1], block_size=32)
del x, xb, a, b
W = torch.randn(11008, 4096, device='cuda') * 0.02
p = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp8_e4m3', 2, [0], 16, check=False)
for _ in range(2):
    y = p.dequantize()
torch.cuda.synchronize()
