"""decode-size GEMMs, kernel-only (CUDA-graph replay): FP4-by-TMA block-scaled mode vs producer-widen mode vs cached
E4M3 copy vs the row-block bf16 kernel vs dense bf16 cuBLAS on the dequantized weight"""
import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import msq_b200
from msq_b200 import _tc
from gpu_smallm import graph_time  # noqa
for K, Nn in ((4096, 4096), (4096, 11008), (8192, 8192)):
    W = torch.randn(Nn, K, device='cuda') * 0.02
    pw = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
    pr = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp8_e4m3', 2, [0], 16)
    pr.hi_index()
    Wd = pw.dequantize(torch.bfloat16)
    nout = int(pw.outlier_mask().sum())
    for Mx in (1, 16, 64, 256):
        x = torch.randn(Mx, K, device='cuda')
        xb = x.to(torch.bfloat16)
        xq = msq_b200.pack_mx(x, 8, 'fp8_e4m3', axes=[-1], block_size=32)
        pw._mx8_dense = None
        _tc.FP4_TMA = True
        t4 = graph_time(lambda: _tc.packed_linear_mx8(xq, pw))
        _tc.FP4_TMA = False
        t8 = graph_time(lambda: _tc.packed_linear_mx8(xq, pw)) if _tc.mx8_weight_ok(pw) else float('nan')
        trb = graph_time(lambda: _tc.packed_linear(xb, pr))
        tc = graph_time(lambda: torch.matmul(xb, Wd.t()))
        wb = K * Nn * (0.5 + 1 / 32)
        print('K=%5d N=%5d M=%3d (%d outliers) | fp4-by-TMA %.1f us (%.0f GB/s of packed W) | producer-widen %.1f us | row-block artefact (0.94 B/w) %.1f us | cuBLAS bf16 dense %.1f us'
              % (K, Nn, Mx, nout, t4 * 1e3, wb / t4 / 1e6, t8 * 1e3, trb * 1e3, tc * 1e3)); sys.stdout.flush()
