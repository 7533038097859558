"""the quantizers' own artefact (row blocks of 16, FP4 inliers / FP8 outliers): row-block kernel vs its exact MX-FP8 form
on the block-scaled kernels, bf16 activations"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
dev = 'cuda'
def tm(fn):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 100 * 1e3
for KN in (4096, 8192):
    W = torch.randn(KN, KN, device=dev) * 0.02
    pw = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp8_e4m3', 2, [0], 16)
    wb = pw.dequantize(torch.bfloat16)
    print("KN %d: artefact %.3f B/weight, outlier fraction %.3f, exact MX-FP8 form: %s" % (
        KN, sum(t.numel() * t.element_size() for t in (pw.codes, pw.bitmap, pw.e_in, pw.e_out, pw.hi)) / W.numel(),
        float(pw.outlier_mask().float().mean()), _tc.rowblock_mxfp8_operand(pw) is not None), flush=True)
    for M in (1, 16, 64, 256):
        x = torch.randn(M, KN, device=dev).to(torch.bfloat16)
        t_c = tm(lambda: torch.matmul(x, wb.t()))
        _tc.ROWBLOCK_VIA_MXFP8 = False
        t_rb = tm(lambda: _tc.packed_linear(x, pw))
        _tc.ROWBLOCK_VIA_MXFP8 = True
        t_new = tm(lambda: _tc.packed_linear(x, pw))
        y = _tc.packed_linear(x, pw); ref = x.float() @ pw.dequantize().t()
        print("   M %3d: cuBLAS bf16 %.1f us | k_gemm_rb %.1f | MX-FP8 form, block-scaled %.1f | rel err %.1e" % (
            M, t_c, t_rb, t_new, float((y - ref).abs().max() / ref.abs().max())), flush=True)
