"""compute-bound block-scaled GEMM: the TMA-fed kernel (msq_gemm_mxfp8_tma; dense E4M3 / packed FP4 weight operand, token-tile
widths) against the producer-assembled kernel (msq_gemm_mxfp8_dense) and cuBLAS bf16 on the dequantized operands"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200
from msq_b200 import _tc
dev = 'cuda'
def tm(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
shapes = [(4096, 4096, 4096), (8192, 8192, 8192), (4096, 11008, 4096), (4096, 4096, 11008), (8192, 4096, 4096)]
if len(sys.argv) > 1: shapes = shapes[:int(sys.argv[1])]
for (M, Nn, K) in shapes:
    W = torch.randn(Nn, K, device=dev) * 0.02
    x = torch.randn(M, K, device=dev)
    xq = msq_b200.pack_mx(x, 8, 'fp8_e4m3', axes=[-1], block_size=32)
    pwp = msq_b200.pack_mx(W, 8, 'fp4_e2m1', axes=[1], block_size=32)
    pwo = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
    fl = 2.0 * M * Nn * K
    xb, wb = xq.dequantize(torch.bfloat16), pwp.dequantize(torch.bfloat16)
    t = tm(lambda: torch.matmul(xb, wb.t()))
    print("M %d N %d K %d | cuBLAS bf16 %.3f ms %.0f TF" % (M, Nn, K, t, fl / t / 1e9), flush=True)
    # old kernel (producer-assembled scales), cached dense operand
    dense = _tc.widen_weight_mx8(pwo)
    pwo._mx8_dense = dense
    t = tm(lambda: _tc.packed_linear_mx8(xq, pwo))
    pwo._mx8_dense = None
    print("   producer-assembled scales (k_gemm_mx8 dense): %.3f ms %.0f TF" % (t, fl / t / 1e9), flush=True)
    for bn in (0, 256, 240, 224, 192, 128):
        if bn: os.environ["MSQ_MX8T_BN"] = str(bn)
        else: os.environ.pop("MSQ_MX8T_BN", None)
        bnr = _tc.N.lib().msq_gemm_mxfp8_tma_tile_tokens(M, Nn, K)
        td = tm(lambda: _tc.packed_linear_mx8(xq, pwo))
        tf = tm(lambda: _tc.packed_linear_mx8(xq, pwp))
        print("   TMA-fed bn %3d (%3d): E4M3 weight %.3f ms %.0f TF | FP4 codes %.3f ms %.0f TF" % (bn, bnr, td, fl / td / 1e9, tf, fl / tf / 1e9), flush=True)
    os.environ.pop("MSQ_MX8T_BN", None)
    y = _tc.packed_linear_mx8(xq, pwp)
    ref = xb.float() @ wb.float().t()
    print("   max rel err vs fp32 matmul of the dequantized operands: %.2e" % float((y - ref).abs().max() / ref.abs().max()))
