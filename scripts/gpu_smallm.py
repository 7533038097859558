"""small-M (decode) packed GEMMs: kernel-only time with the launches replayed from a CUDA graph (the Python / ctypes
call costs more than the kernel at this size), against a dense bf16 cuBLAS GEMM on the dequantized weight"""
import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import msq_b200
from msq_b200 import _tc
def graph_time(f, iters=50):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10): f()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 10
if __name__ == '__main__':
    for K, Nn in ((4096, 4096), (4096, 11008), (11008, 4096), (8192, 8192)):
        W = torch.randn(Nn, K, device='cuda') * 0.02
        pw = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 5, [1], 32, variant=1)
        Wd = pw.dequantize(torch.bfloat16)
        wbytes = pw.nbytes
        for Mx in (1, 16, 64, 256):
            x = torch.randn(Mx, K, device='cuda')
            xb = x.to(torch.bfloat16)
            xq = msq_b200.pack_mx(x, 8, 'fp8_e4m3', axes=[-1], block_size=32)
            pw._mx8_dense = None
            t8 = graph_time(lambda: _tc.packed_linear_mx8(xq, pw))
            _tc.packed_linear_mx8(xq, pw, cache_dense=True)
            t8d = graph_time(lambda: _tc.packed_linear_mx8(xq, pw))
            y_d = _tc.packed_linear_mx8(xq, pw); pw._mx8_dense = None
            y_f = _tc.packed_linear_mx8(xq, pw)
            assert float((y_d - y_f).abs().max()) <= 1e-3 * float(y_f.abs().max()) + 1e-6
            t4 = graph_time(lambda: _tc.packed_linear(xb, pw))
            tc = graph_time(lambda: torch.matmul(xb, Wd.t()))
            print('K=%5d N=%5d M=%3d | mx8 cached-dense %.1f us (%.0f GB/s of E4M3 W) | mx8 %.1f us (%.0f GB/s of packed W) | w4-bf16 %.1f us (%.0f GB/s) | cuBLAS bf16 dense %.1f us (%.0f GB/s of bf16 W)'
                  % (K, Nn, Mx, t8d * 1e3, (K * Nn * 33 / 32) / t8d / 1e6, t8 * 1e3, wbytes / t8 / 1e6, t4 * 1e3, wbytes / t4 / 1e6, tc * 1e3, 2 * K * Nn / tc / 1e6)); sys.stdout.flush()
