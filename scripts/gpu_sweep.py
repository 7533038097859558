"""throughput of every quantizer entry point (GB/s of 4 B read + 4 B written per element), B200"""
import sys, torch
sys.path.insert(0, '.')
import msq_b200
from msq_b200 import _ops
torch.manual_seed(0)
W = torch.randn(8192, 8192, device='cuda') * 0.02
n = W.numel()
def t(fn, iters=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
out = torch.empty_like(W)
st = torch.zeros(1, dtype=torch.int32, device='cuda')
rows = []
def rep(name, ms, bpe=8.0):
    print('%-58s %8.3f ms  %7.0f GB/s  (%.0f%% of 6543)' % (name, ms, n * bpe / ms / 1e6, 100 * n * bpe / ms / 1e6 / 6543)); sys.stdout.flush()
for axis, bs in ((0, 16), (0, 32), (1, 16), (1, 32)):
    rep('outlier fp4/fp8 axis%d bs%d' % (axis, bs), t(lambda: _ops.mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp8_e4m3', 2.0, axis, bs, want=('deq',), check=False, out=out, status=st)))
for axis, bs in ((0, 32), (1, 32), (1, 16)):
    for mode in (0, 1):
        for fmt in ('fp8_e4m3', 'fp4_e2m1', 'int8'):
            rep('plain %s axis%d bs%d mode%d' % (fmt, axis, bs, mode), t(lambda: _ops.mx_plain(W, 8, fmt, axis, bs, mode=mode)))
for native in (False, True):
    rep('elemwise fp8_e4m3 native=%s' % native, t(lambda: _ops.elemwise(W, 5, 4, 448.0, 'nearest', True, True, native)))
    rep('elemwise bf16 native=%s' % native, t(lambda: _ops.elemwise(W, 9, 8, 3.38e38, 'nearest', True, True, native)))
p = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp8_e4m3', 2, [0], 16)
rep('unpack fp4/fp8 axis0 bs16 -> fp32', t(lambda: p.dequantize()), 4 + p.nbytes / n)
p2 = msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 2, [1], 32)
rep('unpack fp4/fp4 axis1 bs32 -> fp32', t(lambda: p2.dequantize()), 4 + p2.nbytes / n)
rep('pack fp4/fp8 axis0 bs16 (python call incl. alloc)', t(lambda: msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp8_e4m3', 2, [0], 16, check=False)), 4 + p.nbytes / n)
rep('pack fp4/fp4 axis1 bs32 (python call incl. alloc)', t(lambda: msq_b200.pack_mx_outlier(W, 8, 8, 'fp4_e2m1', 'fp4_e2m1', 2, [1], 32, check=False)), 4 + p2.nbytes / n)
