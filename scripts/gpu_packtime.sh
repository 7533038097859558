cd $GRAFT_REPO_ROOT
timeout 300 python - <<'PY'
import torch, sys, ctypes as C, time
sys.path.insert(0,'.')
import msq_b200
from msq_b200 import _native as N, _ops, packed
W = (torch.randn(11008, 4096, device='cuda')*0.02)
n = W.numel()
def direct(fi, fo, axis, bs, variant=0, iters=20):
    x = W
    pre, L, post = _ops.geometry(x.shape, axis)
    f_i, f_o = _ops.fmt4(fi), _ops.fmt4(fo)
    pl, wi, hi_w, stride, nseg = packed._alloc_planes(x, pre, L, post, bs, f_i, f_o, False)
    st = torch.zeros(1, dtype=torch.int32, device='cuda'); path = C.c_int(0)
    lo = hi = None
    if variant == 1: lo, hi = _ops.across_block_bounds(x, axis, bs, 5.0)
    def call():
        return N.lib().msq_quantize_mx_outlier_f32(N.ptr(x), None, pre, L, post, bs, 8, 8, f_i[0], f_i[1], f_i[2], f_i[3], f_o[0], f_o[1], f_o[2], f_o[3], 2.0 if variant==0 else 5.0, 0, 0, variant, N.ptr(lo), N.ptr(hi), N.ptr(pl['codes']), N.ptr(pl['e_in']), N.ptr(pl['e_out']), N.ptr(pl['bitmap']), N.ptr(pl['hi']), N.ptr(pl['seg_cnt']), None, None, N.ptr(st), C.byref(path), N.stream_ptr(x.device))
    for _ in range(3): assert call() == 0
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/iters
    p = msq_b200.pack_mx_outlier(W, 8,8,fi,fo,2 if variant==0 else 5,[axis],bs, variant=variant)
    print(f'pack kernel {fi}/{fo} axis{axis} bs{bs} v{variant}: {ms*1e3:.1f} us  {(4*n + p.nbytes)/ms/1e6:.0f} GB/s  ({p.nbytes/n:.3f} B/elem)')
direct('fp4_e2m1','fp8_e4m3',0,16)
direct('int2','fp4',0,16)
direct('fp4_e2m1','fp4_e2m1',0,16)
direct('fp4_e2m1','fp4_e2m1',1,32, variant=1)
direct('fp4_e2m1','fp8_e4m3',1,32)
t0=time.time()
for _ in range(20): p = msq_b200.pack_mx_outlier(W, 8,8,'fp4_e2m1','fp8_e4m3',2,[0],16, check=False)
print('python-side per call (no sync):', (time.time()-t0)/20*1e3, 'ms')
PY
