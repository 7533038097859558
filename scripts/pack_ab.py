"""A/B of the TMA-staged quantize+pack kernel (k_pack_tma, path 5) against k_strided<..,PACK> (MSQ_QUANT_TMA=0):
every plane bit-identical, then GB/s on bench.py's layer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msq_b200
from msq_b200 import _ops, packed

dev = torch.device("cuda:0")
torch.manual_seed(0)
ok = True


def pack(W, fi, fo, tma):
    os.environ["MSQ_QUANT_TMA"] = "1" if tma else "0"
    return msq_b200.pack_mx_outlier(W, 8, 8, fi, fo, 2, [0], 16)


for fi, fo in (("fp4_e2m1", "fp8_e4m3"), ("fp4_e2m1", "fp4_e2m1"), ("fp4_e2m1", "fp6_e3m2"), ("int4", "fp8_e4m3")):
    for shp in ((16, 256), (64, 1024), (528, 768), (100, 256), (4096, 4096), (1000, 11008), (4096, 11008)):
        W = torch.randn(shp, device=dev) * 0.02
        W.view(-1)[::97] *= 30
        a = pack(W, fi, fo, True)
        b = pack(W, fi, fo, False)
        if a.path != 5:
            print("path", a.path, fi, fo, shp)
            continue
        assert b.path == 1
        for name in ("codes", "bitmap", "e_in", "e_out", "seg_cnt"):
            x, y = getattr(a, name), getattr(b, name)
            if not torch.equal(x, y):
                ok = False
                print("MISMATCH", name, fi, fo, shp, int((x != y).sum()))
        if a.hi_w:
            # compare only the used prefix of every segment
            st = a.hi_seg_stride
            for sgi in range(a.nseg):
                nb = (int(a.seg_cnt[sgi]) * a.hi_w + 7) // 8
                if not torch.equal(a.hi[sgi * st: sgi * st + nb], b.hi[sgi * st: sgi * st + nb]):
                    ok = False
                    print("MISMATCH hi seg", sgi, fi, fo, shp)
                    break
            if a.nseg > 64:
                pass
        if not torch.equal(a.dequantize(), b.dequantize()):
            ok = False
            print("MISMATCH dequantize", fi, fo, shp)
print("bit-identical:", ok)

import ctypes as C
from msq_b200 import _native as N
shp = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
Ws = [torch.randn(s, device=dev) * 0.02 for s in shp]
nel = sum(w.numel() for w in Ws)
fi, fo = _ops.fmt4("fp4_e2m1"), _ops.fmt4("fp8_e4m3")
calls, pbytes = [], 0
for w in Ws:
    pre, L, post = _ops.geometry(w.shape, 0)
    pl, wi, hi_w, stride, nseg = packed._alloc_planes(w, pre, L, post, 16, fi, fo, False)
    calls.append((w, pre, L, post, pl, torch.zeros(1, dtype=torch.int32, device=dev)))
    p = msq_b200.pack_mx_outlier(w, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16)
    pbytes += p.nbytes
path = C.c_int(0)


def pack_all():
    for w, pre, L, post, pl, st in calls:
        rc = N.lib().msq_quantize_mx_outlier_f32(
            N.ptr(w), None, pre, L, post, 16, 8, 8, fi[0], fi[1], fi[2], fi[3], fo[0], fo[1], fo[2], fo[3], 2.0, 0, 0, 0,
            None, None, N.ptr(pl["codes"]), N.ptr(pl["e_in"]), N.ptr(pl["e_out"]), N.ptr(pl["bitmap"]), N.ptr(pl["hi"]),
            N.ptr(pl["seg_cnt"]), None, None, N.ptr(st), C.byref(path), N.stream_ptr(dev))
        assert rc == 0


bpe = pbytes / nel
for tma in (0, 1, 0, 1):
    os.environ["MSQ_QUANT_TMA"] = str(tma)
    for _ in range(3):
        pack_all()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        pack_all()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gbps = nel * (4 + bpe) / ms / 1e6
    print("pack tma=%d path %d: %.4f ms  %.0f GB/s  frac %.3f  (%.3f B/elem packed)" % (tma, path.value, ms, gbps, gbps / 6543.1, bpe))
sys.exit(0 if ok else 1)
