import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, msq_b200, bench
dev = torch.device("cuda", 0)
dev_w = [w.to(dev) for w in bench.make_weights(dev, 1000)]
nel = bench.layer_elems()
import json
ex = {}
# reuse the bench's pack leg only
from msq_b200 import _native as N, _ops, packed
import ctypes as C
QCFG = bench.QCFG
fi, fo = _ops.fmt4(QCFG["in_fmt"]), _ops.fmt4(QCFG["out_fmt"])
calls, pbytes = [], 0
for w in dev_w:
    pre, L, post = _ops.geometry(w.shape, 0)
    pl, wi, hi_w, stride, nseg = packed._alloc_planes(w, pre, L, post, 16, fi, fo, False)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    calls.append((w, pre, L, post, pl, st))
    p = msq_b200.pack_mx_outlier(w, 8, 8, QCFG["in_fmt"], QCFG["out_fmt"], 2, [0], 16)
    pbytes += p.nbytes
path = C.c_int(0)
def pack_all():
    for w, pre, L, post, pl, st in calls:
        rc = N.lib().msq_quantize_mx_outlier_f32(N.ptr(w), None, pre, L, post, 16, 8, 8, fi[0], fi[1], fi[2], fi[3], fo[0], fo[1], fo[2], fo[3], 2.0, 0, 0, 0, None, None,
            N.ptr(pl["codes"]), N.ptr(pl["e_in"]), N.ptr(pl["e_out"]), N.ptr(pl["bitmap"]), N.ptr(pl["hi"]), N.ptr(pl["seg_cnt"]), None, None, N.ptr(st), C.byref(path), N.stream_ptr(dev))
        assert rc == 0
ms = bench._time_ms(pack_all, 20)
bpe = pbytes / nel
print("pack: %.3f ms/layer  %.1f GB/s  frac %.3f  (%.1f Gelem/s)" % (ms, nel * (4 + bpe) / ms / 1e6, nel * (4 + bpe) / ms / 1e6 / 6543.1, nel / ms / 1e6))
