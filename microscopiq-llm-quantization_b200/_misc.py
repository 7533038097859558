"""Small helpers of the reference API that are not on the throughput path."""
import ctypes as C

import torch

from . import _native as N
from . import _ops

_p, _ll, _i = C.c_void_p, C.c_longlong, C.c_int
N.register("msq_shared_exponents_f32", _i, [_p, _p, _ll, _ll, _ll, _i, _p])


def shared_exponents(A, method="max", axes=None, ebits=0):
    """utils/quant.py:498-541 (exact floor(log2))"""
    N.require_cuda(A, "A")
    x = A.detach().contiguous().to(torch.float32)
    if method == "none":
        flat = x.reshape(-1, 1)
        out = torch.empty(flat.shape[0], dtype=torch.float32, device=x.device)
        rc = N.lib().msq_shared_exponents_f32(N.ptr(flat), N.ptr(out), flat.shape[0], 1, 1, int(ebits),
                                              N.stream_ptr(x.device))
        N.check(rc, "msq_shared_exponents_f32")
        return out.reshape(A.shape).to(A.dtype)
    if method != "max":
        raise Exception("Unrecognized shared exponent selection method %s" % (method))
    if axes is None:
        pre, L, post, shape = 1, x.numel(), 1, []
    else:
        axis = _ops._single_axis(axes, x.ndim)
        pre, L, post = _ops.geometry(x.shape, axis)
        shape = list(x.shape)
        shape[axis] = 1
    out = torch.empty(pre * post, dtype=torch.float32, device=x.device)
    rc = N.lib().msq_shared_exponents_f32(N.ptr(x), N.ptr(out), pre, L, post, int(ebits), N.stream_ptr(x.device))
    N.check(rc, "msq_shared_exponents_f32")
    return out.reshape(shape).to(A.dtype)
