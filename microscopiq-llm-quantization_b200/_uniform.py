"""GPTQ uniform quantizer on B200 (utils/quant.py:268-390) over the C ABI."""
import ctypes as C

import torch

from . import _native as N

_p, _ll, _i, _f = C.c_void_p, C.c_longlong, C.c_int, C.c_float
N.register("msq_uniform_quantize_f32", _i, [_p, _p, _p, _p, _i, _ll, _ll, _ll, _p])
N.register("msq_uniform_find_params_f32", _i, [_p, _p, _p, _ll, _ll, _i, _i, _i, _f, _i, _f, _p])


def _channel_geometry(x_shape, p_shape):
    """x [pre, C, post] with a per-C parameter: find the single non-unit dim of the
    (right-aligned, broadcastable) parameter shape."""
    nd = len(x_shape)
    ps = [1] * (nd - len(p_shape)) + list(p_shape)
    dims = [d for d in range(nd) if ps[d] != 1]
    if not dims:
        return 1, 1, 1
    if len(dims) > 1 or ps[dims[0]] != x_shape[dims[0]]:
        return None
    d = dims[0]
    pre = 1
    for s in x_shape[:d]:
        pre *= s
    post = 1
    for s in x_shape[d + 1:]:
        post *= s
    return pre, x_shape[d], post


def quantize(x, scale, zero, maxq):
    N.require_cuda(x, "x")
    maxq = int(maxq)
    xf = x.detach().contiguous().to(torch.float32)
    scale = torch.as_tensor(scale, dtype=torch.float32, device=x.device)
    zero = torch.as_tensor(zero, dtype=torch.float32, device=x.device)
    out_shape = torch.broadcast_shapes(xf.shape, scale.shape, zero.shape)
    if tuple(out_shape) != tuple(xf.shape):
        xf = xf.expand(out_shape).contiguous()
    g = _channel_geometry(xf.shape, scale.shape) if scale.shape == zero.shape else None
    if g is None:           # arbitrary broadcast: materialise per-element parameters
        s = scale.expand(out_shape).contiguous().view(-1)
        z = zero.expand(out_shape).contiguous().view(-1)
        pre, cn, post = 1, xf.numel(), 1
    else:
        pre, cn, post = g
        if cn == 1 and scale.numel() == 1:
            pre, cn, post = 1, 1, xf.numel()
        s = scale.contiguous().view(-1)
        z = zero.contiguous().view(-1)
    out = torch.empty_like(xf)
    rc = N.lib().msq_uniform_quantize_f32(N.ptr(xf), N.ptr(out), N.ptr(s), N.ptr(z), maxq, pre, cn, post,
                                          N.stream_ptr(x.device))
    N.check(rc, "msq_uniform_quantize_f32")
    return out


def find_params(x, maxq, perchannel, sym, mse, norm, grid, maxshrink, weight):
    """returns (scale, zero) shaped like Quantizer.find_params leaves them (quant.py:358-379)"""
    N.require_cuda(x, "x")
    shape = x.shape
    if perchannel:
        if weight:
            x = x.flatten(1)
        else:
            if len(shape) == 4:
                x = x.permute([1, 0, 2, 3]).flatten(1)
            if len(shape) == 3:
                x = x.reshape((-1, shape[-1])).t()
            if len(shape) == 2:
                x = x.t()
    else:
        x = x.flatten().unsqueeze(0)
    xf = x.detach().contiguous().to(torch.float32)
    rows, n = xf.shape
    scale = torch.empty(rows, dtype=torch.float32, device=xf.device)
    zero = torch.empty(rows, dtype=torch.float32, device=xf.device)
    rc = N.lib().msq_uniform_find_params_f32(N.ptr(xf), N.ptr(scale), N.ptr(zero), rows, n, int(maxq), int(bool(sym)),
                                             int(bool(mse)), float(norm), int(grid), float(maxshrink),
                                             N.stream_ptr(xf.device))
    N.check(rc, "msq_uniform_find_params_f32")
    if not perchannel:
        rep = shape[0] if weight else (shape[1] if len(shape) != 3 else shape[2])
        scale, zero = scale.repeat(rep), zero.repeat(rep)
    if weight:
        shp = [-1] + [1] * (len(shape) - 1)
        return scale.reshape(shp), zero.reshape(shp)
    if len(shape) == 4:
        return scale.reshape((1, -1, 1, 1)), zero.reshape((1, -1, 1, 1))
    if len(shape) == 3:
        return scale.reshape((1, 1, -1)), zero.reshape((1, 1, -1))
    if len(shape) == 2:
        return scale.unsqueeze(0), zero.unsqueeze(0)
    return scale, zero
