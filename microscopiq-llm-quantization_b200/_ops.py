"""Thin tensor-level wrappers over the C ABI (shape bookkeeping, output allocation,
status -> exception).  Everything numerical happens in libmsq_b200.so."""
import ctypes as C

import torch

from . import _native as N
from .mx.formats import ElemFormat, RoundingMode, _get_format_params

#: set False to skip the 4-byte status read-back (one host sync per call)
CHECK_STATUS = True

_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}

_i, _ll, _f, _p = C.c_int, C.c_longlong, C.c_float, C.c_void_p
N.register("msq_host_quantize_mx_outlier_f32", _i, [_p, _p, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _f, _i, _i, _i, _f,
                                                     _f, _i, _i, _ll, _i])
N.register("msq_quantize_mx_outlier_multi_f32", _i, [_i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _i, _i, _f, _f,
                                                      _i, _i, _p, _p])
N.register("msq_host_sync", _i, [C.POINTER(C.c_int)])
N.register("msq_host_release", _i, [])
N.register("msq_magic_div", _i, [_ll, C.POINTER(C.c_uint), C.POINTER(C.c_int)])


def _rmode(round):
    if isinstance(round, RoundingMode):
        return int(round)
    if round not in N.ROUND:
        raise Exception("Unrecognized round method %s" % (round,))      # elemwise_ops.py:74
    return N.ROUND[round]


def _single_axis(axes, ndim):
    if axes is None:
        raise Exception("axes required in order to determine which dimension to apply block size to")
    axes = [axes] if isinstance(axes, int) else list(axes)
    axes = [a + ndim if a < 0 else a for a in axes]
    if len(axes) != 1:
        raise NotImplementedError("the B200 kernels block along exactly one axis (got axes=%s)" % (axes,))
    return axes[0]


def geometry(shape, axis):
    pre = 1
    for s in shape[:axis]:
        pre *= s
    post = 1
    for s in shape[axis + 1:]:
        post *= s
    return pre, shape[axis], post


def _as_f32(A):
    N.require_cuda(A, "input")
    return A.detach().contiguous().to(torch.float32)


def fmt4(fmt):
    ebits, mbits, emax, max_norm, _ = _get_format_params(fmt)
    return ebits, mbits, emax, float(max_norm)


def elemwise(A, bits, exp_bits, max_norm, round="nearest", saturate_normals=False, allow_denorm=True, native=False):
    """msq_quantize_elemwise on a CUDA tensor of dtype fp32/fp16/bf16 (same dtype out)."""
    N.require_cuda(A, "input")
    if A.dtype not in _DT:
        raise TypeError("unsupported dtype %s" % A.dtype)
    x = A.detach().contiguous()
    out = torch.empty_like(x)
    rc = N.lib().msq_quantize_elemwise(N.ptr(x), N.ptr(out), x.numel(), _DT[x.dtype], int(bits), int(exp_bits),
                                       float(max_norm), _rmode(round), int(bool(saturate_normals)),
                                       int(bool(allow_denorm)), int(bool(native)), N.stream_ptr(x.device))
    N.check(rc, "msq_quantize_elemwise")
    return out


def mx_plain(A, scale_bits, elem_format, axis, block_size, round="nearest", flush=False, mode=0,
             want_exp=False, ext_max=None):
    """plain MX fake-quant (+ optional shared exponents) of a CUDA tensor along `axis`."""
    x = _as_f32(A)
    pre, L, post = geometry(x.shape, axis)
    bs = int(block_size) if block_size and block_size > 0 else L
    nblk = (L + bs - 1) // bs
    ebits, mbits, emax, max_norm = fmt4(elem_format)
    out = torch.empty_like(x)
    e = torch.empty(pre * nblk * post, dtype=torch.int8, device=x.device) if want_exp else None
    st = torch.zeros(1, dtype=torch.int32, device=x.device)
    path = C.c_int(0)
    rc = N.lib().msq_quantize_mx_f32(N.ptr(x), N.ptr(out), pre, L, post, bs, int(scale_bits), ebits, mbits, emax,
                                     max_norm, int(bool(flush)), _rmode(round), int(mode), N.ptr(ext_max), None,
                                     N.ptr(e), N.ptr(st), C.byref(path), N.stream_ptr(x.device))
    N.check(rc, "msq_quantize_mx_f32")
    out = out.to(A.dtype)
    return (out, e, path.value) if want_exp else out


class OutlierResult(object):
    __slots__ = ("deq", "e_in", "e_out", "blk_cnt", "mask", "status", "path", "geom")


def across_block_bounds(x, axis, bs, std_dev):
    pre, L, post = geometry(x.shape, axis)
    nblk = (L + bs - 1) // bs
    if nblk < 2:
        raise AssertionError("std contains NaN values")                  # mx_ops.py:66
    lo = torch.empty(pre * bs * post, dtype=torch.float32, device=x.device)
    hi = torch.empty_like(lo)
    rc = N.lib().msq_outlier_bounds_across_blocks_f32(N.ptr(x), N.ptr(lo), N.ptr(hi), pre, L, post, bs,
                                                      float(std_dev), N.stream_ptr(x.device))
    N.check(rc, "msq_outlier_bounds_across_blocks_f32")
    return lo, hi


def mx_outlier(A, in_sb, out_sb, in_fmt, out_fmt, std_dev, axis, block_size, round="nearest", flush=False,
               variant=0, want=("deq",), check=None, out=None, status=None):
    """MX+outlier fake-quant.  want: subset of {deq, e_in, e_out, blk_cnt, mask}.
    out / status: optional preallocated fp32 output and int32[1] status word (no allocation,
    and with check=False no host synchronisation)."""
    x = _as_f32(A)
    pre, L, post = geometry(x.shape, axis)
    bs = int(block_size) if block_size and block_size > 0 else L
    nblk = (L + bs - 1) // bs
    nb = pre * nblk * post
    fi, fo = fmt4(in_fmt), fmt4(out_fmt)
    dev = x.device
    r = OutlierResult()
    r.geom = (pre, L, post, bs, nblk)
    r.deq = (out if out is not None else torch.empty_like(x)) if "deq" in want else None
    r.e_in = torch.empty(nb, dtype=torch.int8, device=dev) if "e_in" in want else None
    r.e_out = torch.empty(nb, dtype=torch.int8, device=dev) if "e_out" in want else None
    r.blk_cnt = torch.empty(nb, dtype=torch.uint8, device=dev) if "blk_cnt" in want else None
    r.mask = torch.empty(x.shape, dtype=torch.uint8, device=dev) if "mask" in want else None
    st = status if status is not None else torch.zeros(1, dtype=torch.int32, device=dev)
    lo = hi = None
    if variant == 1:
        lo, hi = across_block_bounds(x, axis, bs, std_dev)
    path = C.c_int(0)
    rc = N.lib().msq_quantize_mx_outlier_f32(
        N.ptr(x), N.ptr(r.deq), pre, L, post, bs, int(in_sb), int(out_sb),
        fi[0], fi[1], fi[2], fi[3], fo[0], fo[1], fo[2], fo[3],
        float(std_dev), int(bool(flush)), _rmode(round), int(variant), N.ptr(lo), N.ptr(hi),
        None, N.ptr(r.e_in), N.ptr(r.e_out), None, None, None, N.ptr(r.blk_cnt), N.ptr(r.mask),
        N.ptr(st), C.byref(path), N.stream_ptr(dev))
    N.check(rc, "msq_quantize_mx_outlier_f32")
    r.path = path.value
    r.status = st
    if CHECK_STATUS if check is None else check:
        raise_on_status(int(st.item()))
    if r.deq is not None:
        r.deq = r.deq.to(A.dtype)
    return r


def mx_outlier_host(A, in_sb, out_sb, in_fmt, out_fmt, std_dev, axis, block_size, round="nearest", flush=False,
                    out=None, sync=True, chunk_bytes=0):
    """MX+outlier fake-quant of a HOST tensor (utils/quant.py:147-266 called on a CPU-resident weight): chunked
    upload / kernel / download pipeline inside libmsq_b200.so (csrc/msq_host.cu).  Pinned `A` / `out` run at PCIe
    speed; with sync=False the call only queues work -- call host_sync() before reading `out`."""
    if A.is_cuda:
        raise ValueError("mx_outlier_host takes a host tensor (use mx_outlier for CUDA tensors)")
    if not torch.cuda.is_available():
        raise RuntimeError("msq_b200 needs a CUDA device: there is no CPU implementation")
    x = A.detach().contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    pre, L, post = geometry(x.shape, axis)
    bs = int(block_size) if block_size and block_size > 0 else L
    fi, fo = fmt4(in_fmt), fmt4(out_fmt)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    elif out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or out.shape != x.shape:
        raise ValueError("out must be a contiguous fp32 host tensor of the input's shape")
    rc = N.lib().msq_host_quantize_mx_outlier_f32(
        x.data_ptr(), out.data_ptr(), pre, L, post, bs, int(in_sb), int(out_sb),
        fi[0], fi[1], fi[2], fi[3], fo[0], fo[1], fo[2], fo[3],
        float(std_dev), int(bool(flush)), _rmode(round), int(chunk_bytes), int(bool(sync)))
    N.check(rc, "msq_host_quantize_mx_outlier_f32")
    if sync:
        host_sync()
    else:
        _HOST_KEEPALIVE.append((x, out))      # the queued copies read / write these buffers
    return out if A.dtype == torch.float32 or not sync else out.to(A.dtype)


_HOST_KEEPALIVE = []


def host_sync(check=None):
    """wait for the queued host-pipeline work of the current device; raises like the reference on bad blocks"""
    st = C.c_int(0)
    N.check(N.lib().msq_host_sync(C.byref(st)), "msq_host_sync")
    del _HOST_KEEPALIVE[:]
    if CHECK_STATUS if check is None else check:
        raise_on_status(st.value)
    return st.value


def raise_on_status(s):
    """reproduce the reference's assertion behaviour (utils/quant.py:103-128, :479)"""
    if s & N.ST_NONFINITE_INPUT:
        raise AssertionError("std contains NaN values")
    if s & N.ST_EXP_IN_OVERFLOW:
        raise AssertionError("inlier_val contains NaN values")
    if s & N.ST_EXP_OUT_OVERFLOW:
        raise AssertionError("shared_exp_out contains NaN values")


class MultiQuantPlan(object):
    """argument block of msq_quantize_mx_outlier_multi_f32 for a fixed list of weight matrices (built once, launched
    many times: no per-call host work beyond the ctypes call)"""

    def __init__(self, tensors, outs, in_sb, out_sb, in_fmt, out_fmt, std_dev, block_size, round="nearest", flush=False,
                 status=None):
        n = len(tensors)
        assert 0 < n <= 16 and len(outs) == n
        for t, o in zip(tensors, outs):
            N.require_cuda(t, "input")
            assert t.dim() == 2 and t.dtype == torch.float32 and t.is_contiguous()
            assert o.shape == t.shape and o.dtype == torch.float32 and o.is_contiguous() and o.device == t.device
        self.keep = (list(tensors), list(outs))
        self.n = n
        self.inp = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
        self.out = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        self.rows = (C.c_longlong * n)(*[t.shape[0] for t in tensors])
        self.cols = (C.c_longlong * n)(*[t.shape[1] for t in tensors])
        self.fi, self.fo = fmt4(in_fmt), fmt4(out_fmt)
        self.args = (int(block_size), int(in_sb), int(out_sb))
        self.tail = (float(std_dev), int(bool(flush)), _rmode(round))
        self.status = status if status is not None else torch.zeros(1, dtype=torch.int32, device=tensors[0].device)
        self.device = tensors[0].device

    def launch(self):
        fi, fo = self.fi, self.fo
        rc = N.lib().msq_quantize_mx_outlier_multi_f32(self.n, self.inp, self.out, self.rows, self.cols, *self.args,
                                                       fi[0], fi[1], fi[2], fi[3], fo[0], fo[1], fo[2], fo[3], *self.tail,
                                                       N.ptr(self.status), N.stream_ptr(self.device))
        N.check(rc, "msq_quantize_mx_outlier_multi_f32")


def mx_outlier_multi(tensors, in_sb, out_sb, in_fmt, out_fmt, std_dev, block_size, round="nearest", flush=False,
                     outs=None, check=None):
    """quantize_mx_outlier_v1(axes=[0]) of several 2-D fp32 CUDA weights in one launch; returns the list of outputs"""
    xs = [_as_f32(t) for t in tensors]
    outs = outs if outs is not None else [torch.empty_like(x) for x in xs]
    plan = MultiQuantPlan(xs, outs, in_sb, out_sb, in_fmt, out_fmt, std_dev, block_size, round, flush)
    plan.launch()
    if CHECK_STATUS if check is None else check:
        raise_on_status(int(plan.status.item()))
    return outs
