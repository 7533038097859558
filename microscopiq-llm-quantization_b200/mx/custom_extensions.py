"""`custom_extensions.funcs` -- the seven names of the reference's pybind module
(number_system/mx/custom_extensions.py:8-19, cpp/funcs.cpp:218-226), bound to libmsq_b200.so.

    from mx import custom_extensions as ce
    ce.funcs.quantize_mx_by_tile_func_cuda(A, scale_bits, ebits, mbits, max_norm, tile, axis, flush, rmode)

Same positional signatures, defaults and argument checks (contiguous CUDA fp32 input, `axis < A.dim()`,
`bits <= 24`: funcs.h:11-13, funcs.cpp:146,169,193) and the native arithmetic of cpp/quantize.cuh / cpp/mx.cuh
(the integer-mantissa formulation, what `custom_cuda=True` selects).  Nothing is JIT-compiled: the kernels are
the ones every other entry point of this package uses.  The two `*_func_cpp` names are the reference's CPU
loops; here a host tensor is moved through the GPU and back (there is no CPU implementation in this package).
"""
import ctypes as C
import math

import torch

from .. import _native as N

_RM = {0: 0, 1: 1, 2: 2}          # RoundingMode{nearest=0, floor=1, even=2}: formats.py:15-18 <-> common.cuh:130-134
_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}

N.register("msq_reduce_inner_dim_f32", C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p])


def _check_input(A, what="A"):
    if not isinstance(A, torch.Tensor) or not A.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % what)            # CHECK_CUDA, funcs.h:11
    if not A.is_contiguous():
        raise RuntimeError("%s must be contiguous" % what)               # CHECK_CONTIGUOUS, funcs.h:12


def _geometry(A, axis):
    sizes = list(A.shape)
    pre = 1
    for s in sizes[:axis]:
        pre *= s
    post = 1
    for s in sizes[axis + 1:]:
        post *= s
    return pre, sizes[axis], post


def _emax(max_norm):
    return int(math.floor(math.log2(max_norm))) if max_norm > 0 else 0    # get_unbiased_exponent(elem_max_norm)


def _mx(A, scale_bits, ebits, mbits, max_norm, block, axis, flush, rmode, ext_max):
    if A.dtype == torch.float16:
        raise RuntimeError(" fp16 not supported for MX")                  # mx.cu:53
    if A.dtype != torch.float32:
        raise RuntimeError("expected a float32 tensor")
    if rmode not in _RM:
        raise RuntimeError("unknown rounding mode %r" % (rmode,))
    out = torch.empty_like(A)
    if A.numel() == 0:
        return out
    pre, L, post = _geometry(A, axis)
    st = torch.zeros(1, dtype=torch.int32, device=A.device)
    path = C.c_int(0)
    rc = N.lib().msq_quantize_mx_f32(N.ptr(A), N.ptr(out), pre, L, post, int(block), int(scale_bits), int(ebits),
                                     int(mbits), _emax(max_norm), float(max_norm), int(bool(flush)), _RM[rmode], 1,
                                     N.ptr(ext_max), None, None, N.ptr(st), C.byref(path), N.stream_ptr(A.device))
    N.check(rc, "msq_quantize_mx_f32")
    return out


def quantize_mx_by_tile_func_cuda(A, scale_bits, elem_ebits, elem_mbits, elem_max_norm, tile_size, axis,
                                  flush_fp32_subnorms=False, rmode=0):
    """funcs.cpp:157-176 -> mx.cu:77-172: MX quantization with blocks of `tile_size` along `axis` (zero padded)"""
    _check_input(A)
    if not axis < A.dim():
        raise RuntimeError(" quantize_mx_by_tile axis exceeds input dimensions")
    return _mx(A, scale_bits, elem_ebits, elem_mbits, elem_max_norm, tile_size, axis, flush_fp32_subnorms, rmode, None)


def quantize_mx_func_cuda(A, scale_bits, elem_ebits, elem_mbits, elem_max_norm, max_values, axis,
                          flush_fp32_subnorms=False, rmode=0):
    """funcs.cpp:135-155 -> mx.cu:13-72: the whole of `axis` shares one exponent, taken from the caller's
    `max_values` (shape of A with `axis` reduced to 1)"""
    _check_input(A)
    if not axis < A.dim():
        raise RuntimeError(" quantize_mx_func_cuda axis exceeds input dimensions")
    _check_input(max_values, "max_values")
    pre, L, post = _geometry(A, axis)
    if max_values.numel() != pre * post or max_values.dtype != torch.float32:
        raise RuntimeError("max_values must hold one fp32 value per shared-exponent group")
    return _mx(A, scale_bits, elem_ebits, elem_mbits, elem_max_norm, max(L, 1), axis, flush_fp32_subnorms, rmode,
               max_values)


def quantize_elemwise_func_cuda(A, bits, exp_bits, max_norm, rmode=0, saturate_normals=False, allow_denorm=True):
    """funcs.cpp:178-195 -> elemwise.cu: fp32 or fp16 tensor, rounding of cpp/quantize.cuh:88-149"""
    _check_input(A)
    if not bits <= 24:
        raise RuntimeError(" quantize_elemwise with bits > 24 leads to negative shifts")
    if A.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("expected a float32 or float16 tensor")
    out = torch.empty_like(A)
    rc = N.lib().msq_quantize_elemwise(N.ptr(A), N.ptr(out), A.numel(), _DT[A.dtype], int(bits), int(exp_bits),
                                       float(max_norm), _RM[rmode], int(bool(saturate_normals)), int(bool(allow_denorm)),
                                       1, N.stream_ptr(A.device))
    N.check(rc, "msq_quantize_elemwise")
    return out


def _reduce(A, op):
    _check_input(A)
    if A.dtype == torch.float16:
        raise RuntimeError("fp16 reduce is not supported")                # reduce.cu:52
    if A.dtype != torch.float32 or A.dim() < 1:
        raise RuntimeError("expected a float32 tensor")
    inner = A.shape[-1]
    if inner < 32 or inner % 32 or inner > 1024:
        raise RuntimeError("inner dimension must be a multiple of 32 in [32, 1024]")      # reduce.cu:34-35
    out = torch.zeros(A.shape[:-1], dtype=A.dtype, device=A.device)
    rc = N.lib().msq_reduce_inner_dim_f32(N.ptr(A), N.ptr(out), out.numel(), inner, op, N.stream_ptr(A.device))
    N.check(rc, "msq_reduce_inner_dim_f32")
    return out


def reduce_sum_inner_dim(A):
    """funcs.cpp:197-203: sum over the innermost dimension, two-level shuffle tree in the reference's order"""
    return _reduce(A, 0)


def reduce_max_inner_dim(A):
    """funcs.cpp:205-211 (start value -65504, cpp/reduce.cuh:87)"""
    return _reduce(A, 1)


def _through_gpu(fn, A, *args):
    if A.is_cuda:
        raise RuntimeError("the *_func_cpp entry points take host tensors")
    if not torch.cuda.is_available():
        raise RuntimeError("msq_b200 needs a CUDA device: there is no CPU implementation")
    return fn(A.contiguous().cuda(), *args).cpu()


def quantize_mx_func_cpp(A, scale_bits, elem_ebits, elem_mbits, elem_max_norm, max_values, axis,
                         flush_fp32_subnorms, rmode=0):
    """funcs.cpp:26-96 (a CPU loop in the reference): host tensor in, host tensor out, computed on the GPU"""
    if not 1 <= A.dim() <= 5:
        raise RuntimeError(" number of dims outside of range [1,5]")
    if not 0 <= axis < A.dim():
        raise RuntimeError(" shared axis < 0 or shared axis >= ndim")
    return _through_gpu(lambda a: quantize_mx_func_cuda(a, scale_bits, elem_ebits, elem_mbits, elem_max_norm,
                                                        max_values.contiguous().cuda(), axis, flush_fp32_subnorms, rmode), A)


def quantize_elemwise_func_cpp(A, bits, exp_bits, max_norm, rmode=0, saturate_normals=False, allow_denorm=True):
    """funcs.cpp:98-130 (a CPU loop in the reference): host tensor in, host tensor out, computed on the GPU"""
    return _through_gpu(quantize_elemwise_func_cuda, A, bits, exp_bits, max_norm, rmode, saturate_normals, allow_denorm)


class _Funcs(object):
    """attribute namespace standing in for the JIT-built pybind module `funcs`"""
    quantize_mx_func_cpp = staticmethod(quantize_mx_func_cpp)
    quantize_elemwise_func_cpp = staticmethod(quantize_elemwise_func_cpp)
    quantize_mx_func_cuda = staticmethod(quantize_mx_func_cuda)
    quantize_mx_by_tile_func_cuda = staticmethod(quantize_mx_by_tile_func_cuda)
    quantize_elemwise_func_cuda = staticmethod(quantize_elemwise_func_cuda)
    reduce_sum_inner_dim = staticmethod(reduce_sum_inner_dim)
    reduce_max_inner_dim = staticmethod(reduce_max_inner_dim)


funcs = _Funcs()
