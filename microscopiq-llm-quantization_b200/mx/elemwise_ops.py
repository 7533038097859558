"""Element-wise (bfloat / fpX / custom) rounding on B200.

Mirror of number_system/mx/elemwise_ops.py: same function names, argument order and
defaults; the arithmetic runs in libmsq_b200.so (msq_quantize_elemwise).  Tensors must
live on a CUDA device.  `custom_cuda=True` selects the integer-mantissa formulation of
the reference's native kernels (cpp/quantize.cuh:88-149), `False` the float formulation
of the reference's torch path -- the two differ only on a measure-zero set.
"""
from .. import _ops
from .formats import RoundingMode, _get_format_params, _get_max_norm, _get_min_norm  # noqa: F401


def _quantize_elemwise_core(A, bits, exp_bits, max_norm, round="nearest", saturate_normals=False,
                            allow_denorm=True, custom_cuda=False):
    if A.is_sparse:
        raise NotImplementedError("sparse tensors are not supported by the B200 kernels")
    if round == "dither":
        raise NotImplementedError("dither rounding is random and not part of the B200 contract")
    return _ops.elemwise(A, bits, exp_bits, max_norm, round, saturate_normals, allow_denorm,
                         native=bool(custom_cuda) and round in RoundingMode.string_enums())


def _quantize_elemwise(A, elem_format, round="nearest", custom_cuda=False, saturate_normals=False,
                       allow_denorm=True):
    if elem_format is None:
        return A
    ebits, mbits, _, max_norm, _ = _get_format_params(elem_format)
    return _quantize_elemwise_core(A, mbits, ebits, max_norm, round=round, allow_denorm=allow_denorm,
                                   saturate_normals=saturate_normals, custom_cuda=custom_cuda)


def _quantize_bfloat(A, bfloat, round="nearest", custom_cuda=False, allow_denorm=True):
    if bfloat == 0 or bfloat == 32:
        return A
    return _quantize_elemwise_core(A, bits=bfloat - 7, exp_bits=8, max_norm=_get_max_norm(8, bfloat - 7),
                                   round=round, allow_denorm=allow_denorm, custom_cuda=custom_cuda)


def _quantize_fp(A, exp_bits=None, mantissa_bits=None, round="nearest", custom_cuda=False, allow_denorm=True):
    if exp_bits is None or mantissa_bits is None:
        return A
    return _quantize_elemwise_core(A, bits=mantissa_bits + 2, exp_bits=exp_bits,
                                   max_norm=_get_max_norm(exp_bits, mantissa_bits + 2), round=round,
                                   allow_denorm=allow_denorm, custom_cuda=custom_cuda)


def quantize_elemwise_op(A, mx_specs, round=None):
    """bfloat / fp rounding selected by mx_specs (elemwise_ops.py:237-266)"""
    if mx_specs is None:
        return A
    if round is None:
        round = mx_specs["round"]
    bf, fp = mx_specs["bfloat"], mx_specs["fp"]
    if bf > 0 and fp > 0:
        raise ValueError("Cannot set both [bfloat] and [fp] in mx_specs.")
    if bf > 9:
        return _quantize_bfloat(A, bfloat=bf, round=round, custom_cuda=mx_specs["custom_cuda"],
                                allow_denorm=mx_specs["bfloat_subnorms"])
    if 0 < bf <= 9:
        raise ValueError("Cannot set [bfloat] <= 9 in mx_specs.")
    if fp > 6:
        return _quantize_fp(A, exp_bits=5, mantissa_bits=fp - 6, round=round, custom_cuda=mx_specs["custom_cuda"],
                            allow_denorm=mx_specs["bfloat_subnorms"])
    if 0 < fp <= 6:
        raise ValueError("Cannot set [fp] <= 6 in mx_specs.")
    return A
