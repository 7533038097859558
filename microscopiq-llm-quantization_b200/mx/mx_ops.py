"""MX block quantization on B200 -- mirror of number_system/mx/mx_ops.py.

`_quantize_mx`, `quantize_mx_op`, `_quantize_mx_outlier_v1`, `quantize_mx_outlier_op`
keep the reference's signatures.  New keyword-only options are additive.
"""
from .. import _ops
from .formats import ElemFormat, RoundingMode, _get_format_params  # noqa: F401
from .specs import mx_assert_test


def _quantize_mx(A, scale_bits, elem_format, shared_exp_method="max", axes=None, block_size=0, round="nearest",
                 flush_fp32_subnorms=False, custom_cuda=False, *, compat_eps=None):
    """mx_ops.py:332-457.  Arithmetic mode: custom_cuda=False reproduces the reference's
    torch path including the fork's `2^e + 1e-6` divisor (mx_ops.py:444); custom_cuda=True
    reproduces the native path (cpp/mx.cuh) that the reference's golden tests pin.
    compat_eps (bool) overrides that choice explicitly."""
    if elem_format is None:
        return A
    assert scale_bits > 0
    if round == "dither":
        raise NotImplementedError("dither rounding is not part of the B200 contract")
    spec = bool(custom_cuda) and round in RoundingMode.string_enums()
    if compat_eps is not None:
        spec = not compat_eps
    if shared_exp_method == "none":
        x = A.reshape(-1, 1)
        return _ops.mx_plain(x, scale_bits, elem_format, 1, 1, round, flush_fp32_subnorms, int(spec)).reshape(A.shape)
    if shared_exp_method != "max":
        raise Exception("Unrecognized shared exponent selection method %s" % (shared_exp_method))
    axis = _ops._single_axis(axes, A.ndim)
    return _ops.mx_plain(A, scale_bits, elem_format, axis, block_size, round, flush_fp32_subnorms, int(spec))


def quantize_mx_op(A, mx_specs, elem_format=None, block_size=None, axes=None, round="nearest",
                   expand_and_reshape=False):
    """mx_ops.py:460-490 (scale_bits 0 -> 8)"""
    mx_assert_test(mx_specs)
    if elem_format is None:
        return A
    if isinstance(elem_format, str):
        elem_format = ElemFormat.from_str(elem_format)
    if block_size is None:
        block_size = mx_specs["block_size"]
    scale_bits = mx_specs["scale_bits"] or 8
    return _quantize_mx(A, scale_bits, elem_format, block_size=block_size, axes=axes, round=round,
                        shared_exp_method=mx_specs["shared_exp_method"],
                        flush_fp32_subnorms=mx_specs["mx_flush_fp32_subnorms"], custom_cuda=mx_specs["custom_cuda"])


def _quantize_mx_outlier_v1(A, inlier_scale_bits, outlier_scale_bits, inlier_elem_format, outlier_elem_format,
                            shared_exp_method="max", std_dev=5, axes=None, block_size=0, round="nearest",
                            flush_fp32_subnorms=False, custom_cuda=False):
    """mx_ops.py:210-330: detector statistics over the block-count axis (mx_ops.py:248),
    underflowing exponents clamp to -scale_emax."""
    if inlier_elem_format is None:
        return A
    assert inlier_scale_bits > 0 and outlier_scale_bits > 0
    if shared_exp_method != "max":
        raise NotImplementedError("shared_exp_method=%r for the outlier quantizer" % (shared_exp_method,))
    axis = _ops._single_axis(axes, A.ndim)
    if not block_size:
        raise NotImplementedError("block_size=0 with the across-block detector")
    return _ops.mx_outlier(A, inlier_scale_bits, outlier_scale_bits, inlier_elem_format, outlier_elem_format,
                           std_dev, axis, block_size, round, flush_fp32_subnorms, variant=1).deq


def quantize_mx_outlier_op(A, mx_specs, inlier_elem_format=None, outlier_elem_format=None, block_size=None,
                           axes=None, round="nearest", expand_and_reshape=False):
    """mx_ops.py:492-533 (scale_bits 0 -> 4, outlier scale bits = inlier's)"""
    mx_assert_test(mx_specs)
    if inlier_elem_format is None or outlier_elem_format is None:
        return A
    if isinstance(inlier_elem_format, str):
        inlier_elem_format = ElemFormat.from_str(inlier_elem_format)
    if isinstance(outlier_elem_format, str):
        outlier_elem_format = ElemFormat.from_str(outlier_elem_format)
    if block_size is None:
        block_size = mx_specs["block_size"]
    sb = mx_specs["scale_bits"] or 4
    return _quantize_mx_outlier_v1(A, sb, sb, inlier_elem_format, outlier_elem_format, block_size=block_size,
                                   axes=axes, round=round, shared_exp_method=mx_specs["shared_exp_method"],
                                   flush_fp32_subnorms=mx_specs["mx_flush_fp32_subnorms"],
                                   custom_cuda=mx_specs["custom_cuda"])
