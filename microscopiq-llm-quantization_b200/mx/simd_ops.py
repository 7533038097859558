"""SIMD element-wise operators with bfloat rounding of operands and results (mirror of the forward passes of
number_system/mx/simd_ops.py:85-545; this package is inference-only)."""
import torch

from .specs import apply_mx_specs, mx_assert_test
from .vector_ops import (vec_add, vec_div, vec_exp, vec_mul, vec_quantize, vec_reduce_mean, vec_reduce_sum, vec_sqrt,
                         vec_sub)

torch_sqrt, torch_exp, torch_log, torch_square = torch.sqrt, torch.exp, torch.log, torch.square


def _binary(op, vec_op, in1, in2, mx_specs):
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return op(in1, in2)
    mx_specs = apply_mx_specs(mx_specs)
    assert isinstance(in1, torch.Tensor)
    q1 = vec_quantize(in1, mx_specs=mx_specs)
    q2 = vec_quantize(in2, mx_specs=mx_specs) if isinstance(in2, torch.Tensor) else in2
    return vec_op(q1, q2, mx_specs=mx_specs)


def simd_add(in1, in2, mx_specs=None):
    return _binary(lambda a, b: a + b, vec_add, in1, in2, mx_specs)


def simd_sub(in1, in2, mx_specs=None):
    return _binary(lambda a, b: a - b, vec_sub, in1, in2, mx_specs)


def simd_mul(in1, in2, mx_specs=None):
    return _binary(lambda a, b: a * b, vec_mul, in1, in2, mx_specs)


def simd_div(in1, in2, mx_specs=None):
    return _binary(lambda a, b: a / b, vec_div, in1, in2, mx_specs)


def simd_split(in1, mx_specs=None):
    """forward: two references to the same tensor (simd_ops.py:262-283: the backward adds the two gradients)"""
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return in1, in1
    return in1, in1.clone()


def _unary(op, vec_fn, in1, mx_specs):
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return op(in1)
    mx_specs = apply_mx_specs(mx_specs)
    return vec_fn(vec_quantize(in1, mx_specs=mx_specs), mx_specs)


def simd_square(in1, mx_specs=None):
    return _unary(torch_square, lambda q, s: vec_mul(q, q, mx_specs=s), in1, mx_specs)


def simd_sqrt(in1, mx_specs=None):
    return _unary(torch_sqrt, lambda q, s: vec_sqrt(q, mx_specs=s), in1, mx_specs)


def simd_exp(in1, mx_specs=None):
    return _unary(torch_exp, lambda q, s: vec_exp(q, mx_specs=s), in1, mx_specs)


def simd_log(in1, mx_specs=None):
    from .elemwise_ops import quantize_elemwise_op
    return _unary(torch_log, lambda q, s: quantize_elemwise_op(torch_log(q), mx_specs=s), in1, mx_specs)


def simd_reduce_sum(in1, dim=None, keepdim=False, mx_specs=None):
    mx_assert_test(mx_specs)
    if dim is None:
        dim = list(range(in1.ndim))
    if mx_specs is None:
        return torch.sum(in1, dim, keepdim=keepdim)
    mx_specs = apply_mx_specs(mx_specs)
    dim = [dim] if type(dim) is int else dim
    return vec_reduce_sum(vec_quantize(in1, mx_specs=mx_specs), dim, keepdim=keepdim, mx_specs=mx_specs)


def simd_reduce_mean(in1, dim=None, keepdim=False, mx_specs=None):
    mx_assert_test(mx_specs)
    if dim is None:
        dim = list(range(in1.ndim))
    if mx_specs is None:
        return torch.mean(in1, dim, keepdim=keepdim)
    mx_specs = apply_mx_specs(mx_specs)
    dim = [dim] if type(dim) is int else dim
    return vec_reduce_mean(vec_quantize(in1, mx_specs=mx_specs), dim, keepdim=keepdim, mx_specs=mx_specs)


def simd_norm(in1, keepdim=False, mx_specs=None):
    """sqrt(sum(in1**2)) over all elements (simd_ops.py:544-553)"""
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return torch.linalg.norm(in1)
    mx_specs = apply_mx_specs(mx_specs)
    q = vec_quantize(in1, mx_specs=mx_specs)
    s = vec_reduce_sum(vec_mul(q, q, mx_specs=mx_specs), list(range(q.ndim)), keepdim=keepdim, mx_specs=mx_specs)
    return vec_sqrt(s, mx_specs=mx_specs)
