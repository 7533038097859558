"""Activation functions with bfloat rounding after every vector operation (mirror of the forward passes of
number_system/mx/activations.py:28-108, 239-520)."""
import torch
import torch.nn.functional as F

from .specs import apply_mx_specs, mx_assert_test
from .vector_ops import vec_add, vec_exp, vec_mul, vec_quantize, vec_recip, vec_tanh


def _sigmoid_core(q_in, mx_specs):
    e = vec_exp(-q_in, mx_specs=mx_specs)
    e = vec_add(e, 1., mx_specs=mx_specs)
    return vec_recip(e, mx_specs=mx_specs)


def sigmoid(input, mx_specs=None, name=None):
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return torch.sigmoid(input)
    mx_specs = apply_mx_specs(mx_specs)
    return _sigmoid_core(vec_quantize(input, mx_specs=mx_specs), mx_specs)


def tanh(input, mx_specs=None, name=None):
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return torch.tanh(input)
    mx_specs = apply_mx_specs(mx_specs)
    return vec_tanh(vec_quantize(input, mx_specs=mx_specs), mx_specs=mx_specs)


def relu(input, inplace=False, mx_specs=None, name=None):
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return F.relu(input, inplace=inplace)
    mx_specs = apply_mx_specs(mx_specs)
    return F.relu(vec_quantize(input, mx_specs=mx_specs))


def relu6(input, inplace=False, mx_specs=None, name=None):
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return F.relu6(input, inplace=inplace)
    mx_specs = apply_mx_specs(mx_specs)
    return F.relu6(vec_quantize(input, mx_specs=mx_specs))


def leaky_relu(input, negative_slope=0.01, inplace=False, mx_specs=None, name=None):
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return F.leaky_relu(input, negative_slope=negative_slope, inplace=inplace)
    mx_specs = apply_mx_specs(mx_specs)
    q_in = vec_quantize(input, mx_specs=mx_specs)
    return vec_quantize(F.leaky_relu(q_in, negative_slope=negative_slope), mx_specs=mx_specs)


def silu(input, inplace=False, mx_specs=None, name=None):
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return F.silu(input, inplace=inplace)
    mx_specs = apply_mx_specs(mx_specs)
    q_in = vec_quantize(input, mx_specs=mx_specs)
    return vec_mul(q_in, _sigmoid_core(q_in, mx_specs), mx_specs=mx_specs)


def gelu(input, mx_specs=None, first_order_gelu=False, approximate=None, name=None):
    """activations.py:85-107 / GELUFunction.forward :460-506: x * sigmoid(1.5958 (x + 0.044715 x^3)) with bfloat16
    coefficients (first_order_gelu: x * sigmoid(1.702 x))"""
    mx_assert_test(mx_specs)
    if mx_specs is None and not first_order_gelu:
        return F.gelu(input, approximate="tanh")
    mx_specs = apply_mx_specs(mx_specs)
    q_in = vec_quantize(input, mx_specs=mx_specs)
    if first_order_gelu:
        s = vec_mul(1.703125, q_in, mx_specs=mx_specs)
    else:
        s = vec_mul(q_in, q_in, mx_specs=mx_specs)
        s = vec_mul(s, q_in, mx_specs=mx_specs)
        s = vec_mul(0.044677734, s, mx_specs=mx_specs)
        s = vec_add(s, q_in, mx_specs=mx_specs)
        s = vec_mul(1.59375, s, mx_specs=mx_specs)
    phi = vec_exp(-s, mx_specs=mx_specs)
    phi = vec_add(phi, 1., mx_specs=mx_specs)
    phi = vec_recip(phi, mx_specs=mx_specs)
    return vec_mul(q_in, phi, mx_specs=mx_specs)


class GELU(torch.nn.GELU):
    def __init__(self, mx_specs=None, first_order_gelu=False, name=None):
        super().__init__()
        mx_assert_test(mx_specs)
        self.mx_specs, self.first_order_gelu, self.name = mx_specs, first_order_gelu, name

    def forward(self, input):
        return gelu(input, mx_specs=self.mx_specs, first_order_gelu=self.first_order_gelu)
