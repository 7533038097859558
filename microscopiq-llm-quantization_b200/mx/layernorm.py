"""LayerNorm / RMSNorm with bfloat rounding after every vector operation (mirror of the forward passes of
number_system/mx/layernorm.py:18-40,68-99,102-142 and norm_utils._norm_forward :26-110)."""
import torch
import torch.nn.functional as F

from .specs import apply_mx_specs, mx_assert_test
from .vector_ops import (vec_add, vec_mul, vec_quantize, vec_recip, vec_reduce_mean, vec_sqrt, vec_sub)


def _norm_forward(x, axes, weight, bias, eps, mx_specs):
    """z = (x - mean) / sqrt(var + eps); output = z * weight + bias   (norm_utils.py:26-110, no groups / running stats)"""
    axes = [axes] if type(axes) is not list else axes
    x_mean = vec_reduce_mean(x, axes, keepdim=True, mx_specs=mx_specs)
    x_shift = vec_sub(x, x_mean, mx_specs=mx_specs)
    x_shift_pow2 = vec_mul(x_shift, x_shift, mx_specs=mx_specs)
    x_var = vec_reduce_mean(x_shift_pow2, axes, keepdim=True, mx_specs=mx_specs)
    x_vare = vec_add(x_var, eps, mx_specs=mx_specs)
    x_std = vec_sqrt(x_vare, mx_specs=mx_specs)
    x_std_inv = vec_recip(x_std, mx_specs=mx_specs)
    x_norm = vec_mul(x_shift, x_std_inv, mx_specs=mx_specs)
    x_scale = vec_mul(weight, x_norm, mx_specs=mx_specs)
    output = vec_add(x_scale, bias, mx_specs=mx_specs)
    return output, x_shift, x_norm, x_std_inv, x_mean, x_vare


def layer_norm(x, weight, bias, eps, mx_specs=None):
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)
    mx_specs = apply_mx_specs(mx_specs)
    x = vec_quantize(x, mx_specs=mx_specs)
    bf_weight = vec_quantize(weight, mx_specs=mx_specs)
    bf_bias = vec_quantize(bias, mx_specs=mx_specs)
    return _norm_forward(x, -1, bf_weight, bf_bias, eps, mx_specs)[0]


def rms_norm(x, weight, bias, eps, mx_specs=None):
    """layernorm.py:102-142 (forward)"""
    mx_specs = apply_mx_specs(mx_specs)
    x = vec_quantize(x, mx_specs=mx_specs)
    x2 = vec_mul(x, x, mx_specs=mx_specs)
    x_ms = vec_reduce_mean(x2, -1, keepdim=True, mx_specs=mx_specs)
    x_mse = vec_add(x_ms, eps, mx_specs=mx_specs)
    x_rms = vec_sqrt(x_mse, mx_specs=mx_specs)
    x_rms_inv = vec_recip(x_rms, mx_specs=mx_specs)
    x_norm = vec_mul(x, x_rms_inv, mx_specs=mx_specs)
    bf_weight = vec_quantize(weight, mx_specs=mx_specs)
    bf_bias = vec_quantize(bias, mx_specs=mx_specs)
    x_scale = vec_mul(bf_weight, x_norm, mx_specs=mx_specs)
    return vec_add(x_scale, bf_bias, mx_specs=mx_specs)


class LayerNorm(torch.nn.LayerNorm):
    def __init__(self, hidden_size, eps=1e-12, mx_specs=None, name=None):
        mx_assert_test(mx_specs)
        self.mx_none = mx_specs is None
        self.name = name
        self.mx_specs = apply_mx_specs(mx_specs)
        super().__init__(normalized_shape=hidden_size, eps=eps)

    def apply_mx_specs(self, mx_specs):
        self.mx_none = mx_specs is None
        self.mx_specs = apply_mx_specs(mx_specs)

    def append_name(self, postfix):
        self.name += postfix

    def forward(self, x):
        if self.mx_none:
            return super().forward(x)
        with torch.no_grad():
            return layer_norm(x, self.weight, self.bias, self.eps, self.mx_specs)


class RMSNorm(LayerNorm):
    def forward(self, x):
        if self.mx_none:
            raise NotImplementedError("RMSNorm needs mx_specs (the reference has no baseline path either)")
        with torch.no_grad():
            return rms_norm(x, self.weight, self.bias, self.eps, self.mx_specs)
