"""Non-differentiable vector operations (mirror of number_system/mx/vector_ops.py): every result is rounded to the
bfloat / fp format of mx_specs by the B200 element-wise kernel (msq_quantize_elemwise); the arithmetic between two
roundings is one torch element-wise op on the device, exactly the op the reference issues."""
import numpy as np
import torch

from .elemwise_ops import quantize_elemwise_op

LN_2_EXACT = 0.69314718056
LOG2_E_BF16 = 1.4453125  # 1 + 2**-2 + 2**-3 + 2**-4 + 2**-7


def vec_quantize(input, mx_specs=None, round=None):
    return quantize_elemwise_op(input, mx_specs=mx_specs, round=round)


def vec_add(a, b, mx_specs=None, round=None):
    return quantize_elemwise_op(a + b, mx_specs=mx_specs, round=round)


def vec_sub(a, b, mx_specs=None, round=None):
    return quantize_elemwise_op(a - b, mx_specs=mx_specs, round=round)


def vec_mul(a, b, mx_specs=None, round=None):
    return quantize_elemwise_op(a * b, mx_specs=mx_specs, round=round)


def vec_div(a, b, mx_specs=None, round=None):
    if mx_specs and mx_specs["vec_use_recip"]:
        recip_b = vec_recip(b, mx_specs=mx_specs, round=round)
        return vec_mul(a, recip_b, mx_specs=mx_specs, round=round)
    return quantize_elemwise_op(a / b, mx_specs=mx_specs, round=round)


def vec_exp(input, mx_specs=None, round=None):
    if mx_specs and mx_specs["vec_use_exp2"]:
        phi = quantize_elemwise_op(LOG2_E_BF16 * input, mx_specs=mx_specs, round=round)
        return vec_exp2(phi, mx_specs=mx_specs, round=round)
    return quantize_elemwise_op(torch.exp(input), mx_specs=mx_specs, round=round)


def vec_exp2(input, mx_specs=None, round=None):
    return quantize_elemwise_op(torch.exp2(input), mx_specs=mx_specs, round=round)


def vec_recip(input, mx_specs=None, round=None):
    return quantize_elemwise_op(1. / input, mx_specs=mx_specs, round=round)


def vec_sqrt(input, mx_specs=None, round=None):
    return quantize_elemwise_op(torch.sqrt(input), mx_specs=mx_specs, round=round)


def vec_tanh(input, mx_specs=None, round=None):
    return quantize_elemwise_op(torch.tanh(input), mx_specs=mx_specs, round=round)


def vec_reduce_sum(input, dim, keepdim=False, mx_specs=None, round=None):
    return quantize_elemwise_op(input.sum(dim, keepdim=keepdim), mx_specs=mx_specs, round=round)


def vec_reduce_mean(input, dim, keepdim=False, mx_specs=None, round=None):
    dim = dim if type(dim) is list else [dim]
    denom = np.prod([input.shape[i] for i in dim])
    s = vec_reduce_sum(input, dim, keepdim=keepdim, mx_specs=mx_specs, round=round)
    return vec_div(s, denom, mx_specs=mx_specs, round=round)
