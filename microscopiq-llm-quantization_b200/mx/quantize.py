"""quantize_bfloat (mirror of number_system/mx/quantize.py:13-24, forward)"""
from .elemwise_ops import quantize_elemwise_op
from .specs import apply_mx_specs, mx_assert_test


def quantize_bfloat(x, mx_specs, round=None):
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return x
    mx_specs = apply_mx_specs(mx_specs)
    if round is None:
        round = mx_specs["round"]
    return quantize_elemwise_op(x, mx_specs=mx_specs, round=round)
