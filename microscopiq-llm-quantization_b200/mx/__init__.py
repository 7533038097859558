"""`mx` namespace of the reference (number_system/mx/__init__.py): format table, specs and
quantization entry points, backed by libmsq_b200.so."""
from .formats import ElemFormat, RoundingMode, _get_format_params  # noqa: F401
from .specs import (MxSpecs, add_mx_args, apply_mx_specs, finalize_mx_specs, get_backwards_mx_specs,  # noqa: F401
                    get_default_mx_specs, get_mx_specs, mx_assert_test)
from .elemwise_ops import quantize_elemwise_op  # noqa: F401
from .mx_ops import quantize_mx_op, quantize_mx_outlier_op  # noqa: F401
from .linear import MXLinear, linear  # noqa: F401,E402
from . import custom_extensions  # noqa: F401,E402  (custom_extensions.funcs: the reference's seven pybind names)
from .linear import MXLinear as Linear  # noqa: F401,E402  (`from mx import Linear`, examples/scratch_3.py:9)
from .layernorm import LayerNorm, RMSNorm  # noqa: F401,E402
from .activations import gelu, relu, relu6, leaky_relu, silu, sigmoid, tanh, GELU  # noqa: F401,E402
from .simd_ops import (simd_add, simd_sub, simd_mul, simd_div, simd_split, simd_square, simd_sqrt, simd_exp,  # noqa: F401,E402
                       simd_log, simd_reduce_sum, simd_reduce_mean, simd_norm)
from .quantize import quantize_bfloat  # noqa: F401,E402
