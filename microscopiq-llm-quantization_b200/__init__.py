"""microscopiq-llm-quantization_b200 -- B200-native MicroScopiQ quantization hot path.

Import as `msq_b200` (the repo-root shim package) or put this directory on sys.path to use
the reference's own module names (`utils.quant`, `mx.mx_ops`, `llm.gptq`).
"""
from . import _native  # noqa: F401
from .mx.formats import ElemFormat, RoundingMode, _get_format_params  # noqa: F401
from .mx.specs import MxSpecs, finalize_mx_specs, apply_mx_specs, add_mx_args, get_mx_specs  # noqa: F401
from .mx.elemwise_ops import quantize_elemwise_op, _quantize_elemwise_core  # noqa: F401
from .mx.mx_ops import _quantize_mx, quantize_mx_op, _quantize_mx_outlier_v1, quantize_mx_outlier_op  # noqa: F401
from .utils.quant import (quantize_mx_outlier_v1, quantize_mx_outlier_hessian, quantize, Quantizer,  # noqa: F401
                          MXQuantizer, quantize_mx_outlier_host, host_sync)

__version__ = "0.1.0"
from .llm.gptq import GPTQ  # noqa: F401,E402
from .packed import PackedMX, pack_mx, pack_mx_outlier  # noqa: F401,E402
from .mx.linear import MXLinear, linear  # noqa: F401,E402
