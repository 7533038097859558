"""MxSpecs configuration dictionary (mirror of number_system/mx/specs.py:61-355).

Key names, defaults and the dependency rules of `finalize_mx_specs` are the contract.
"""
import argparse
import collections
import json
import os
import traceback

_ASSERT_MODE = os.environ.get("MX_ASSERT", "False")

# key -> (default, help)
_SPEC_TABLE = collections.OrderedDict([
    ("scale_bits", (0, "Bits (sign + magnitude) to use for shared exponent/scale")),
    ("w_elem_format", (None, "Weight MX elem format, one of {fp8_e5m2, fp8_e4m3, fp6_e3m2, fp6_e2m3, fp4_e2m1, int8, int4}")),
    ("a_elem_format", (None, "Activation MX elem format. See w_elem_format")),
    ("w_elem_format_bp", (None, "Backpass weight MX elem format. See w_elem_format")),
    ("a_elem_format_bp_ex", (None, "Backpass act MX elem format. See w_elem_format")),
    ("a_elem_format_bp_os", (None, "Backpass act MX elem format. See w_elem_format")),
    ("mx_flush_fp32_subnorms", (False, "MX quantization flushes blocks with subnormal shared scale to zero")),
    ("shared_exp_method", ("max", "Shared exponent calculation method. Options: max, none")),
    ("block_size", (0, "mx shared exponent block size")),
    ("bfloat", (0, "BfloatX format (8exp + sign + mantissa). Only one of bfloat or fp can be used")),
    ("fp", (0, "fpX format (5exp + sign + mantissa). Only one of bfloat or fp can be used")),
    ("bfloat_subnorms", (True, "Bfloat/FP supports subnorms")),
    ("quantize_backprop", (True, "Enable mx/bfloat quantization on backward pass")),
    ("round", ("nearest", "Global rounding mode. Choices: nearest, floor")),
    ("round_m", ("nearest", "ADAM optimizer m and v rounding mode")),
    ("round_weight", ("nearest", "Weight bfloat rounding mode (W in WAGE)")),
    ("round_output", ("nearest", "Activation bfloat rounding mode (A in WAGE)")),
    ("round_grad_weight", ("nearest", "Weight update rounding mode (G in WAGE)")),
    ("round_grad_input", ("nearest", "Error gradient rounding mode (E in WAGE)")),
    ("round_mx_output", ("nearest", "Forward pass mx rounding mode")),
    ("round_mx_input_grad_input", ("nearest", "")),
    ("round_mx_weight_grad_input", ("nearest", "")),
    ("round_mx_grad_output_grad_input", ("nearest", "")),
    ("round_mx_input_grad_weight", ("nearest", "")),
    ("round_mx_grad_output_grad_weight", ("nearest", "")),
    ("softmax_exp2", (False, "Softmax uses 2^x instead of e^x")),
    ("vec_use_exp2", (False, "Use 2^x to compute e^x")),
    ("vec_use_recip", (False, "Use 1/x to compute division")),
    ("custom_cuda", (False, "Enable custom CUDA kernels for quantization")),
])

# (dependent key, key it inherits from) -- applied in order by finalize_mx_specs
_INHERIT = [
    ("w_elem_format_bp", "w_elem_format"), ("a_elem_format_bp_os", "a_elem_format"),
    ("a_elem_format_bp_ex", "a_elem_format"),
    ("round_m", "round"), ("round_output", "round"), ("round_grad_weight", "round"),
    ("round_grad_input", "round"), ("round_weight", "round"), ("round_mx_output", "round"),
    ("round_mx_input_grad_input", "round_grad_input"), ("round_mx_weight_grad_input", "round_grad_input"),
    ("round_mx_grad_output_grad_input", "round_grad_input"), ("round_mx_input_grad_weight", "round_grad_input"),
    ("round_mx_grad_output_grad_weight", "round_grad_input"),
]
_QUANT_KEYS = ("w_elem_format", "a_elem_format", "w_elem_format_bp", "a_elem_format_bp_os",
               "a_elem_format_bp_ex", "bfloat", "fp")


class MxSpecs(collections.UserDict):
    """dict of quantization parameters; unknown keys are rejected"""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.help_strings = {k: h for k, (_, h) in _SPEC_TABLE.items()}
        for k, (default, _) in _SPEC_TABLE.items():
            self.data.setdefault(k, default)
        for k in self.data:
            assert k in _SPEC_TABLE

    def safe_json(self, indent=None):
        return json.dumps(self.data, indent=indent,
                          default=lambda o: "<<non-serializable: %s>>" % type(o).__qualname__)

    def __str__(self):
        return self.safe_json(indent=4)


def get_default_mx_specs():
    return MxSpecs()


def get_backwards_mx_specs(specs):
    """a spec that disables quantization when quantize_backprop is False"""
    b = specs.copy()
    if b["quantize_backprop"] is False or b["quantize_backprop"] == False:  # noqa: E712
        for k in ("w_elem_format", "a_elem_format", "w_elem_format_bp", "a_elem_format_bp_os", "a_elem_format_bp_ex"):
            b[k] = None
        b["block_size"] = 0
        b["bfloat"] = 0
        b["fp"] = 0
    return b


def apply_mx_specs(mx_specs, default_mx_specs=None):
    base = default_mx_specs if default_mx_specs else get_default_mx_specs()
    if not mx_specs:
        return base
    for k in mx_specs:
        if mx_specs[k] is not None:
            if k not in base:
                raise KeyError("Unknown key '%s' passed to mx specs" % k)
            base[k] = mx_specs[k]
    return base


def add_mx_args(parser):
    """one --flag per spec key, typed from the default (bools become store_true / --no_*)"""
    grp = parser.add_argument_group("mx", "MX specs")
    grp.add_argument("--mx_dir", type=str, default=None, help="Path to mx library")
    for k, (default, help_str) in _SPEC_TABLE.items():
        help_str = help_str or "No help string"
        if "elem_format" in k:
            grp.add_argument("--" + k, type=str, default=default, help=help_str)
        elif default is False:
            grp.add_argument("--" + k, action="store_true", help=help_str)
        elif default is True:
            grp.add_argument("--no_" + k, action="store_true", help=help_str)
        else:
            grp.add_argument("--" + k, type=type(default), default=None, help=help_str)
    grp.add_argument("--skip_early_exit", action="store_true", default=False,
                     help="Don't early exit if no quantization is specified")
    return parser


def finalize_mx_specs(specs, early_exit=True):
    """fill dependent keys; returns None when nothing is quantized (and early_exit)"""
    if early_exit and not any(specs.get(k, 0) for k in _QUANT_KEYS):
        return None
    for dst, src in _INHERIT:
        if (dst not in specs or specs[dst] is None) and src in specs:
            specs[dst] = specs[src]
    return apply_mx_specs(specs, get_default_mx_specs())


def get_mx_specs(parsed_args):
    parsed = {}
    for k, (default, _) in _SPEC_TABLE.items():
        name = "no_" + k if default is True else k
        if hasattr(parsed_args, name):
            v = getattr(parsed_args, name)
            parsed[k] = (not v) if default is True else v
    early_exit = not getattr(parsed_args, "skip_early_exit", False)
    return finalize_mx_specs(parsed, early_exit=early_exit)


def mx_assert_test(mx_specs):
    if _ASSERT_MODE == "True" and mx_specs is None:
        stack = traceback.extract_stack()
        f1, f2 = stack[-2], stack[-3]
        raise ValueError("MX assert test failed!\nmx_specs is None in function %s\nCalled from %s, line %d\n  %s"
                         % (f1.name, f2.filename, f2.lineno, f2.line))
