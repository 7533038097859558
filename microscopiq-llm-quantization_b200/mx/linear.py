"""MX linear layer on B200 -- mirror of number_system/mx/linear.py (forward path).

    linear(input, weight, bias=None, mx_specs=None, name=None)       linear.py:204-217
    MXLinear(in_features, out_features, bias=True, mx_specs=None)     linear.py:220-254

Forward = the reference's LinearFunction.forward (linear.py:20-108): bfloat/fp rounding of input,
weight and bias, MX+outlier quantization of both operands along the input-feature axis (axes=[1],
across-block detector of mx_ops.py), fp32 matmul, rounding of the output, bias add, rounding.
The matmul runs on the packed weight with tcgen05 (csrc/msq_gemm_w4.cu, csrc/msq_gemm_mx8.cu) whenever the
weight is a 2-D tensor blocked by 32 along K with an element format of at most 8 bits; otherwise
(block size 128 of llm/opt_direct.py:98-105, ragged K, w_elem_format None ...) on the dequantized operands
with the dense tcgen05 GEMM (_tc.dense_linear).  The reference's backward is broken as shipped and is not
part of the contract: this layer is inference-only.
"""
import torch
import torch.nn.functional as F

from .. import _tc
from ..packed import pack_mx, pack_mx_outlier
from .elemwise_ops import quantize_elemwise_op
from .mx_ops import quantize_mx_outlier_op
from .specs import apply_mx_specs, mx_assert_test

f_linear = F.linear
torch_matmul = torch.matmul


def _pack_weight(bf_weight, mx_specs):
    fmt = mx_specs["w_elem_format"]
    sb = mx_specs["scale_bits"] or 4                       # mx_ops.py:519-524
    return pack_mx_outlier(bf_weight, sb, sb, fmt, fmt, 5, [1], mx_specs["block_size"], round=mx_specs["round_mx_output"],
                           flush_fp32_subnorms=mx_specs["mx_flush_fp32_subnorms"], variant=1)


def _packable(weight, mx_specs):
    return (mx_specs["w_elem_format"] is not None and weight.dim() == 2 and mx_specs["block_size"] == 32
            and weight.shape[1] % 32 == 0 and weight.shape[1] // 32 >= 2 and mx_specs["shared_exp_method"] == "max")


def _mx8_activation_ok(x, mx_specs):
    from .formats import _get_format_params
    a = mx_specs["a_elem_format"]
    return (a is not None and x.dim() == 2 and mx_specs["block_size"] == 32 and x.shape[1] % 32 == 0
            and x.shape[1] // 32 >= 2 and _get_format_params(a)[:2] == (4, 5) and x.dtype == torch.float32)


def _mx4_activation_ok(x, mx_specs):
    from .formats import _get_format_params
    a = mx_specs["a_elem_format"]
    return (a is not None and x.dim() == 2 and mx_specs["block_size"] == 32 and x.shape[1] % 256 == 0
            and _get_format_params(a)[:2] == (2, 3) and x.dtype == torch.float32)


def _bf16_exact_activations(mx_specs):
    """are the (fake-quantized) activations exactly representable in bfloat16?  MX element formats with at most 8
    significant bits times a power of two are (fp32 subnormals aside); unquantized activations are when the spec rounds
    them to bfloat16 first (quantize_elemwise_op, number_system/mx/linear.py:57)."""
    from .formats import _get_format_params
    a = mx_specs["a_elem_format"]
    if a is None:
        return mx_specs["bfloat"] == 16
    return _get_format_params(a)[1] - 1 <= 8


def linear_forward(input, weight, bias, mx_specs, packed_weight=None):
    bf_in = quantize_elemwise_op(input, mx_specs=mx_specs, round=mx_specs["round_output"])
    bf_bias = None
    if bias is not None:
        bf_bias = quantize_elemwise_op(bias, mx_specs=mx_specs, round=mx_specs["round_weight"])
    # MX quantize the activations along the input-feature axis (linear.py:66-73: axes=[1])
    qis_input = quantize_mx_outlier_op(bf_in, mx_specs, inlier_elem_format=mx_specs["a_elem_format"],
                                       outlier_elem_format=mx_specs["a_elem_format"], axes=[1],
                                       round=mx_specs["round_mx_output"])
    if packed_weight is None and _packable(weight, mx_specs):
        bf_weight = quantize_elemwise_op(weight, mx_specs=mx_specs, round=mx_specs["round_weight"])
        packed_weight = _pack_weight(bf_weight, mx_specs)
    if packed_weight is not None and qis_input.shape[-1] == packed_weight.shape[1]:
        output = None
        if _mx4_activation_ok(bf_in, mx_specs) and _tc.mx4_pair_ok(packed_weight, bf_in.shape[0]):
            # W4A4: both operands as packed FP4 codes on kind::mxf4 (twice the tensor rate of the FP8 pairing)
            sb = mx_specs["scale_bits"] or 4
            xq = pack_mx_outlier(bf_in, sb, sb, mx_specs["a_elem_format"], mx_specs["a_elem_format"], 5, [1], 32,
                                 round=mx_specs["round_mx_output"], flush_fp32_subnorms=mx_specs["mx_flush_fp32_subnorms"],
                                 variant=1)
            if not bool((xq.bitmap != 0).any()):
                output = _tc.packed_linear_mx4(xq, packed_weight, assume_clean=True).to(input.dtype)
        if output is None and _mx8_activation_ok(bf_in, mx_specs) and _tc.mx8_weight_ok(packed_weight):
            # native block-scaled path: needs the activations as MX-FP8 codes with no outliers of their own
            sb = mx_specs["scale_bits"] or 4
            xq = pack_mx_outlier(bf_in, sb, sb, mx_specs["a_elem_format"], mx_specs["a_elem_format"], 5, [1], 32,
                                 round=mx_specs["round_mx_output"], flush_fp32_subnorms=mx_specs["mx_flush_fp32_subnorms"],
                                 variant=1)
            if not bool((xq.bitmap != 0).any()):
                output = _tc.packed_linear_mx8(xq, packed_weight).to(input.dtype)
        if output is None:
            xg = qis_input
            if xg.dtype == torch.float32 and _bf16_exact_activations(mx_specs):
                # every value is q * 2^e with <= 8 significant bits (or already rounded to bfloat16): the bf16 cast is
                # exact, one GEMM pass instead of three bf16 terms, and the W4A16 block-scaled path applies
                xg = xg.to(torch.bfloat16)
            output = _tc.packed_linear(xg, packed_weight).to(input.dtype)
    else:
        # not packable (block size other than 32, ragged K, w_elem_format None, ...): linear.py:78-91 on the
        # dequantized operands -- fake-quantize the weight, then the dense tcgen05 contraction
        bf_weight = quantize_elemwise_op(weight, mx_specs=mx_specs, round=mx_specs["round_weight"])
        qis_weight = quantize_mx_outlier_op(bf_weight, mx_specs, inlier_elem_format=mx_specs["w_elem_format"],
                                            outlier_elem_format=mx_specs["w_elem_format"], axes=[1],
                                            round=mx_specs["round_mx_output"])
        output = _tc.dense_linear(qis_input, qis_weight).to(input.dtype)
    output = quantize_elemwise_op(output, mx_specs=mx_specs, round=mx_specs["round_output"])
    if bf_bias is not None:
        output = quantize_elemwise_op(output + bf_bias, mx_specs=mx_specs, round=mx_specs["round_output"])
    return output


def linear(input, weight, bias=None, mx_specs=None, name=None):
    mx_assert_test(mx_specs)
    if mx_specs is None:
        return f_linear(input, weight, bias=bias)
    mx_specs = apply_mx_specs(mx_specs)
    with torch.no_grad():
        return linear_forward(input, weight, bias, mx_specs)


class MXLinear(torch.nn.Linear):
    def __init__(self, in_features, out_features, bias=True, mx_specs=None, name=None):
        mx_assert_test(mx_specs)
        self.mx_none = mx_specs is None
        self.name = name
        self.mx_specs = apply_mx_specs(mx_specs)
        self._packed = None
        self._packed_version = None
        super().__init__(in_features, out_features, bias)

    @classmethod
    def from_packed(cls, pw, bias=None, mx_specs=None, name=None, keep_dense=False, fast_activations=False):
        """inference module over a packed weight (PackedMX [out_features, in_features], e.g. loaded from an MSQPACK
        checkpoint or emitted by GPTQ / pack_mx_outlier): the stored planes ARE the weight -- nothing is re-quantized,
        the forward multiplies the (MX-quantized) activations with the packed planes on tcgen05.
        keep_dense=True also materialises the dequantized fp32 weight in `.weight` (state_dict compatibility);
        otherwise `.weight` is an empty placeholder."""
        out_f, in_f = pw.shape
        specs = dict(mx_specs) if mx_specs is not None else {}
        specs.setdefault("block_size", 32)
        self = cls.__new__(cls)
        torch.nn.Module.__init__(self)
        self.in_features, self.out_features = in_f, out_f
        dev = pw.codes.device
        w = pw.dequantize() if keep_dense else torch.empty(0, device=dev)
        self.weight = torch.nn.Parameter(w, requires_grad=False)
        self.bias = None if bias is None else torch.nn.Parameter(bias.detach().to(dev).float(), requires_grad=False)
        self.mx_none = False
        self.name = name
        mx_assert_test(specs)
        self.mx_specs = apply_mx_specs(specs)
        self._packed, self._packed_version, self._frozen = pw, None, True
        self.fast_activations = bool(fast_activations)
        return self

    def apply_mx_specs(self, mx_specs):
        mx_assert_test(mx_specs)
        self.mx_none = mx_specs is None
        self.mx_specs = apply_mx_specs(mx_specs)
        if not getattr(self, "_frozen", False):
            self._packed = None

    def append_name(self, postfix):
        self.name += postfix

    def packed_weight(self):
        """quantize + pack the weight once; reused until the parameter is modified"""
        if getattr(self, "_frozen", False):
            return self._packed
        ver = (self.weight._version, self.weight.data_ptr())
        if self._packed is None or self._packed_version != ver:
            with torch.no_grad():
                bf_w = quantize_elemwise_op(self.weight.data, mx_specs=self.mx_specs, round=self.mx_specs["round_weight"])
                self._packed = _pack_weight(bf_w, self.mx_specs) if _packable(self.weight, self.mx_specs) else None
            self._packed_version = ver
        return self._packed

    #: serving mode (opt-in; from_packed(..., fast_activations=True) or set on the module): the activations go through
    #: the one-launch plain-MX activation quantizer (msq_quantize_mx_act: the MX spec's shared-exponent blocks of 32, no
    #: outlier detection on the ACTIVATIONS) straight into the block-scaled GEMM on the packed weight -- W4A4 on
    #: kind::mxf4, W4A8 on kind::mxf8f6f4.  linear_forward (the default) keeps the reference's semantics, whose
    #: activation quantizer (linear.py:66-73) is the outlier-aware one: an across-block statistics pass, the pack, and a
    #: host check that no activation outlier was found before the same GEMM can run.
    fast_activations = False

    def _forward_fast(self, inputs):
        from .formats import _get_format_params
        pw = self.packed_weight()
        a = self.mx_specs["a_elem_format"]
        if pw is None or a is None or not _tc.mx4_weight_ok(pw):
            return None
        fp = tuple(_get_format_params(a)[:2])
        if fp not in ((2, 3), (4, 5)) or inputs.shape[-1] != pw.shape[1] or not inputs.is_cuda:
            return None
        x2 = inputs.reshape(-1, inputs.shape[-1])
        xq = pack_mx(x2, self.mx_specs["scale_bits"] or 8, a, axes=[-1], block_size=32,
                     flush_fp32_subnorms=self.mx_specs["mx_flush_fp32_subnorms"])
        try:
            if fp == (2, 3) and _tc.mx4_pair_ok(pw, x2.shape[0]):
                y = _tc.packed_linear_mx4(xq, pw, bias=self.bias, out_dtype=inputs.dtype)
            elif fp == (4, 5):
                y = _tc.packed_linear_mx8(xq, pw, bias=self.bias, out_dtype=inputs.dtype)
            else:
                return None
        except NotImplementedError:          # a weight the block-scaled kernels cannot take: the default route decides
            return None
        return y.reshape(*inputs.shape[:-1], pw.shape[0]).to(inputs.dtype)

    def forward(self, inputs):
        if self.mx_none:
            return super().forward(inputs)
        with torch.no_grad():
            if self.fast_activations:
                y = self._forward_fast(inputs)
                if y is not None:
                    return y
            return linear_forward(inputs, self.weight, self.bias, self.mx_specs, packed_weight=self.packed_weight())
