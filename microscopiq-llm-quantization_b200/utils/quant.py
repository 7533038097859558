"""MicroScopiQ quantizers on B200 -- mirror of the reference's utils/quant.py.

Public names, positional order, defaults and return arity follow utils/quant.py:23-454.
All arithmetic runs in libmsq_b200.so on the GPU (no CPU fallback).  CUDA tensors are quantized in place on
the device; quantize_mx_outlier_v1 also accepts a HOST tensor, as the reference's drivers pass it
(llm/llama.py:94-121), and then streams it through the device in chunks (csrc/msq_host.cu).
"""
import torch
import torch.nn as nn

from .. import _ops
from ..mx.formats import (ElemFormat, FP32_EXPONENT_BIAS, FP32_MIN_NORMAL, RoundingMode,  # noqa: F401
                          _get_format_params)
from ..mx.specs import finalize_mx_specs


def _mx_outlier(A, isb, osb, ifmt, ofmt, shared_exp_method, std_dev, axes, block_size, round, flush, want):
    assert isb > 0 and osb > 0
    if shared_exp_method != "max":
        raise NotImplementedError("shared_exp_method=%r for the outlier quantizer" % (shared_exp_method,))
    if round == "dither":
        raise NotImplementedError("dither rounding is not part of the B200 contract")
    axis = _ops._single_axis(axes, A.ndim)
    return _ops.mx_outlier(A, isb, osb, ifmt, ofmt, std_dev, axis, block_size, round, flush, variant=0, want=want)


def quantize_mx_outlier_v1(A, inlier_scale_bits, outlier_scale_bits, inlier_elem_format, outlier_elem_format,
                           shared_exp_method="max", std_dev=2, axes=None, block_size=0, round="nearest",
                           flush_fp32_subnorms=False, custom_cuda=False):
    """utils/quant.py:147-266 -- MX inliers + two-level-scaled outliers, dequantized."""
    if inlier_elem_format is None:
        return A
    if isinstance(A, torch.Tensor) and not A.is_cuda:
        return quantize_mx_outlier_host(A, inlier_scale_bits, outlier_scale_bits, inlier_elem_format,
                                        outlier_elem_format, shared_exp_method, std_dev, axes, block_size, round,
                                        flush_fp32_subnorms)
    return _mx_outlier(A, inlier_scale_bits, outlier_scale_bits, inlier_elem_format, outlier_elem_format,
                       shared_exp_method, std_dev, axes, block_size, round, flush_fp32_subnorms, ("deq",)).deq


def quantize_mx_outlier_host(A, inlier_scale_bits, outlier_scale_bits, inlier_elem_format, outlier_elem_format,
                             shared_exp_method="max", std_dev=2, axes=None, block_size=0, round="nearest",
                             flush_fp32_subnorms=False, out=None, sync=True):
    """quantize_mx_outlier_v1 for a host-resident tensor: same arguments, host tensor in, host tensor out
    (pinned unless `out` is given).  sync=False queues the work and returns immediately; call host_sync()
    before reading the result (lets consecutive tensors share one upload/compute/download pipeline)."""
    assert inlier_scale_bits > 0 and outlier_scale_bits > 0
    if shared_exp_method != "max":
        raise NotImplementedError("shared_exp_method=%r for the outlier quantizer" % (shared_exp_method,))
    if round == "dither":
        raise NotImplementedError("dither rounding is not part of the B200 contract")
    axis = _ops._single_axis(axes, A.ndim)
    return _ops.mx_outlier_host(A, inlier_scale_bits, outlier_scale_bits, inlier_elem_format, outlier_elem_format,
                                std_dev, axis, block_size, round, flush_fp32_subnorms, out=out, sync=sync)


host_sync = _ops.host_sync


def quantize_mx_outlier_hessian(A, inlier_scale_bits, outlier_scale_bits, inlier_elem_format, outlier_elem_format,
                                shared_exp_method="max", std_dev=2, axes=None, block_size=0, round="nearest",
                                flush_fp32_subnorms=False, prune_inliers=False, custom_cuda=False):
    """utils/quant.py:23-146 -- as v1, plus `num_outliers` (int8): the outlier counts of
    blocks 0, bs, 2bs, ... only (quant.py:66).  Like the reference, returns a bare tensor
    when inlier_elem_format is None and ignores prune_inliers."""
    if inlier_elem_format is None:
        return A
    r = _mx_outlier(A, inlier_scale_bits, outlier_scale_bits, inlier_elem_format, outlier_elem_format,
                    shared_exp_method, std_dev, axes, block_size, round, flush_fp32_subnorms, ("deq", "blk_cnt"))
    pre, L, post, bs, nblk = r.geom
    # outlier_pos is [..., nblk, bs, ...]; the reference slices its FIRST dim with step
    # block_size and sums over dim -2.  For the layouts the drivers use (2-D input, axes=[0])
    # the first dim is the block index and dim -2 the in-block axis.
    if pre != 1 or A.ndim != 2:
        raise NotImplementedError("num_outliers is defined for 2-D inputs blocked along axis 0 (quant.py:66)")
    cnt = r.blk_cnt.view(nblk, post)[::bs].flatten()
    return r.deq, cnt.to(torch.int8)


def quantize(x, scale, zero, maxq):
    """GPTQ uniform affine fake-quant (utils/quant.py:268-272)."""
    from .. import _uniform
    return _uniform.quantize(x, scale, zero, maxq)


class Quantizer(nn.Module):
    """utils/quant.py:274-390 -- min/max (+ optional MSE grid) uniform quantizer."""

    def __init__(self, shape=1):
        super().__init__()
        self.register_buffer("maxq", torch.tensor(0))
        self.register_buffer("scale", torch.zeros(shape))
        self.register_buffer("zero", torch.zeros(shape))

    def configure(self, bits, perchannel=False, sym=True, mse=False, norm=2.4, grid=100, maxshrink=.8, trits=False):
        self.maxq = torch.tensor(2 ** bits - 1)
        self.perchannel = perchannel
        self.sym = sym
        self.mse = mse
        self.norm = norm
        self.grid = grid
        self.maxshrink = maxshrink
        if trits:
            self.maxq = torch.tensor(-1)

    def find_params(self, x, weight=False):
        from .. import _uniform
        self.maxq = self.maxq.to(x.device)
        self.scale, self.zero = _uniform.find_params(
            x, int(self.maxq), self.perchannel, self.sym, self.mse, self.norm, self.grid, self.maxshrink, weight)

    def quantize(self, x):
        if self.ready():
            return quantize(x, self.scale, self.zero, self.maxq)
        return x

    def enabled(self):
        return self.maxq > 0

    def ready(self):
        return torch.all(self.scale != 0)


class MXQuantizer(nn.Module):
    """utils/quant.py:393-454 -- configuration holder for the MX+outlier quantizer."""

    def __init__(self, shape=1):
        super().__init__()
        self.mx_specs = finalize_mx_specs({"w_elem_format": "int2", "a_elem_format": "fp16", "block_size": 128,
                                           "custom_cuda": False, "quantize_backprop": False})

    def configure(self, inlier_scale_bits, outlier_scale_bits, inlier_elem_format, outlier_elem_format,
                  shared_exp_method="max", std_dev=2, axes=None, block_size=0, round="nearest",
                  flush_fp32_subnorms=False, custom_cuda=False):
        self.inlier_scale_bits = inlier_scale_bits
        self.inlier_elem_format = inlier_elem_format
        self.outlier_scale_bits = outlier_scale_bits
        self.outlier_elem_format = outlier_elem_format
        self.shared_exp_method = shared_exp_method
        self.std_dev = std_dev
        self.axes = axes
        self.block_size = block_size
        self.round = round
        self.flush_fp32_subnorms = flush_fp32_subnorms
        self.custom_cuda = custom_cuda

    def find_params(self, x, weight=False):
        pass

    def quantize(self, x):
        if self.ready():
            return quantize_mx_outlier_v1(x, self.inlier_scale_bits, self.outlier_scale_bits,
                                          self.inlier_elem_format, self.outlier_elem_format, self.shared_exp_method,
                                          self.std_dev, self.axes, self.block_size, self.round,
                                          self.flush_fp32_subnorms, self.custom_cuda)
        return x

    def enabled(self):
        pass

    def ready(self):
        return True


# ---------------------------------------------------------------------------------------
# block view helpers (pure tensor views; utils/quant.py:544-616)
# ---------------------------------------------------------------------------------------
def _reshape_to_blocks(A, axes, block_size):
    if axes is None:
        raise Exception("axes required in order to determine which dimension toapply block size to")
    if block_size == 0:
        raise Exception("block_size == 0 in _reshape_to_blocks")
    axes = sorted(a + A.ndim if a < 0 else a for a in axes)
    for n, a in enumerate(axes):
        axes[n] = a + n                       # every earlier insert shifts later axes by one
        A = A.unsqueeze(axes[n] + 1)
    orig_shape = A.size()
    pad = [0, 0] * A.ndim                     # F.pad order: last dim first
    for a in axes:
        rem = orig_shape[a] % block_size
        if rem:
            pad[2 * (A.ndim - 1 - a) + 1] = block_size - rem
    if any(pad):
        A = torch.nn.functional.pad(A, pad, mode="constant")
    padded_shape = A.size()
    view = list(padded_shape)
    for a in axes:
        if view[a] >= block_size:
            assert view[a] % block_size == 0
            view[a + 1] = block_size
            view[a] //= block_size
        else:
            view[a + 1], view[a] = view[a], 1
    return A.view(view), axes, orig_shape, padded_shape


def _undo_reshape_to_blocks(A, padded_shape, orig_shape, axes):
    A = A.view(padded_shape)
    if list(padded_shape) != list(orig_shape):
        A = A[tuple(slice(0, n) for n in orig_shape)]
    for a in reversed(axes):
        A = A.squeeze(a + 1)
    return A


def _extract_outlier_indices(A, std_dev=1, axes=None):
    """utils/quant.py:460-495 on an already-blocked tensor: 0/1 mask (A.dtype) of elements
    outside mean(|A|) +- std_dev*std(|A|) over `axes` (single axis)."""
    axis = _ops._single_axis(axes, A.ndim)
    r = _ops.mx_outlier(A, 8, 8, "int8", "int8", std_dev, axis, 0, want=("mask",), check=False)
    return r.mask.to(A.dtype)


def _shared_exponents(A, method="max", axes=None, ebits=0):
    """utils/quant.py:498-541: floor(log2(max|A|)) over `axes` (keepdim), exact."""
    from .. import _misc
    return _misc.shared_exponents(A, method, axes, ebits)
