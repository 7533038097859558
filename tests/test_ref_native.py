"""Second oracle: the REFERENCE's own native extension (oracle/_ref, built from the sources under
/root/reference by oracle/build_ref.py).  CPU entry points pin the oracle's "spec" arithmetic on the
CPU tier; on the GPU box the reference's CUDA kernels are compared with ours on the same tensors."""
import os
import sys

import pytest
import torch

import golden_core as G
import msq_oracle as O
from util import assert_exact

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import build_ref  # noqa: E402

REF = build_ref.load_ref()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (python oracle/build_ref.py)")


@needs_ref
@pytest.mark.parametrize("fmt", ["fp8_e4m3", "fp8_e5m2", "fp6_e3m2", "fp6_e2m3", "fp4_e2m1", "int8", "int4", "int2", "bfloat16"])
@pytest.mark.parametrize("rmode", [0, 1, 2])
def test_elemwise_cpp_vs_oracle(fmt, rmode):
    rnd = ["nearest", "floor", "even"][rmode]
    eb, mb, _, mx, _ = O.format_params(fmt)
    g = torch.Generator().manual_seed(1)
    x = torch.cat([torch.randn(50000, generator=g) * 4, torch.from_numpy(G.all_encodings(5, 7))[:4000]])
    x = x[torch.isfinite(x)].contiguous()
    for sat in (False, True):
        y = REF.quantize_elemwise_func_cpp(x, mb, eb, mx, rmode, sat, True)
        assert_exact(O.qelem_native(x, mb, eb, mx, rnd, sat, True), y, "%s %s sat=%s" % (fmt, rnd, sat))
        # the float formulation (what the reference's torch path computes) agrees on this sample too
        assert_exact(O.qelem(x, mb, eb, mx, rnd, sat, True), y)


@needs_ref
@pytest.mark.parametrize("fmt,bs,sb", [("fp4_e2m1", 32, 8), ("fp6_e3m2", 32, 8), ("int8", 10, 8), ("int2", 8, 5),
                                       ("fp8_e4m3", 64, 8), ("int4", 16, 4)])
@pytest.mark.parametrize("rmode", [0, 1, 2])
def test_mx_cpp_vs_oracle_spec_mode(fmt, bs, sb, rmode):
    rnd = ["nearest", "floor", "even"][rmode]
    eb, mb, _, mx, _ = O.format_params(fmt)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(64, 40 * bs, generator=g) * torch.logspace(-8, 4, 64).unsqueeze(1)
    xb = x.reshape(64, 40, bs).contiguous()
    maxv = xb.abs().max(dim=2, keepdim=True).values
    y = REF.quantize_mx_func_cpp(xb, sb, eb, mb, mx, maxv, 2, False, rmode).reshape(x.shape)
    assert_exact(O.quantize_mx(x, sb, fmt, 1, bs, rnd, mode="spec"), y, "%s bs%d sb%d %s" % (fmt, bs, sb, rnd))


@needs_ref
def test_golden_tables_through_reference_cpp():
    x = torch.tensor(G.MX_HW_X).reshape(7, 1, 10).contiguous()
    y = REF.quantize_mx_func_cpp(x, 8, 0, 8, 1.984375, x.abs().max(dim=2, keepdim=True).values, 2, False, 0)
    assert_exact(y.reshape(7, 10), torch.tensor(G.MX_HW_Y))


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("fmt,bs,axis", [("fp4_e2m1", 32, 1), ("fp6_e2m3", 16, 1), ("int8", 64, 1), ("fp8_e4m3", 16, 0),
                                         ("int4", 32, 0), ("fp4_e2m1", 10, 1)])
def test_reference_cuda_kernels_vs_ours(fmt, bs, axis):
    """third oracle: the reference's own CUDA kernels, run on this GPU"""
    import msq_b200
    eb, mb, _, mx, _ = O.format_params(fmt)
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(512, 640, generator=g) * torch.logspace(-6, 3, 512).unsqueeze(1)).cuda()
    for rmode, rnd in enumerate(["nearest", "floor", "even"]):
        ref = REF.quantize_mx_by_tile_func_cuda(x, 8, eb, mb, mx, bs, axis, False, rmode)
        ours = msq_b200._quantize_mx(x, 8, fmt, axes=[axis], block_size=bs, round=rnd, custom_cuda=True)
        assert_exact(ours, ref, "%s %s" % (fmt, rnd))
        ref = REF.quantize_elemwise_func_cuda(x, mb, eb, mx, rmode, True, True)
        ours = msq_b200._quantize_elemwise_core(x, mb, eb, mx, rnd, True, True, custom_cuda=True)
        assert_exact(ours, ref, "elemwise %s %s" % (fmt, rnd))
