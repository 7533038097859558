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


@needs_ref
@pytest.mark.gpu
def test_custom_extensions_funcs_shim_vs_reference_module():
    """the seven names of the reference's pybind module (cpp/funcs.cpp:218-226) behind `custom_extensions.funcs`,
    each called exactly like mx_ops.py:371-422 / elemwise_ops.py:120-129 call it, against the reference's own
    compiled module on the same tensors"""
    from msq_b200.mx import custom_extensions as ce
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(6, 96, 320, generator=g) * torch.logspace(-5, 3, 96).reshape(1, 96, 1)).cuda()
    for fmt, tile in (("fp4_e2m1", 32), ("fp6_e3m2", 16), ("int8", 10), ("fp8_e4m3", 64)):
        eb, mb, _, mx, _ = O.format_params(fmt)
        for rmode in (0, 1, 2):
            for axis in (0, 1, 2):
                ref = REF.quantize_mx_by_tile_func_cuda(x, 8, eb, mb, mx, tile, axis, False, rmode)
                assert_exact(ce.funcs.quantize_mx_by_tile_func_cuda(x, 8, eb, mb, mx, tile, axis, False, rmode), ref,
                             "by_tile %s axis %d rmode %d" % (fmt, axis, rmode))
                mv = x.abs().max(dim=axis, keepdim=True).values.contiguous()
                ref = REF.quantize_mx_func_cuda(x, 8, eb, mb, mx, mv, axis, True, rmode)
                assert_exact(ce.funcs.quantize_mx_func_cuda(x, 8, eb, mb, mx, mv, axis, True, rmode), ref,
                             "mx_func %s axis %d rmode %d" % (fmt, axis, rmode))
            ref = REF.quantize_elemwise_func_cuda(x, mb, eb, mx, rmode, False, True)
            assert_exact(ce.funcs.quantize_elemwise_func_cuda(x, mb, eb, mx, rmode, False, True), ref)
            xh = x.half()
            ref = REF.quantize_elemwise_func_cuda(xh, mb, eb, mx, rmode, True, False)
            assert_exact(ce.funcs.quantize_elemwise_func_cuda(xh, mb, eb, mx, rmode, True, False).float(), ref.float())
    # host-tensor names: same results as the reference's CPU loops
    xc = x[:2].cpu().contiguous()
    eb, mb, _, mx, _ = O.format_params("fp6_e2m3")
    mv = xc.abs().max(dim=2, keepdim=True).values
    assert_exact(ce.funcs.quantize_mx_func_cpp(xc, 8, eb, mb, mx, mv, 2, False, 0),
                 REF.quantize_mx_func_cpp(xc, 8, eb, mb, mx, mv, 2, False, 0))
    assert_exact(ce.funcs.quantize_elemwise_func_cpp(xc, mb, eb, mx, 2, True, True),
                 REF.quantize_elemwise_func_cpp(xc, mb, eb, mx, 2, True, True))
    # reductions: bit-identical sums (same association order), maxima with the -65504 floor
    # (the reference's kernel is only correct when inner divides its 1024-thread block: at inner = 320 it
    #  returns zeros for most rows, so the comparison is on the sizes it handles)
    for inner in (32, 64, 128, 256, 512, 1024):
        a = (torch.randn(37, 5, inner, generator=g) * 100).cuda()
        assert_exact(ce.funcs.reduce_sum_inner_dim(a), REF.reduce_sum_inner_dim(a), "sum %d" % inner)
        assert_exact(ce.funcs.reduce_max_inner_dim(a), REF.reduce_max_inner_dim(a), "max %d" % inner)
    a320 = (torch.randn(11, 320, generator=g) * 100).cuda()
    assert float((ce.funcs.reduce_sum_inner_dim(a320) - a320.double().sum(-1).float()).abs().max()) < 0.05
    assert_exact(ce.funcs.reduce_max_inner_dim(a320), a320.max(-1).values)
    a = torch.full((3, 64), -1.0e5).cuda()
    assert_exact(ce.funcs.reduce_max_inner_dim(a), REF.reduce_max_inner_dim(a))
    with pytest.raises(RuntimeError):
        ce.funcs.reduce_sum_inner_dim(torch.zeros(4, 48).cuda())
    with pytest.raises(RuntimeError):
        ce.funcs.quantize_elemwise_func_cuda(x, 25, 8, 1.0)
    with pytest.raises(RuntimeError):
        ce.funcs.quantize_mx_by_tile_func_cuda(x.transpose(0, 1), 8, 2, 3, 6.0, 32, 0)


def test_custom_extensions_funcs_names():
    """CPU tier: the shim exposes exactly the reference's seven names and refuses host tensors on the *_cuda names"""
    from msq_b200.mx import custom_extensions as ce
    names = ["quantize_mx_func_cpp", "quantize_elemwise_func_cpp", "quantize_mx_func_cuda", "quantize_mx_by_tile_func_cuda",
             "quantize_elemwise_func_cuda", "reduce_sum_inner_dim", "reduce_max_inner_dim"]
    assert sorted(n for n in dir(ce.funcs) if not n.startswith("_")) == sorted(names)
    with pytest.raises(RuntimeError):
        ce.funcs.quantize_mx_by_tile_func_cuda(torch.zeros(4, 32), 8, 2, 3, 6.0, 32, 1)
    with pytest.raises(RuntimeError):
        ce.funcs.reduce_max_inner_dim(torch.zeros(4, 32))
