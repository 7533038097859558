"""GPU tier: W4A4 on tcgen05.mma kind::mxf4.block_scale (msq_gemm_mxfp4_tma / packed_linear_mx4): both operands packed
FP4 codes.  Reference = fp64 matmul of the DEQUANTIZED operands (number_system/mx/linear.py:91 multiplies the two
fake-quantized tensors); products and power-of-two block scales are exact, only fp32 summation order differs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import msq_b200
    return msq_b200


def check(y, x, w):
    ref = x.double() @ w.double().t()
    bound = 4 * x.shape[-1] * 2.0 ** -24 * (x.double().abs() @ w.double().abs().t()) + 1e-30
    bad = ((y.double() - ref).abs() > bound)
    assert not bool(bad.any()), (int(bad.sum()), float(((y.double() - ref).abs() / bound).max()))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("Mx,N,K", [(1, 128, 256), (16, 128, 256), (33, 200, 512), (100, 1000, 768), (256, 256, 256),
                                    (300, 384, 1024), (513, 640, 2048), (1100, 384, 256), (2048, 1024, 512),
                                    (700, 4096, 4096), (64, 11008, 4096)])
def test_w4a4_linear(M, Mx, N, K):
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(N + K + Mx)
    W = torch.randn(N, K, generator=g) * 0.02 * torch.logspace(-1, 1, N).unsqueeze(1)
    x = torch.randn(Mx, K, generator=g) * torch.logspace(-2, 1, Mx).unsqueeze(1)
    pw = M.pack_mx(W.cuda(), 8, "fp4_e2m1", axes=[1], block_size=32)
    xq = M.pack_mx(x.cuda(), 8, "fp4_e2m1", axes=[-1], block_size=32)
    assert _tc.mx4_pair_ok(pw)
    y = _tc.packed_linear_mx4(xq, pw)
    assert y.shape == (Mx, N)
    check(y.cpu(), xq.dequantize().cpu(), pw.dequantize().cpu())
    # the same operands through the FP8-activation kernel (an FP4 value is an E4M3 value): same sums, other order
    b = torch.randn(N, generator=g).cuda()
    yb = _tc.packed_linear_mx4(xq, pw, bias=b)
    assert torch.allclose(yb, y + b)


def test_w4a4_rejects(M):
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(5)
    W = torch.randn(128, 384, generator=g) * 0.02           # K % 256 != 0
    pw = M.pack_mx(W.cuda(), 8, "fp4_e2m1", axes=[1], block_size=32)
    xq = M.pack_mx(torch.randn(4, 384, generator=g).cuda(), 8, "fp4_e2m1", axes=[-1], block_size=32)
    assert not _tc.mx4_pair_ok(pw)
    with pytest.raises(NotImplementedError):
        _tc.packed_linear_mx4(xq, pw)
    W = torch.randn(128, 512, generator=g) * 0.02
    pw = M.pack_mx(W.cuda(), 8, "fp4_e2m1", axes=[1], block_size=32)
    x8 = M.pack_mx(torch.randn(4, 512, generator=g).cuda(), 8, "fp8_e4m3", axes=[-1], block_size=32)
    with pytest.raises(NotImplementedError):                # FP8 activations belong to packed_linear_mx8
        _tc.packed_linear_mx4(x8, pw)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("Mx,N,K", [(16, 256, 1024), (300, 384, 2048), (400, 512, 4096)])
def test_w4a4_weight_with_outliers(M, Mx, N, K):
    """FP4 outliers of the weight (own shared exponent) = sparse correction behind the kind::mxf4 GEMM
    (msq_sparse_outlier_correct_fp4); beyond the outliers x tokens budget the pairing is refused"""
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(N + K + Mx)
    W = torch.randn(N, K, generator=g) * 0.02
    W.view(-1)[torch.randperm(W.numel(), generator=g)[: max(4, W.numel() // 400)]] *= 20
    x = torch.randn(Mx, K, generator=g) * torch.logspace(-2, 1, Mx).unsqueeze(1)
    pw = M.pack_mx_outlier(W.cuda(), 8, 8, "fp4_e2m1", "fp4_e2m1", 5, [1], 32, variant=1)
    assert bool((pw.bitmap != 0).any())
    xq = M.pack_mx(x.cuda(), 8, "fp4_e2m1", axes=[-1], block_size=32)
    assert _tc.mx4_pair_ok(pw, Mx)
    y = _tc.packed_linear_mx4(xq, pw)
    check(y.cpu(), xq.dequantize().cpu(), pw.dequantize().cpu())
    old = _tc.MX8_TMA_SPARSE_BUDGET
    _tc.MX8_TMA_SPARSE_BUDGET = 1.0
    try:
        assert not _tc.mx4_pair_ok(pw, Mx)
        with pytest.raises(NotImplementedError):
            _tc.packed_linear_mx4(xq, pw)
    finally:
        _tc.MX8_TMA_SPARSE_BUDGET = old


@pytest.mark.timeout(300)
def test_mx_linear_w4a4_route(M):
    """mx.linear with a_elem_format = w_elem_format = fp4_e2m1 takes the kind::mxf4 kernel and gives the result of the
    kind::mxf8f6f4 / dequantized routes (linear.py:66-91: F.linear of the fake-quantized operands)"""
    from msq_b200 import _tc
    from msq_b200 import mx
    specs = mx.specs.finalize_mx_specs({"w_elem_format": "fp4_e2m1", "a_elem_format": "fp4_e2m1", "block_size": 32,
                                        "scale_bits": 8, "bfloat": 0, "custom_cuda": False})
    g = torch.Generator().manual_seed(11)
    W = (torch.randn(384, 1024, generator=g) * 0.02).cuda()
    x = torch.randn(48, 1024, generator=g).cuda()
    b = torch.randn(384, generator=g).cuda()
    calls = []
    orig = _tc.packed_linear_mx4

    def spy(*a, **k):
        calls.append(1)
        return orig(*a, **k)
    _tc.packed_linear_mx4 = spy
    try:
        y4 = mx.linear(x, W, b, mx_specs=specs)
    finally:
        _tc.packed_linear_mx4 = orig
    assert calls, "the W4A4 pairing did not reach the kind::mxf4 kernel"
    _tc.MX4_TMA = False
    try:
        y8 = mx.linear(x, W, b, mx_specs=specs)
    finally:
        _tc.MX4_TMA = True
    assert torch.allclose(y4, y8, rtol=1e-5, atol=1e-6 * float(y8.abs().max()))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("a_fmt", ["fp4_e2m1", "fp8_e4m3"])
def test_mxlinear_serving_mode(M, a_fmt):
    """MXLinear.from_packed(..., fast_activations=True): one-launch plain-MX activation quantizer + block-scaled GEMM;
    result = the fp64 matmul of the dequantized operands (+ bias) within the summation-order bound"""
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(21)
    W = (torch.randn(384, 1024, generator=g) * 0.02).cuda()
    b = torch.randn(384, generator=g).cuda()
    x = (torch.randn(3, 50, 1024, generator=g) * 2).cuda()
    pw = M.pack_mx(W, 8, "fp4_e2m1", axes=[1], block_size=32)
    lin = M.MXLinear.from_packed(pw, bias=b, mx_specs={"w_elem_format": "fp4_e2m1", "a_elem_format": a_fmt, "block_size": 32,
                                                        "scale_bits": 8}, fast_activations=True)
    calls, saved = [], {}
    for name in ("packed_linear_mx4", "packed_linear_mx8"):
        saved[name] = getattr(_tc, name)

        def spy(*a, _o=saved[name], _n=name, **k):
            calls.append(_n)
            return _o(*a, **k)
        setattr(_tc, name, spy)
    try:
        y = lin(x)
    finally:
        for name, fn in saved.items():
            setattr(_tc, name, fn)
    assert y.shape == (3, 50, 384)
    assert calls and calls[0] == ("packed_linear_mx4" if a_fmt == "fp4_e2m1" else "packed_linear_mx8")
    xq = M.pack_mx(x.reshape(-1, 1024), 8, a_fmt, axes=[-1], block_size=32)
    check((y.reshape(-1, 384) - b).cpu(), xq.dequantize().cpu(), pw.dequantize().cpu())
    for bf in (torch.bfloat16, torch.float16):           # 16-bit activations are read as they are
        yb = lin(x.to(bf))
        assert yb.dtype == bf and torch.isfinite(yb).all()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("Mx,N,K", [(48, 512, 4096), (64, 4096, 4096), (33, 384, 8192), (64, 1000, 8192)])
def test_split_k_between_decode_and_compute_bound_sizes(M, Mx, N, K):
    """33..64 tokens: the plan cuts every tile into K slices that add into a zero-filled C by atomics (both pairings);
    result within the summation-order bound, and a second call into the SAME output buffer starts from zero again"""
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(N + K + Mx)
    W = torch.randn(N, K, generator=g) * 0.02
    x = torch.randn(Mx, K, generator=g) * torch.logspace(-1, 1, Mx).unsqueeze(1)
    pw = M.pack_mx(W.cuda(), 8, "fp4_e2m1", axes=[1], block_size=32)
    for fmt, fn, w4a4 in (("fp8_e4m3", _tc.packed_linear_mx8, 0), ("fp4_e2m1", _tc.packed_linear_mx4, 1)):
        ks = _tc.N.lib().msq_gemm_mx_tma_k_slices(Mx, N, K, w4a4)
        if K // (256 if w4a4 else 128) >= 32:
            assert ks > 1, (fmt, ks)
        xq = M.pack_mx(x.cuda(), 8, fmt, axes=[-1], block_size=32)
        y = fn(xq, pw)
        check(y.cpu(), xq.dequantize().cpu(), pw.dequantize().cpu())
        y2 = fn(xq, pw)
        check(y2.cpu(), xq.dequantize().cpu(), pw.dequantize().cpu())


@pytest.mark.timeout(300)
@pytest.mark.parametrize("Mx,N,K", [(5, 128, 256), (48, 512, 8192), (300, 640, 1024), (1000, 1000, 2048)])
def test_w4a4_bf16_result(M, Mx, N, K):
    """msq_gemm_mxfp4_tma_bf16: the accumulator leaves the epilogue as bf16 -- the fp32 result rounded once (identical bits
    where both calls use the same plan; within half a bf16 ulp where the fp32 call splits K)"""
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(N + K + Mx)
    W = torch.randn(N, K, generator=g) * 0.02
    x = torch.randn(Mx, K, generator=g)
    pw = M.pack_mx(W.cuda(), 8, "fp4_e2m1", axes=[1], block_size=32)
    xq = M.pack_mx(x.cuda(), 8, "fp4_e2m1", axes=[-1], block_size=32)
    y32 = _tc.packed_linear_mx4(xq, pw)
    y16 = _tc.packed_linear_mx4(xq, pw, out_dtype=torch.bfloat16)
    assert y16.dtype == torch.bfloat16 and y16.shape == y32.shape
    if _tc.N.lib().msq_gemm_mx_tma_k_slices(Mx, N, K, 1) == 1:
        assert torch.equal(y16, y32.to(torch.bfloat16))
    else:
        assert float((y16.float() - y32).abs().max()) <= 2.0 ** -8 * float(y32.abs().max())
