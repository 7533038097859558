"""GPU tier: native block-scaled tensor-core linear (tcgen05.mma kind::mxf8f6f4.block_scale).
Reference = fp64 matmul of the DEQUANTIZED operands (number_system/mx/linear.py:91 multiplies fake-quantized
tensors).  Element products and power-of-two block scales are exact; only fp32 summation order differs."""
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
@pytest.mark.parametrize("Mx,N,K", [(16, 128, 128), (1, 256, 256), (48, 384, 512), (100, 200, 160), (256, 512, 1024),
                                    (300, 1024, 4096), (8, 4096, 4096)])
@pytest.mark.parametrize("outliers", [False, True])
def test_blockscaled_linear(M, Mx, N, K, outliers):
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(N + K + Mx)
    W = torch.randn(N, K, generator=g) * 0.02
    if outliers:       # make the across-block detector fire: a few entries 20x larger
        W.view(-1)[torch.randperm(W.numel(), generator=g)[: max(4, W.numel() // 400)]] *= 20
    x = torch.randn(Mx, K, generator=g) * torch.logspace(-2, 1, Mx).unsqueeze(1)
    pw = M.pack_mx_outlier(W.cuda(), 8, 8, "fp4_e2m1", "fp4_e2m1", 5, [1], 32, variant=1)
    if outliers and K >= 1024:      # the 5-sigma across-block detector needs enough blocks per row to fire
        assert bool((pw.bitmap != 0).any())
    assert _tc.mx8_weight_ok(pw)
    xq = M.pack_mx(x.cuda(), 8, "fp8_e4m3", axes=[-1], block_size=32)
    y = _tc.packed_linear_mx8(xq, pw)
    check(y.cpu(), xq.dequantize().cpu(), pw.dequantize().cpu())


def test_blockscaled_rejects_unrepresentable_outliers(M):
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(1)
    W = torch.randn(16, 4096, generator=g) * 0.02
    W[3, 5] = 50.0                  # ~1000x the inlier block scale: q * 2^(e_out - 2 e_in) is not an E4M3 value
    pw = M.pack_mx_outlier(W.cuda(), 8, 8, "fp4_e2m1", "fp4_e2m1", 5, [1], 32, variant=1)
    assert bool((pw.bitmap != 0).any())
    assert not _tc.mx8_weight_ok(pw)
    x = torch.randn(4, 4096, generator=g)
    xq = M.pack_mx(x.cuda(), 8, "fp8_e4m3", axes=[-1], block_size=32)
    with pytest.raises(NotImplementedError):
        _tc.packed_linear_mx8(xq, pw)
    # ... and the bf16 kernel still serves it exactly
    y = _tc.packed_linear(xq.dequantize().to(torch.bfloat16), pw)
    check(y.cpu(), xq.dequantize().to(torch.bfloat16).float().cpu(), pw.dequantize().cpu())
