"""GPU tier: packed-weight tcgen05 linear vs the fp32 matmul on the dequantized weight
(what number_system/mx/linear.py:91 computes).  Products are exact, only fp32 summation order
differs: tolerance = 4*K*2^-24 * (|x| @ |w|^T)."""
import pytest
import torch

import msq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import msq_b200
    return msq_b200


def check(y, x, w):
    ref = x.double() @ w.double().t()
    bound = 4 * x.shape[-1] * 2.0 ** -24 * (x.double().abs() @ w.double().abs().t()) + 1e-30
    assert bool(((y.double() - ref).abs() <= bound).all()), float(((y.double() - ref).abs() / bound).max())


@pytest.mark.timeout(300)
@pytest.mark.parametrize("Mx,N,K", [(1, 256, 128), (16, 128, 64), (48, 384, 256), (100, 200, 96), (256, 512, 1024),
                                    (300, 4096, 4096), (8, 11008, 4096)])
@pytest.mark.parametrize("fmt", ["fp4_e2m1", "int4", "fp6_e3m2", "int2"])
def test_packed_linear_outlier(M, Mx, N, K, fmt):
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(N + K)
    W = torch.randn(N, K, generator=g) * 0.02
    W.view(-1)[torch.randperm(W.numel(), generator=g)[: W.numel() // 500]] *= 20
    x = torch.randn(Mx, K, generator=g).to(torch.bfloat16)
    pw = M.pack_mx_outlier(W.cuda(), 8, 8, fmt, fmt, 5, [1], 32, variant=1)
    wd = pw.dequantize()
    y = _tc.packed_linear(x.cuda(), pw)
    check(y.cpu(), x.float(), wd.cpu())


@pytest.mark.timeout(300)
@pytest.mark.parametrize("dt", [torch.float16, torch.float32])
def test_packed_linear_plain_and_split_activations(M, dt):
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(9)
    W = torch.randn(640, 512, generator=g) * 0.05
    x = (torch.randn(3, 20, 512, generator=g) * 3).to(dt)
    b = torch.randn(640, generator=g)
    pw = M.pack_mx(W.cuda(), 8, "fp4_e2m1", axes=[1], block_size=32)
    y = _tc.packed_linear(x.cuda(), pw, bias=b.cuda())
    assert y.shape == (3, 20, 640)
    check((y.cpu() - b).reshape(-1, 640), x.float().reshape(-1, 512), pw.dequantize().cpu())


@pytest.mark.timeout(300)
def test_mx_linear_matches_reference_fixtures(M):
    """mx.linear / MXLinear against outputs of the reference's own mx.linear (tests/golden/linear.npz)"""
    from util import load_json, load_npz
    z = load_npz("linear.npz")
    for n, spec in enumerate(load_json("linear_cases.json")):
        specs = M.finalize_mx_specs(dict(spec))
        x, w, b, ref = z["x_%d" % n].cuda(), z["w_%d" % n].cuda(), z["b_%d" % n].cuda(), z["y_%d" % n]
        y = M.linear(x, w, b, specs)
        # fp32 accumulation order is the only difference (bfloat-rounded outputs may flip one ulp of bf16)
        tol = 1e-5 if not spec.get("bfloat") else 2.0 ** -8
        assert float((y.cpu() - ref).abs().max()) <= tol * float(ref.abs().max()) + 1e-6, n
        lin = M.MXLinear(256, 96, bias=True, mx_specs=dict(spec)).cuda()
        lin.weight.data.copy_(w)
        lin.bias.data.copy_(b)
        y2 = lin(x)
        assert torch.equal(y2, y)
        assert lin.packed_weight() is lin.packed_weight()          # cached
    assert torch.equal(M.linear(x, w, b, None), torch.nn.functional.linear(x, w, b))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("dt,T", [(torch.bfloat16, 2), (torch.bfloat16, 3), (torch.float16, 3)])
def test_split_a16_terms_reproduce_the_activations(M, dt, T):
    """msq_split_a16_e4m3: sum_t a_t * 2^e_t == x EXACTLY while an element is within 2^12 (bf16, two terms) / 2^15 (bf16,
    three) / 2^11 (fp16, three) of its block's maximum; below that the truncation error is < 2^-19 / 2^-22 of the block
    maximum and never increases a magnitude.  Inf / NaN poison their block."""
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(3)
    Mx, K = 37, 512
    spread = {(torch.bfloat16, 2): 12, (torch.bfloat16, 3): 15, (torch.float16, 3): 11}[(dt, T)]
    floor = 2.0 ** -19 if T == 2 else 2.0 ** -22
    x = torch.randn(Mx, K, generator=g) * torch.exp2(torch.randint(-spread // 2, spread // 2 + 1, (Mx, K), generator=g).float())
    x[0, :32] = 0.0                                        # an all-zero block
    x[1, 5] = 0.0
    x = x.to(dt)
    x[2, 40:48] = torch.tensor([2.0 ** -i for i in range(8)], dtype=dt) * x[2, 40:48].abs().max() + 0    # exact powers of two

    def rebuild(t):
        sa = _tc.split_a16(t.cuda(), terms=T)
        r, c = t.shape
        assert sa.shape == (T * r, c) and sa.terms == T
        terms = sa.codes.view(torch.float8_e4m3fn).float().view(T, r, c // 32, 32) * torch.exp2(sa.e_in.float()).view(T, r, c // 32, 1)
        return terms.sum(0).view(r, c), sa

    rec, _ = rebuild(x)
    xf = x.float().cuda()
    blockmax = xf.abs().view(Mx, K // 32, 32).amax(-1, keepdim=True).expand(Mx, K // 32, 32).reshape(Mx, K)
    near = xf.abs() * 2.0 ** spread >= blockmax
    assert torch.equal(rec[near], xf[near])
    assert bool(((rec - xf).abs() <= blockmax * floor).all())
    # wide dynamic range inside a block: truncation at `floor` of the block maximum, never more, never away from zero
    y = (torch.randn(8, 256, generator=g) * torch.exp2(torch.randint(-30, 1, (8, 256), generator=g).float())).to(dt)
    rec, _ = rebuild(y)
    yf = y.float().cuda()
    bm = yf.abs().view(8, 8, 32).amax(-1, keepdim=True).expand(8, 8, 32).reshape(8, 256)
    assert bool(((rec - yf).abs() <= bm * floor).all()) and bool((rec.abs() <= yf.abs()).all())
    z = torch.ones(2, 64, dtype=dt)
    z[1, 3] = float("inf")
    _, sz = rebuild(z)
    assert all(int(sz.e_in[t * 2 + 1, 0]) == -128 for t in range(T)) and int(sz.e_in[1, 1]) != -128


@pytest.mark.timeout(300)
@pytest.mark.parametrize("Mx,N,K", [(1, 256, 128), (8, 512, 1024), (16, 4096, 4096), (48, 384, 256), (200, 1024, 2048), (500, 256, 512)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("outliers", [False, True])
def test_w4a16_on_the_block_scaled_path(M, monkeypatch, Mx, N, K, dt, outliers):
    """16-bit activations x K-blocked FP4 weight: packed_linear splits the activations into three block-scaled E4M3 terms
    and runs the block-scaled kernels on the packed codes (no bf16 copy of the weight, no producer dequantization).
    Same tolerance as the bf16 kernel: the split is exact for these inputs, products are exact, fp32 accumulation."""
    from msq_b200 import _tc
    g = torch.Generator().manual_seed(N + K + Mx)
    W = torch.randn(N, K, generator=g) * 0.02
    if outliers:
        W.view(-1)[torch.randperm(W.numel(), generator=g)[: W.numel() // 2000]] *= 20
        pw = M.pack_mx_outlier(W.cuda(), 8, 8, "fp4_e2m1", "fp4_e2m1", 5, [1], 32, variant=1)
    else:
        pw = M.pack_mx(W.cuda(), 8, "fp4_e2m1", axes=[1], block_size=32)
    x = (torch.randn(Mx, K, generator=g) * torch.logspace(-2, 1, Mx).unsqueeze(1)).to(dt)
    calls = []
    real = _tc.split_a16
    monkeypatch.setattr(_tc, "split_a16", lambda t: (calls.append(1), real(t))[1])
    y = _tc.packed_linear(x.cuda(), pw)
    assert calls, "packed_linear did not take the split-activation path"
    check(y.cpu(), x.float(), pw.dequantize().cpu())
    monkeypatch.setattr(_tc, "SPLIT_A16", False)           # the producer-dequantizing bf16 kernel agrees
    y0 = _tc.packed_linear(x.cuda(), pw)
    check(y0.cpu(), x.float(), pw.dequantize().cpu())
