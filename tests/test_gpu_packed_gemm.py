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
