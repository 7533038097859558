"""GPU tier: the activation quantizer of the block-scaled linears (csrc/msq_quant_act.cu, msq_quantize_mx_act) against the
generic plain-MX path (msq_quantize_mx_f32, mode 1 -- itself pinned to the oracle and to the reference's compiled kernels
by test_gpu_mxops.py / test_ref_native.py): codes and shared exponents bit for bit, for every input dtype."""
import pytest
import torch

import msq_oracle as O
from util import assert_exact

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import msq_b200
    return msq_b200


def _inputs(kind, rows, K):
    g = torch.Generator().manual_seed(rows * 7 + K)
    if kind == "gauss":
        return torch.randn(rows, K, generator=g) * torch.logspace(-3, 2, rows).unsqueeze(1)
    if kind == "ties":           # few mantissa bits and their fp32 neighbours: rounding ties of both grids everywhere
        m = torch.randint(-64, 65, (rows, K), generator=g).float() / 16.0
        x = m * torch.exp2(torch.randint(-12, 6, (rows, K), generator=g).float() + torch.randint(-30, 31, (rows, 1), generator=g).float())
        d = torch.randint(-1, 2, (rows, K), generator=g, dtype=torch.int32)
        return torch.where(x != 0, (x.view(torch.int32) + d).view(torch.float32), x)
    if kind == "degenerate":
        x = torch.randn(rows, K, generator=g)
        x[0] = 0
        x[1 % rows, :32] = 0
        x[2 % rows] = torch.where(torch.rand(K, generator=g) < 0.5, -1e-9, 1.0)
        x[3 % rows] *= 1e-42                                   # fp32 subnormals
        x[4 % rows] *= 1e37
        x[5 % rows, 7] = float("nan")
        x[6 % rows, 40 % K] = float("inf")
        x[7 % rows] = -x[7 % rows].abs() * 1e-30
        return x
    raise ValueError(kind)


@pytest.mark.parametrize("fmt", ["fp4_e2m1", "fp8_e4m3"])
@pytest.mark.parametrize("kind", ["gauss", "ties", "degenerate"])
@pytest.mark.parametrize("rows,K", [(1, 32), (3, 96), (16, 4096), (130, 1056), (257, 160)])
def test_act_kernel_equals_generic_path(M, fmt, kind, rows, K):
    from msq_b200 import packed
    x = _inputs(kind, rows, K).cuda()
    for sb, flush in ((8, False), (8, True), (5, False), (4, False)):
        a = M.pack_mx(x, sb, fmt, axes=[-1], block_size=32, flush_fp32_subnorms=flush)
        assert a.path == 7
        packed.ACT_KERNEL = False
        try:
            b = M.pack_mx(x, sb, fmt, axes=[-1], block_size=32, flush_fp32_subnorms=flush)
        finally:
            packed.ACT_KERNEL = True
        assert b.path != 7
        assert torch.equal(a.e_in, b.e_in), (sb, flush)
        assert torch.equal(a.codes, b.codes), (sb, flush, int((a.codes != b.codes).sum()))
        assert int(a.status.item()) == int(b.status.item())
        ok = torch.isfinite(b.dequantize())
        assert torch.equal(a.dequantize()[ok], b.dequantize()[ok])


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("fmt", ["fp4_e2m1", "fp8_e4m3"])
def test_act_kernel_16bit_inputs_and_oracle(M, dt, fmt):
    """bf16 / fp16 activations are read as they are: same planes as quantizing their fp32 copy; values = the oracle's"""
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(5, 7, 256, generator=g) * 3).to(dt)
    x[0, 0, :5] = torch.tensor([0.0, -0.0, 6.0, -6.0, 0.25]).to(dt)
    a = M.pack_mx(x.cuda(), 8, fmt, axes=[-1], block_size=32)
    b = M.pack_mx(x.float().cuda(), 8, fmt, axes=[-1], block_size=32)
    assert a.path == 7 and a.shape == (5, 7, 256)
    assert torch.equal(a.codes, b.codes) and torch.equal(a.e_in, b.e_in)
    y = O.quantize_mx(x.float(), 8, fmt, 2, 32, mode="spec")
    assert_exact(a.dequantize(), y, "dequantized activations")
