"""GPU tier: tcgen05 GEMM and Hessian accumulation against torch fp32/fp64 references.
Floating-point kernels: tolerance stated per test."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def TC():
    from msq_b200 import _tc
    return _tc


@pytest.mark.timeout(180)
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (384, 768, 1024), (100, 200, 72), (4096, 4096, 512)])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_gemm_nt(TC, M, N, K, dt):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g).to(dt)
    B = torch.randn(N, K, device="cuda", generator=g).to(dt)
    C = TC.gemm_nt(A, B)
    ref = A.double() @ B.double().t()
    # fp16/bf16 products are exact; only the fp32 accumulation order differs: |err| <= K * 2^-24 * sum|a||b|
    tol = 4 * K * 2.0 ** -24 * float((A.double().abs() @ B.double().abs().t()).max())
    assert float((C.double() - ref).abs().max()) <= tol
    C2 = TC.gemm_nt(A, B, out=C.clone(), alpha=0.5, beta=2.0)
    assert float((C2.double() - (0.5 * ref + 2.0 * C.double())).abs().max()) <= 4 * tol


@pytest.mark.timeout(180)
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("T,K", [(256, 768), (2048, 4096), (100, 96)])
def test_hessian_accumulate(TC, dt, T, K):
    """GPTQ.add_batch (llm/gptq.py:53-58): running mean of 2 X^T X over 3 batches"""
    import msq_oracle as O
    g = torch.Generator().manual_seed(2)
    H = torch.zeros(K, K, device="cuda")
    Href, ns, n = torch.zeros(K, K, dtype=torch.float64), 0, 0
    for b in range(3):
        X = torch.randn(1, T, K, generator=g).to(dt)
        X[..., :4] *= 5.0
        n_new = n + 1
        TC.hessian_accumulate(H, X[0].cuda(), alpha=2.0 / n_new, beta=n / n_new)
        n = n_new
        Href = Href * (ns / (ns + 1)) + (2.0 / (ns + 1)) * (X[0].double().t() @ X[0].double())
        ns += 1
    err = float((H.cpu().double() - Href).abs().max()) / float(Href.abs().max())
    # 16-bit inputs: exact products, fp32 tensor-core accumulation over T terms (measured ~1e-5 at T=2048);
    # fp32 inputs: two-term fp16 split (~2^-21 per product) on top of that
    assert err <= (5e-5 if dt != torch.float32 else 1e-4), err
    assert torch.equal(H, H.t())
