"""GPU tier: GPTQ (Hessian + fasterquant) against the reference fixtures and the oracle."""
import pytest
import torch

import msq_oracle as O
from util import assert_exact, load_json, load_npz, mismatches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import msq_b200
    msq_b200.GPTQ.verbose = False
    return msq_b200


def make_gptq(M, W, fi, fo, bs):
    layer = torch.nn.Linear(W.shape[1], W.shape[0], bias=False).cuda()
    layer.weight.data = W.clone().cuda()
    gq = M.GPTQ(layer)
    gq.quantizer = M.MXQuantizer()
    gq.quantizer.configure(8, 8, fi, fo, axes=[0], block_size=bs)
    return layer, gq


@pytest.mark.timeout(300)
def test_hessian_matches_reference_fixture(M):
    z = load_npz("gptq.npz")
    for n, c in enumerate(load_json("gptq_cases.json")):
        layer, gq = make_gptq(M, z["W_%d" % n], c["in_fmt"], c["out_fmt"], c["bs"])
        X = z["X_%d" % n]
        for b in range(X.shape[0]):
            gq.add_batch(X[b:b + 1].cuda(), None)
        assert gq.nsamples == X.shape[0]
        ref = z["H_%d" % n]
        assert float((gq.H.cpu() - ref).abs().max()) <= 5e-5 * float(ref.abs().max())
        # batched accumulation gives the same matrix
        layer2, g2 = make_gptq(M, z["W_%d" % n], c["in_fmt"], c["out_fmt"], c["bs"])
        g2.hessian_batch = 4
        for b in range(X.shape[0]):
            g2.add_batch(X[b:b + 1].cuda(), None)
        assert float((g2.H.cpu() - ref).abs().max()) <= 5e-5 * float(ref.abs().max())


@pytest.mark.timeout(300)
def test_fused_solver_on_reference_fixtures(M):
    """same Hinv in.  The fused kernel must (a) refuse exactly when some column has fewer zeros than
    the reference's k (then the prune changes values), (b) otherwise reproduce the reference;
    bit for bit when there is a single GPTQ block (no trailing GEMM)."""
    z = load_npz("gptq.npz")
    for n, c in enumerate(load_json("gptq_cases.json")):
        if c["blocksize"] != 128:
            continue
        cfg = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format=c["in_fmt"],
                   outlier_elem_format=c["out_fmt"], block_size=c["bs"])
        _, _, info = O.fasterquant(z["W_%d" % n], z["Hinv_%d" % n], cfg, 128)
        layer, gq = make_gptq(M, z["W_%d" % n], c["in_fmt"], c["out_fmt"], c["bs"])
        Q, loss = gq._solve_fused(z["W_%d" % n].cuda().clone(), z["Hinv_%d" % n].cuda().contiguous())
        if info["columns_with_zeros_lt_k"] > 0:
            assert Q is None, n
        else:
            assert_exact(Q, z["Q_%d" % n], "case %d" % n)
            assert abs(loss - float(z["err_%d" % n])) <= 1e-5 * float(z["err_%d" % n])


@pytest.mark.timeout(300)
def test_generic_solver_matches_reference(M):
    z = load_npz("gptq.npz")
    for n, c in enumerate(load_json("gptq_cases.json")):
        layer, gq = make_gptq(M, z["W_%d" % n], c["in_fmt"], c["out_fmt"], c["bs"])
        Q, loss = gq._solve_generic(z["W_%d" % n].cuda().clone(), z["Hinv_%d" % n].cuda().contiguous(), c["blocksize"],
                                    -1, False, False, None, None)
        frac = mismatches(Q, z["Q_%d" % n]) / Q.numel()
        cfg = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format=c["in_fmt"],
                   outlier_elem_format=c["out_fmt"], block_size=c["bs"])
        _, _, info = O.fasterquant(z["W_%d" % n], z["Hinv_%d" % n], cfg, c["blocksize"])
        if c["cols"] <= c["blocksize"] and info["columns_with_zeros_lt_k"] == 0:
            assert frac == 0, (n, frac)
        else:
            # the prune zeroes non-zero values and torch.topk's order among tied importances is
            # unspecified (CPU and CUDA differ -- SURVEY 7 hard parts), plus fp32 summation order of
            # the trailing GEMM: positions may differ, the objective may not
            assert frac < 0.10, (n, frac)
        assert abs(loss - float(z["err_%d" % n])) <= 5e-2 * float(z["err_%d" % n])


@pytest.mark.timeout(300)
def test_fp6_prune_falls_back_to_generic(M):
    z = load_npz("gptq.npz")
    cs = load_json("gptq_cases.json")
    n = [i for i, c in enumerate(cs) if c["in_fmt"].startswith("fp6")][0]
    c = cs[n]
    layer, gq = make_gptq(M, z["W_%d" % n], c["in_fmt"], c["out_fmt"], c["bs"])
    gq.H = z["H_%d" % n].cuda().clone()
    gq.nsamples = 4
    gq.fasterquant(blocksize=128)
    assert gq.last_path == "generic"          # zeros < k: the prune changes values, fused path must refuse
    assert torch.isfinite(layer.weight).all()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("rows,cols,fi,fo,bs", [(1024, 384, "int2", "fp4", 16), (2048, 256, "fp4_e2m1", "fp8_e4m3", 16),
                                                (1000, 300, "int4", "fp8_e4m3", 32), (512, 256, "int2", "fp4_e2m1", 8),
                                                (512, 96, "int2", "fp4", 16)])
def test_fasterquant_vs_oracle_same_hinv(M, rows, cols, fi, fo, bs):
    """fused column loop on the GPU vs the oracle column loop on the CPU, fed the SAME Hinv"""
    g = torch.Generator().manual_seed(rows + cols)
    W = torch.randn(rows, cols, generator=g) * 0.02
    X = torch.randn(1024, cols, generator=g).half()
    X[:, :8] *= 4
    layer, gq = make_gptq(M, W, fi, fo, bs)
    gq.add_batch(X.cuda().unsqueeze(0), None)
    H = gq.H.clone()
    Hinv_gpu = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(
        H + 0.01 * torch.mean(torch.diag(H)) * torch.eye(cols, device="cuda"))), upper=True).contiguous()
    Q, loss = gq._solve_fused(W.cuda().clone(), Hinv_gpu)
    cfg = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format=fi, outlier_elem_format=fo, block_size=bs)
    Qo, loss_o, info = O.fasterquant(W, Hinv_gpu.cpu(), cfg, 128)
    if info["columns_with_zeros_lt_k"] > 0:
        assert Q is None
        return
    assert Q is not None
    # first GPTQ block never saw a trailing update: bit exact
    assert_exact(Q[:, :128], Qo[:, :128])
    frac = mismatches(Q, Qo) / Q.numel()
    assert frac < 0.01, frac                   # only the fp32 summation order of the trailing update differs
    assert abs(loss - loss_o) <= 1e-2 * loss_o


@pytest.mark.timeout(600)
def test_fasterquant_llama_shape_improves_on_rtn(M):
    """LLaMA-2-7B attention-sized linear: GPTQ output error must beat round-to-nearest"""
    g = torch.Generator().manual_seed(5)
    W = torch.randn(4096, 4096, generator=g) * 0.02
    X = (torch.randn(2048, 4096, generator=g) * (1 + 3 * torch.rand(4096, generator=g))).half()
    layer, gq = make_gptq(M, W, "int2", "fp4", 16)
    gq.add_batch(X.cuda().unsqueeze(0), None)
    gq.fasterquant()
    assert gq.last_path == "fused"
    Wq = layer.weight.data.float()
    rtn = M.quantize_mx_outlier_v1(W.cuda(), 8, 8, "int2", "fp4", "max", 2, [0], 16)
    Xf = X.cuda().float()
    e_gptq = float(((W.cuda() - Wq) @ Xf.t()).pow(2).sum())
    e_rtn = float(((W.cuda() - rtn) @ Xf.t()).pow(2).sum())
    assert e_gptq < e_rtn


@pytest.mark.timeout(300)
def test_no_prune_variant_vs_oracle(M):
    """GPTQ(layer, prune=False) = the reference's vlm/gptq.py:130-144 column loop (quantize_mx_outlier_v1 per column, no
    inlier zeroed afterwards): the fused solver never refuses -- also on the fixture where prune=True must fall back --
    and its first column block equals the oracle's no-prune loop bit for bit"""
    z = load_npz("gptq.npz")
    for n, c in enumerate(load_json("gptq_cases.json")):
        W, Hinv = z["W_%d" % n], z["Hinv_%d" % n]
        layer = torch.nn.Linear(W.shape[1], W.shape[0], bias=False).cuda()
        layer.weight.data = W.clone().cuda()
        gq = M.GPTQ(layer, prune=False)
        gq.quantizer = M.MXQuantizer()
        gq.quantizer.configure(8, 8, c["in_fmt"], c["out_fmt"], axes=[0], block_size=c["bs"])
        Q, loss = gq._solve_fused(W.cuda().clone(), Hinv.cuda().contiguous())
        assert Q is not None, n
        cfg = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format=c["in_fmt"],
                   outlier_elem_format=c["out_fmt"], block_size=c["bs"])
        Qo, loss_o, _ = O.fasterquant(W, Hinv, cfg, c["blocksize"], prune=False)
        b0 = min(c["blocksize"], W.shape[1])
        assert mismatches(Q[:, :b0], Qo[:, :b0]) == 0, n
        assert mismatches(Q, Qo) / Q.numel() < 0.01, n          # beyond block 0: fp32 order of the trailing update
        assert abs(loss - float(loss_o)) <= 1e-2 * float(loss_o), n
