"""GPU tier: the packed-weight GEMM that consumes what the quantizers EMIT -- row blocks along the output features
(axes=[0], block 16 / 32: utils/quant.py:147-266 as llm/llama.py:130-137 calls it), inlier-width slots, bitmap and
the segment-compacted hi stream -- plus the rank index that gives it random access into that stream.
Reference = fp64 matmul of the DEQUANTIZED weight (number_system/mx/linear.py:91 multiplies fake-quantized operands);
the dequantized weight itself is pinned to the oracle in test_gpu_pack.py.  Tolerance 4 K 2^-24 (|x| |W|^T)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import msq_b200
    return msq_b200


def weights(rows, cols, seed):
    g = torch.Generator().manual_seed(seed)
    W = torch.randn(rows, cols, generator=g) * 0.02
    idx = torch.randperm(W.numel(), generator=g)[: max(1, W.numel() // 1000)]
    W.view(-1)[idx] *= 20
    return W


@pytest.mark.parametrize("N,K,fi,fo,bs", [(128, 64, "fp4_e2m1", "fp8_e4m3", 16), (256, 192, "fp4_e2m1", "fp8_e4m3", 16),
                                          (1024, 4096, "fp4_e2m1", "fp8_e4m3", 16), (4096, 1024, "int2", "fp4", 16),
                                          (512, 512, "fp4_e2m1", "fp4_e2m1", 32), (384, 640, "fp4_e2m1", "fp8_e4m3", 32),
                                          (2048, 2048, "int2", "fp4_e2m1", 16)])
def test_hi_index_matches_the_bitmap(M, N, K, fi, fo, bs):
    pw = M.pack_mx_outlier(weights(N, K, N + K).cuda(), 8, 8, fi, fo, 2, [0], bs)
    idx = pw.hi_index().cpu().long()
    mask = pw.outlier_mask().cpu()                                   # [N, K]
    cnt = mask.reshape(N // bs, bs, K).sum(1).reshape(-1)            # outliers per block, block-linear order b*K + k
    seg = 1024
    nb = cnt.numel()
    pad = (-nb) % seg
    c = torch.nn.functional.pad(cnt, (0, pad)).reshape(-1, seg)
    ex = (torch.cumsum(c, 1) - c).reshape(-1)[:nb]
    assert torch.equal(idx, ex[::8])
    if pw.hi_w:
        assert torch.equal(pw.seg_cnt.cpu().long(), c.sum(1))


@pytest.mark.parametrize("N,K,fi,fo,bs", [(128, 64, "fp4_e2m1", "fp8_e4m3", 16), (256, 192, "fp4_e2m1", "fp8_e4m3", 16),
                                          (1024, 4096, "fp4_e2m1", "fp8_e4m3", 16), (4096, 1024, "int2", "fp4", 16),
                                          (512, 512, "fp4_e2m1", "fp4_e2m1", 32), (384, 640, "fp4_e2m1", "fp8_e4m3", 32)])
@pytest.mark.parametrize("Mx", [1, 16, 100, 300])
@pytest.mark.parametrize("via_mxfp8", [False, True])
def test_rowblock_gemm_against_dequantized_weight(M, monkeypatch, N, K, fi, fo, bs, Mx, via_mxfp8):
    """via_mxfp8: the artefact re-encoded once, exactly, as K-blocked MX-FP8 and run on the block-scaled kernels with the
    bf16 activations as two block-scaled E4M3 terms (K % 128 == 0); else the row-block kernel (k_gemm_rb)"""
    from msq_b200 import _tc
    monkeypatch.setattr(_tc, "ROWBLOCK_VIA_MXFP8", via_mxfp8)
    pw = M.pack_mx_outlier(weights(N, K, N * 3 + K).cuda(), 8, 8, fi, fo, 2, [0], bs)
    assert _tc.rowblock_ok(pw)
    if via_mxfp8 and K % 128 == 0:
        op = _tc.rowblock_mxfp8_operand(pw)
        assert op is not None                                     # exact by construction (checked inside), and it is:
        wb, wsf = op
        w8 = (wb.cpu().view(torch.float8_e4m3fn).float().view(N, K // 32, 32) * torch.exp2(wsf.cpu().float() - 127).view(N, K // 32, 1)).view(N, K)
        if pw._mx8_dense_resid is not None:                       # elements far below their block's maximum: exact remainders
            rn, rk, rdw = (t.cpu() for t in pw._mx8_dense_resid)
            w8[rn.long(), rk.long()] += rdw
        assert torch.equal(w8, pw.dequantize().cpu())
    g = torch.Generator().manual_seed(Mx)
    x = torch.randn(Mx, K, generator=g).to(torch.bfloat16).cuda()
    y = _tc.packed_linear(x, pw).double().cpu()
    Wd = pw.dequantize().double().cpu()
    ref = x.double().cpu() @ Wd.t()
    bound = 4 * K * 2.0 ** -24 * (x.double().cpu().abs() @ Wd.abs().t()) + 1e-30
    assert bool(((y - ref).abs() <= bound).all()), float(((y - ref).abs() / bound).max())
    assert (getattr(pw, "_mx8_dense", None) is not None) == (via_mxfp8 and K % 128 == 0)


def test_rowblock_weight_without_an_exact_mxfp8_form_keeps_the_rowblock_kernel(M):
    """a 32-block along K whose elements span more binades than E4M3 holds has no exact MX-FP8 form: the conversion
    reports it (once) and packed_linear stays on the row-block kernel"""
    from msq_b200 import _tc
    W = weights(256, 256, 4)
    W[:, 7::32] *= 2.0 ** 30                                         # every block holds a column 30 binades above its neighbours
    pw = M.pack_mx_outlier(W.cuda(), 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16)
    assert _tc.rowblock_mxfp8_operand(pw) is None and pw._mx8_dense_failed
    x = torch.randn(8, 256).to(torch.bfloat16).cuda()
    y = _tc.packed_linear(x, pw).double().cpu()
    Wd = pw.dequantize().double().cpu()
    ref = x.double().cpu() @ Wd.t()
    bound = 4 * 256 * 2.0 ** -24 * (x.double().cpu().abs() @ Wd.abs().t()) + 1e-30
    assert bool(((y - ref).abs() <= bound).all())


def test_rowblock_gemm_other_dtypes_large_m_and_fallback(M):
    from msq_b200 import _tc
    N, K = 512, 256
    pw = M.pack_mx_outlier(weights(N, K, 9).cuda(), 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16)
    Wd = pw.dequantize().double().cpu()
    g = torch.Generator().manual_seed(1)
    for dt, m in ((torch.float32, 40), (torch.float16, 40), (torch.bfloat16, 700)):      # 700 rows: dense bf16 route
        x = torch.randn(m, K, generator=g).to(dt).cuda()
        y = _tc.packed_linear(x, pw, bias=torch.ones(N).cuda()).double().cpu()
        ref = x.double().cpu() @ Wd.t() + 1.0
        bound = 4 * K * 2.0 ** -24 * (x.double().cpu().abs() @ Wd.abs().t()) + 1e-6
        assert bool(((y - ref).abs() <= bound).all())
    # a geometry the kernel does not cover (rows not a multiple of the block, ragged K) goes through unpack + dense GEMM
    pw2 = M.pack_mx_outlier(weights(200, 96, 3).cuda(), 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16)
    assert not _tc.rowblock_ok(pw2)
    x = torch.randn(5, 96, generator=g).to(torch.bfloat16).cuda()
    y = _tc.packed_linear(x, pw2).double().cpu()
    Wd2 = pw2.dequantize().double().cpu()
    assert float((y - x.double().cpu() @ Wd2.t()).abs().max()) <= 1e-4


def test_mxlinear_from_packed_and_checkpoint_round_trip(M, tmp_path):
    """f3: save -> load -> MXLinear.from_packed runs the packed GEMM on the stored planes; load(save(p)).dequantize()
    is the reference's fake-quantized weight bit for bit (pack parity is pinned in test_gpu_pack.py)."""
    from msq_b200 import packed
    from util import assert_exact
    W = weights(256, 128, 21).cuda()
    ref = M.quantize_mx_outlier_v1(W, 8, 8, "fp4_e2m1", "fp8_e4m3", "max", 2, [0], 16)
    pw = M.pack_mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16)
    pw.hi_index()
    path = str(tmp_path / "w.msqpack")
    packed.save_packed(path, {"fc.weight": pw, "fc.bias": torch.arange(256.0)}, {"note": "test"})
    d = packed.load_packed(path, "cuda")
    assert_exact(d["fc.weight"].dequantize(), ref, "load(save(p)).dequantize()")
    assert torch.equal(d["fc.weight"].hi_off8, pw.hi_off8)
    lin = M.MXLinear.from_packed(d["fc.weight"], bias=d["fc.bias"], mx_specs={"a_elem_format": None, "block_size": 32})
    x = torch.randn(7, 128, generator=torch.Generator().manual_seed(0)).cuda()
    y = lin(x)
    yref = torch.nn.functional.linear(x.double(), ref.double(), torch.arange(256.0, dtype=torch.float64).cuda())
    assert float((y.double() - yref).abs().max()) <= 1e-4 * float(yref.abs().max())
    # quantize_model swaps nn.Linear for MXLinear built on the packed planes
    from msq_b200.utils.quant_model import quantize_model
    net = torch.nn.Sequential(torch.nn.Linear(128, 256)).cuda()
    qnet = quantize_model(net, {"a_elem_format": None, "block_size": 32}, packed={"0.weight": d["fc.weight"], "0.bias": d["fc.bias"]})
    assert type(qnet[0]) is M.MXLinear
    assert torch.equal(qnet(x), y)
