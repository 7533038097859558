"""GPU tier: quantize+pack and unpack.  Packed planes decode (in pure torch, following the layout
written in DESIGN.md) to exactly the oracle's mask / exponents / element codes, and the CUDA unpack
kernel reproduces the reference's dequantized tensor bit for bit."""
import pytest
import torch

import msq_oracle as O
from util import assert_exact

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


def decode_planes(p, M):
    """reference decoder of the packed layout, written from the spec (slow, torch on CPU)"""
    n = 1
    for s in p.shape:
        n *= s
    wi = p.code_width
    raw = p.codes.cpu().to(torch.int64)
    bits = ((raw.unsqueeze(-1) >> torch.arange(8)) & 1).reshape(-1)[: n * wi].reshape(n, wi)
    slot = (bits << torch.arange(wi)).sum(-1)
    mask = p.outlier_mask().cpu().reshape(-1)
    code = slot.clone()
    if p.hi_w:
        pre, L, post = M._ops.geometry(p.shape, p.axis)
        bs = p.block_size
        nblk = (L + bs - 1) // bs
        lin = torch.arange(n).reshape(pre, L, post)
        hi_raw = p.hi.cpu().to(torch.int64)
        segb = M._native.lib().msq_seg_blocks()
        seg_cnt = p.seg_cnt.cpu()
        # block-major order inside every segment
        order = []
        for pb in range(pre * nblk):
            pp, b = divmod(pb, nblk)
            rows = lin[pp, b * bs:(b + 1) * bs, :]          # [<=bs, post]
            order.append(rows.t())                           # [post, <=bs]: block j, then position i
        blocks = [blk for o in order for blk in o]           # block-linear order
        for s in range(p.nseg):
            k = 0
            base_bit = s * p.hi_seg_stride * 8
            for blk in blocks[s * segb:(s + 1) * segb]:
                for e in blk.tolist():
                    if mask[e]:
                        bit = base_bit + k * p.hi_w
                        v = 0
                        for t in range(p.hi_w):
                            v |= ((int(hi_raw[(bit + t) >> 3]) >> ((bit + t) & 7)) & 1) << t
                        code[e] |= v << wi
                        k += 1
            assert k == int(seg_cnt[s])
    return mask.reshape(p.shape), code.reshape(p.shape)


CASES = [
    # in_fmt, out_fmt, bs, axis, shape, hi_w
    ("int2", "fp4", 16, 0, (96, 64), 2),
    ("fp4_e2m1", "fp8_e4m3", 16, 0, (64, 96), 4),
    ("fp4_e2m1", "fp4_e2m1", 16, 0, (48, 32), 0),
    ("int2", "fp8_e4m3", 8, 0, (40, 64), 8),
    ("fp4_e2m1", "fp8_e4m3", 32, 1, (24, 128), 4),
    ("int2", "fp4", 16, 1, (16, 256), 2),
    ("fp6_e3m2", "fp6_e3m2", 32, 1, (16, 128), 0),
    ("int4", "int8", 64, 1, (8, 256), 4),
    ("fp4_e2m1", "fp8_e4m3", 32, 0, (80, 32), 4),      # ragged rows (80 = 2.5 blocks of 32)
]


@pytest.mark.parametrize("fi,fo,bs,axis,shape,hi_w", CASES)
def test_pack_layout_and_roundtrip(M, fi, fo, bs, axis, shape, hi_w):
    W = weights(shape[0], shape[1], 21)
    p, deq = M.pack_mx_outlier(W.cuda(), 8, 8, fi, fo, 2, [axis], bs, want_deq=True)
    assert p.hi_w == hi_w
    y, art = O.quantize_mx_outlier(W, 8, 8, fi, fo, 2, axis, bs, artefacts=True)
    assert_exact(deq, y, "fused dequant output")
    assert_exact(p.dequantize(), y, "unpack(pack(W))")
    assert_exact(p.dequantize(torch.bfloat16).float(), y.to(torch.bfloat16).float(), "bf16 unpack")
    mask, code = decode_planes(p, M)
    assert_exact(mask.float(), art["mask"], "outlier bitmap")
    pre, L, post = M._ops.geometry(shape, axis)
    nblk = (L + bs - 1) // bs
    assert_exact(p.e_in.view(pre, nblk, post).float(), art["e_in"], "e_in")
    assert_exact(p.e_out.view(pre, nblk, post).float(), art["e_out"], "e_out")
    q = torch.where(art["mask"].bool(), art["qout"], art["qin"])
    dec = torch.where(art["mask"].bool(), O.decode_elem(code, fo), O.decode_elem(code, fi))
    assert_exact(dec, q, "element codes")


def test_pack_plain_roundtrip(M):
    W = weights(64, 256, 3)
    for fmt, bs in (("fp4_e2m1", 32), ("fp6_e2m3", 32), ("int8", 16), ("fp8_e4m3", 64)):
        p = M.pack_mx(W.cuda(), 8, fmt, axes=[1], block_size=bs)
        y, art = O.quantize_mx(W, 8, fmt, 1, bs, mode="spec", artefacts=True)
        assert_exact(p.dequantize(), y, fmt)
        assert_exact(p.e_in.view(64, -1).float(), art["e"].squeeze(-1), fmt)
    p = M.pack_mx(W.cuda(), 8, "fp4_e2m1", axes=[0], block_size=16)
    assert_exact(p.dequantize(), O.quantize_mx(W, 8, "fp4_e2m1", 0, 16, mode="spec"))


def test_pack_llama_shape_bytes_and_roundtrip(M):
    """full-size property: unpack(pack(W)) == fake-quant(W) and the byte count matches DESIGN.md"""
    W = weights(4096, 4096, 0).cuda()
    p = M.pack_mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16)
    y = M.quantize_mx_outlier_v1(W, 8, 8, "fp4_e2m1", "fp8_e4m3", "max", 2, [0], 16)
    assert_exact(p.dequantize(), y)
    rho = float(p.outlier_mask().float().mean())
    per_elem = p.nbytes / W.numel()
    expect = 0.5 + 2 / 16 + 1 / 8 + 0.5 * rho           # codes + 2 exponents / 16 + bitmap + 4-bit hi * rho
    assert abs(per_elem - expect) < 0.01, (per_elem, expect, rho)
    with pytest.raises(RuntimeError):                   # packed fast path needs 32-column words
        M.pack_mx_outlier(W[:, :100].contiguous(), 8, 8, "int2", "fp4", 2, [0], 16)


@pytest.mark.parametrize("fi,fo,bs,axis,shape", [
    ("fp4_e2m1", "fp8_e4m3", 16, 0, (1000, 1056)),     # many segments, ragged last block row, strided layout
    ("fp4_e2m1", "fp8_e4m3", 32, 1, (520, 2048)),      # many segments, innermost layout with a hi stream
    ("fp4_e2m1", "fp4_e2m1", 32, 1, (768, 4096)),      # GEMM weight operand layout (no hi stream)
    ("int2", "fp8_e4m3", 8, 0, (264, 2080)),           # 8-bit... 6-bit hi field (run-time hi width), bs 8
    ("int2", "fp4", 16, 1, (300, 1024)),               # 2-bit slots, 2-bit hi field, innermost
    ("fp6_e3m2", "fp6_e3m2", 32, 0, (640, 1024)),      # 8-bit containers
    ("int4", "int8", 64, 1, (96, 4096)),               # 64-element blocks, innermost
])
def test_multi_segment_roundtrip(M, fi, fo, bs, axis, shape):
    """unpack(pack(W)) == fake-quant(W) bit for bit across segment boundaries, both unpack layouts, fp32 and bf16"""
    W = weights(shape[0], shape[1], 33).cuda()
    p = M.pack_mx_outlier(W, 8, 8, fi, fo, 2, [axis], bs)
    y = M.quantize_mx_outlier_v1(W, 8, 8, fi, fo, "max", 2, [axis], bs)
    assert_exact(p.dequantize(), y, "fp32")
    assert_exact(p.dequantize(torch.bfloat16).float(), y.to(torch.bfloat16).float(), "bf16")
    # the oracle on a slice that crosses the first segment boundary
    ref = O.quantize_mx_outlier(W.cpu()[:, :64] if axis == 0 else W.cpu()[:8], 8, 8, fi, fo, 2, axis, bs)
    assert_exact(y[:, :64] if axis == 0 else y[:8], ref, "oracle slice")


def _ulp(x, k):
    return (x.view(torch.int32) + k).view(torch.float32)


def _adversarial(kind, L=64, C=256):
    """inputs aimed at the conversion-instruction core (core_outlier_cvt): exact rounding ties of both formats and their
    fp32 neighbours, the magnitude below half the smallest subnormal where the reference's fl(|y| + 0.5) rounds up,
    negative values that round to zero, saturation, zero / constant / all-negative blocks, shared exponents at and
    beyond the edge of the core's range, non-finite values"""
    g = torch.Generator().manual_seed(77)
    if kind == "ties":
        m = torch.randint(-64, 65, (L, C), generator=g).float() / 16.0
        j = torch.randint(-12, 6, (L, C), generator=g).float()
        W = m * torch.exp2(j + torch.randint(-30, 31, (1, C), generator=g).float())
        d = torch.randint(-1, 2, (L, C), generator=g, dtype=torch.int32)
        return torch.where(W != 0, (W.view(torch.int32) + d).view(torch.float32), W).contiguous()
    if kind == "specials":
        W = torch.randn(L, C, generator=g) * 0.02
        _, art = O.quantize_mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 0, 16, artefacts=True)
        e_in = art["e_in"].reshape(L // 16, C).repeat_interleave(16, 0)
        e_out = art["e_out"].reshape(L // 16, C).repeat_interleave(16, 0)
        sel = torch.rand(L, C, generator=g)
        sgn = torch.where(torch.rand(L, C, generator=g) < 0.5, -1.0, 1.0)
        W = torch.where(sel < 0.05, _ulp(torch.exp2(e_in - 2), -1) * sgn, W)
        W = torch.where((sel >= 0.05) & (sel < 0.08), torch.exp2(e_in - 2) * sgn, W)
        W = torch.where((sel >= 0.08) & (sel < 0.11) & (e_out > -100), -_ulp(torch.exp2(e_out - e_in - 10), -1), W)
        return W.contiguous()
    if kind == "degenerate":
        W = torch.randn(L, C, generator=g)
        W[:, :32] = 0
        W[:16, 32:64] = 0
        W[16:32, 64:96] = 1.0
        W[32:48, 64:96] = -3.0
        W[:, 96:128] = torch.where(torch.rand(L, 32, generator=g) < 0.5, -1e-9, 1.0)
        W[:, 128:160] *= torch.exp2(torch.randint(-8, 9, (L, 32), generator=g).float())
        W[:, 160:192] = -torch.rand(L, 32, generator=g) - 0.5
        return W.contiguous()
    if kind == "scales":
        W = torch.randn(L, C, generator=g)
        e = torch.cat([torch.randint(-44, -34, (1, C // 4), generator=g), torch.randint(36, 46, (1, C // 4), generator=g),
                       torch.randint(-120, 110, (1, C // 2), generator=g)], 1).float()
        W = W * torch.exp2(e)
        W[:, ::7] *= 2.0 ** 40
        return torch.nan_to_num(W, posinf=3e38, neginf=-3e38).contiguous()
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["ties", "specials", "degenerate", "scales"])
def test_pack_conversion_core_vs_oracle(M, kind, monkeypatch):
    """k_pack_tma<16, CVT> (fp4_e2m1 / fp8_e4m3, nearest: rounding and codes by cvt.rn.satfinite.e2m1x2 / e4m3x2) gives
    the oracle's mask, exponents and element codes, and the planes of the generic arithmetic, on adversarial inputs"""
    W = _adversarial(kind)
    monkeypatch.setenv("MSQ_QUANT_TMA", "1")
    monkeypatch.delenv("MSQ_PACK_CVT", raising=False)
    p = M.pack_mx_outlier(W.cuda(), 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16, check=False)
    assert p.path == 5
    monkeypatch.setenv("MSQ_PACK_CVT", "0")
    r = M.pack_mx_outlier(W.cuda(), 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16, check=False)
    assert r.path == 5
    for name in ("codes", "bitmap", "e_in", "e_out", "seg_cnt"):
        assert torch.equal(getattr(p, name), getattr(r, name)), name
    for s in range(p.nseg):
        nb = (int(p.seg_cnt[s]) * p.hi_w + 7) // 8
        assert torch.equal(p.hi[s * p.hi_seg_stride: s * p.hi_seg_stride + nb], r.hi[s * p.hi_seg_stride: s * p.hi_seg_stride + nb])
    assert int(p.status.item()) == int(r.status.item())
    if kind == "scales":
        # sums that overflow fp32: the reference asserts (and so does the oracle); the planes are those of the direct-load
        # kernel (k_strided<.., PACK>, compared with the oracle by the tests above)
        monkeypatch.setenv("MSQ_QUANT_TMA", "0")
        d = M.pack_mx_outlier(W.cuda(), 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16, check=False)
        assert d.path == 1
        for name in ("codes", "bitmap", "e_in", "e_out", "seg_cnt"):
            assert torch.equal(getattr(p, name), getattr(d, name)), name
        assert int(p.status.item()) == int(d.status.item())
        return
    y, art = O.quantize_mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 0, 16, artefacts=True)
    mask, code = decode_planes(p, M)
    L, C = W.shape
    ok = torch.isfinite(y).reshape(L // 16, 16, C).all(1)            # blocks the reference would assert on are NaN in the oracle
    assert_exact(mask.float()[ok.repeat_interleave(16, 0)], art["mask"][ok.repeat_interleave(16, 0)], "outlier bitmap")
    assert_exact(p.e_in.view(L // 16, C).float()[ok], art["e_in"].reshape(L // 16, C)[ok], "e_in")
    q = torch.where(art["mask"].bool(), art["qout"], art["qin"])
    dec = torch.where(art["mask"].bool(), O.decode_elem(code, "fp8_e4m3"), O.decode_elem(code, "fp4_e2m1"))
    okE = ok.repeat_interleave(16, 0)
    assert_exact(dec[okE], q[okE], "element codes")
    assert_exact(p.dequantize().cpu()[okE], y[okE], "unpack(pack(W))")
    if kind == "specials":
        e_in = art["e_in"].reshape(L // 16, C).repeat_interleave(16, 0)
        hit = (W.abs() == _ulp(torch.exp2(e_in - 2), -1)) & ~art["mask"].bool()
        assert int(hit.sum()) > 50                    # the patched magnitude really occurs ...
        assert bool((dec[hit].abs() == 0.5).all())    # ... and rounds UP to the smallest subnormal, as the reference does


@pytest.mark.parametrize("sb", [4, 5, 6])
@pytest.mark.parametrize("kind", ["ties", "degenerate"])
def test_pack_conversion_core_narrow_scales(M, kind, sb, monkeypatch):
    """scale widths whose clamp binds inside the conversion core's exponent range: those blocks must take the generic
    arithmetic (clamped exponents, overflow status) -- planes and status equal to the generic kernel's"""
    W = _adversarial(kind).cuda()
    monkeypatch.setenv("MSQ_QUANT_TMA", "1")
    monkeypatch.delenv("MSQ_PACK_CVT", raising=False)
    p = M.pack_mx_outlier(W, sb, sb, "fp4_e2m1", "fp8_e4m3", 2, [0], 16, check=False)
    monkeypatch.setenv("MSQ_PACK_CVT", "0")
    r = M.pack_mx_outlier(W, sb, sb, "fp4_e2m1", "fp8_e4m3", 2, [0], 16, check=False)
    assert p.path == 5 and r.path == 5
    for name in ("codes", "bitmap", "e_in", "e_out", "seg_cnt"):
        assert torch.equal(getattr(p, name), getattr(r, name)), name
    for s in range(p.nseg):
        nb = (int(p.seg_cnt[s]) * p.hi_w + 7) // 8
        assert torch.equal(p.hi[s * p.hi_seg_stride: s * p.hi_seg_stride + nb], r.hi[s * p.hi_seg_stride: s * p.hi_seg_stride + nb])
    assert int(p.status.item()) == int(r.status.item())


@pytest.mark.parametrize("shape", [(16, 256), (48, 512), (64, 1024), (528, 768)])
def test_unpack_packed_words_small_and_partial_segments(M, shape):
    """k_unpack_strided_cvt on tensors smaller than one segment / with a partial last segment (threads past the last
    block read the first row block and store nothing): unpack(pack(W)) == the oracle, fp32 and bf16"""
    W = weights(shape[0], shape[1], 5)
    p = M.pack_mx_outlier(W.cuda(), 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16)
    y = O.quantize_mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 0, 16)
    assert_exact(p.dequantize(), y, "fp32")
    assert_exact(p.dequantize(torch.bfloat16).float(), y.to(torch.bfloat16).float(), "bf16")
