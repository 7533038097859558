"""conversion-instruction pack core (core_outlier_cvt, k_pack_tma<16, true>) against the generic arithmetic
(MSQ_PACK_CVT=0, same kernel) and against k_strided<.., PACK> (MSQ_QUANT_TMA=0) on adversarial inputs: exact rounding
ties of both formats, the neighbours of every tie, the one magnitude per format where fl(|y| + 0.5) rounds up
(half the smallest subnormal minus an ulp), negative values that round to zero, saturation, zero blocks, blocks
without outliers / without inliers, shared exponents at the edge of the conversion path's range, NaN / Inf."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import msq_b200

dev = torch.device("cuda:0")
ok = True


def pack(W, mode):
    os.environ["MSQ_QUANT_TMA"] = "0" if mode == "strided" else "1"
    os.environ["MSQ_PACK_CVT"] = "0" if mode == "generic" else "1"
    p = msq_b200.pack_mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", 2, [0], 16, check=False)
    os.environ.pop("MSQ_PACK_CVT", None)
    os.environ["MSQ_QUANT_TMA"] = "1"
    return p


def same(a, b, tag):
    global ok
    good = True
    for name in ("codes", "bitmap", "e_in", "e_out", "seg_cnt"):
        x, y = getattr(a, name), getattr(b, name)
        if not torch.equal(x, y):
            good = False
            print("MISMATCH", tag, name, int((x != y).sum()))
    if int(a.status.item()) != int(b.status.item()):
        good = False
        print("MISMATCH", tag, "status", int(a.status.item()), int(b.status.item()))
    st = a.hi_seg_stride
    for sgi in range(a.nseg):
        nb = (int(a.seg_cnt[sgi]) * a.hi_w + 7) // 8
        if not torch.equal(a.hi[sgi * st: sgi * st + nb], b.hi[sgi * st: sgi * st + nb]):
            good = False
            print("MISMATCH", tag, "hi segment", sgi)
            break
    ok = ok and good
    return good


def ulp_shift(x, k):
    return (x.view(torch.int32) + k).view(torch.float32)


g = torch.Generator(device=dev).manual_seed(1234)
cases = {}
L, C = 256, 1024
# 1. dyadic grid values (few mantissa bits) and their fp32 neighbours: ties of both formats everywhere
m = torch.randint(-64, 65, (L, C), device=dev, generator=g).float() / 16.0
j = torch.randint(-12, 6, (L, C), device=dev, generator=g).float()
colscale = torch.randint(-30, 31, (1, C), device=dev, generator=g).float()
W = m * torch.exp2(j + colscale)
d = torch.randint(-1, 2, (L, C), device=dev, generator=g, dtype=torch.int32)
W = torch.where(W != 0, (W.view(torch.int32) + d).view(torch.float32), W)
cases["dyadic +-1ulp"] = W.contiguous()
# 2. gaussian weights with planted half-subnormal-minus-an-ulp magnitudes (found from a first pass's exponents)
W = torch.randn(L, C, device=dev, generator=g) * 0.02
p0 = pack(W, "strided")
e_in = p0.e_in.view(L // 16, C).float().repeat_interleave(16, 0)
e_out = p0.e_out.view(L // 16, C).float().repeat_interleave(16, 0)
sel = torch.rand(L, C, device=dev, generator=g)
sgn = torch.where(torch.rand(L, C, device=dev, generator=g) < 0.5, -1.0, 1.0)
sp_in = ulp_shift(torch.exp2(e_in - 2), -1)
W = torch.where(sel < 0.04, sp_in * sgn, W)                       # pred(2^(e_in - 2)): inlier special (if it stays an inlier)
W = torch.where((sel >= 0.04) & (sel < 0.06), torch.exp2(e_in - 2) * sgn, W)   # the exact tie next to it
sp_out = ulp_shift(torch.exp2(e_out - e_in - 10), -1)
W = torch.where((sel >= 0.06) & (sel < 0.08) & (e_out > -100), -sp_out, W)     # negative => below the lower bound => outlier special
cases["planted specials"] = W.contiguous()
# 3. tiny negatives (round to -0), zeros, zero blocks, constant blocks (no outliers), saturating values
W = torch.randn(L, C, device=dev, generator=g)
W[:, :64] = 0
W[:16, 64:128] = 0
W[16:32, 128:192] = 1.0
W[32:48, 128:192] = -3.0
W[:, 192:256] = torch.where(torch.rand(L, 64, device=dev, generator=g) < 0.5, -1e-9, 1.0)
W[:, 256:320] *= torch.exp2(torch.randint(-8, 9, (L, 64), device=dev, generator=g).float())
W[:, 320:384] = -torch.rand(L, 64, device=dev, generator=g).abs() - 0.5       # all negative: every element below the lower bound
cases["zeros / constants / negatives"] = W.contiguous()
# 4. shared exponents around the limits of the conversion path (|e_in| <= 40, |e_out| <= 80) and far beyond
W = torch.randn(L, C, device=dev, generator=g)
W *= torch.exp2(torch.randint(-125, 120, (1, C), device=dev, generator=g).float())
W[:, ::7] *= torch.exp2(torch.tensor(40.0, device=dev))
cases["extreme scales"] = torch.nan_to_num(W, posinf=3e38, neginf=-3e38).contiguous()
W2 = torch.randn(L, C, device=dev, generator=g) * torch.exp2(torch.randint(36, 46, (1, C), device=dev, generator=g).float())
cases["scales near +40"] = W2.contiguous()
W2 = torch.randn(L, C, device=dev, generator=g) * torch.exp2(torch.randint(-44, -34, (1, C), device=dev, generator=g).float())
cases["scales near -40"] = W2.contiguous()
# 5. non-finite input
W = torch.randn(L, C, device=dev, generator=g)
W[3, 5] = float("nan"); W[40, 300] = float("inf"); W[100, 700] = -float("inf")
cases["nan / inf"] = W.contiguous()
# 6. heavy-tailed
W = torch.randn(L, C, device=dev, generator=g) ** 3
cases["cubed gaussian"] = W.contiguous()

for tag, W in cases.items():
    a, b, c = pack(W, "cvt"), pack(W, "generic"), pack(W, "strided")
    assert a.path == 5 and b.path == 5 and c.path == 1, (a.path, b.path, c.path)
    r1 = same(a, b, tag + " [cvt vs generic]")
    r2 = same(a, c, tag + " [cvt vs strided]")
    extra = ""
    if tag == "planted specials":
        ein = a.e_in.view(L // 16, C).float().repeat_interleave(16, 0)
        eo = a.e_out.view(L // 16, C).float().repeat_interleave(16, 0)
        bm = a.dequantize() * 0  # placeholder to force a decode
        hit_in = int((W.abs() == ulp_shift(torch.exp2(ein - 2), -1)).sum())
        hit_out = int(((W.abs() == ulp_shift(torch.exp2(eo - ein - 10), -1)) & (eo > -100)).sum())
        extra = "  (magnitudes equal to the inlier special: %d, outlier special: %d)" % (hit_in, hit_out)
    print("%-32s %s%s" % (tag, "identical" if (r1 and r2) else "DIFFERENT", extra), flush=True)
print("all identical:", ok)
sys.exit(0 if ok else 1)
