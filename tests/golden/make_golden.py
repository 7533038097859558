"""Generate tests/golden/*.npz by running the REFERENCE itself (imported in place from
/root/reference, CPU) on seeded inputs.  Run in the authoring container only:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, these fixtures do.  Nothing here is imported
by the product.  mx_core_vectors.json (the reference's own known-answer tables) is written
by hand-transcription in make_core_vectors() with the source file:line of each table.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
REF = "/root/reference"
sys.path[:0] = [REF + "/number_system", REF, REF + "/llm"]
torch.cuda.synchronize = lambda *a, **k: None          # llm/gptq.py:173 calls it unconditionally
HERE = os.path.dirname(os.path.abspath(__file__))

from utils.quant import (quantize_mx_outlier_v1, quantize_mx_outlier_hessian, Quantizer, quantize,  # noqa: E402
                         MXQuantizer)
from mx.mx_ops import _quantize_mx, _quantize_mx_outlier_v1  # noqa: E402
from mx.elemwise_ops import _quantize_bfloat, _quantize_elemwise_core  # noqa: E402
from mx.formats import _get_format_params  # noqa: E402
import gptq as ref_gptq  # noqa: E402


def weights(rows, cols, seed, outliers=True):
    g = torch.Generator().manual_seed(seed)
    W = torch.randn(rows, cols, generator=g) * 0.02
    if outliers:
        idx = torch.randperm(W.numel(), generator=g)[: max(1, W.numel() // 1000)]
        W.view(-1)[idx] *= 20
    return W


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        out[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


OUTLIER_CASES = [
    # name, in_fmt, out_fmt, bs, axis, in_sb, out_sb, std_dev, round, shape
    ("ref_default", "int2", "fp4", 16, 0, 8, 8, 2, "nearest", (256, 96)),
    ("fp4_fp8", "fp4_e2m1", "fp8_e4m3", 16, 0, 8, 8, 2, "nearest", (128, 64)),
    ("fp4_fp8_b32_inner", "fp4_e2m1", "fp8_e4m3", 32, 1, 8, 8, 2, "nearest", (48, 256)),
    ("fp4_fp8_b32_strided", "fp4_e2m1", "fp8_e4m3", 32, 0, 8, 8, 2, "nearest", (128, 64)),
    ("fp6_fp6_sb4", "fp6_e3m2", "fp6_e3m2", 32, 1, 4, 4, 2, "nearest", (32, 128)),
    ("int4_int8_even", "int4", "int8", 16, 1, 8, 8, 3, "even", (40, 64)),
    ("int4_fp8e5_floor_b128", "int4", "fp8_e5m2", 128, 1, 8, 8, 2, "floor", (16, 512)),
    ("ragged_generic", "int8", "int8", 10, 1, 8, 8, 2, "nearest", (33, 57)),
    ("ragged_strided", "fp4_e2m1", "fp6_e2m3", 16, 0, 8, 8, 1.5, "nearest", (50, 36)),
    ("threed_mid_axis", "int2", "fp4", 16, 1, 8, 8, 2, "nearest", (3, 64, 20)),
]


def make_outlier():
    d = {}
    meta = []
    for n, (name, fi, fo, bs, ax, sbi, sbo, sd, rnd, shape) in enumerate(OUTLIER_CASES):
        W = weights(int(np.prod(shape[:-1])), shape[-1], 100 + n).reshape(shape)
        y = quantize_mx_outlier_v1(W, sbi, sbo, fi, fo, "max", sd, [ax], bs, rnd, False, False)
        d["x_" + name] = W
        d["y_" + name] = y
        meta.append(dict(name=name, in_fmt=fi, out_fmt=fo, bs=bs, axis=ax, in_sb=sbi, out_sb=sbo, std_dev=sd,
                         round=rnd))
    # hessian variant: column vectors, as GPTQ.fasterquant calls it (llm/gptq.py:130-143)
    for n, rows in enumerate((4096, 768, 1000)):
        w = weights(rows, 1, 200 + n)
        y, cnt = quantize_mx_outlier_hessian(w, 8, 8, "int2", "fp4", "max", 2, [0], 16, "nearest", False, False)
        d["hx_%d" % rows] = w
        d["hy_%d" % rows] = y
        d["hn_%d" % rows] = cnt
    # flush + tiny magnitudes + an all-negative block
    W = weights(64, 32, 300)
    W[:16] = -W[:16].abs()
    W[16:32] *= 1e-38
    y = quantize_mx_outlier_v1(W, 8, 8, "fp4_e2m1", "fp8_e4m3", "max", 2, [0], 16, "nearest", True, False)
    d["x_flush"] = W
    d["y_flush"] = y
    save("outlier.npz", **d)
    json.dump(meta, open(os.path.join(HERE, "outlier_cases.json"), "w"), indent=1)


def make_mxops():
    d = {}
    for n, (f, bs, sb, shape) in enumerate([("fp4_e2m1", 32, 4, (24, 256)), ("fp6_e3m2", 32, 4, (16, 128)),
                                            ("fp4_e2m1", 32, 8, (24, 256)), ("int4", 16, 8, (8, 96))]):
        W = weights(shape[0], shape[1], 400 + n)
        y = _quantize_mx_outlier_v1(W, sb, sb, f, f, block_size=bs, axes=[1])
        d["x_%d" % n] = W
        d["y_%d" % n] = y
    save("mxops_outlier.npz", **d)


def make_plain():
    d = {}
    cases = [("fp4_e2m1", 32, 8, "nearest", 1), ("fp6_e3m2", 32, 8, "even", 1), ("int8", 64, 8, "floor", 0),
             ("fp8_e4m3", 9, 5, "nearest", 1), ("int4", 16, 8, "nearest", 0), ("fp6_e2m3", 32, 4, "nearest", 1)]
    for n, (f, bs, sb, rnd, ax) in enumerate(cases):
        W = weights(64, 96, 500 + n)
        W[0, :4] = torch.tensor([0.0, -0.0, 1e-40, 3e38])
        d["x_%d" % n] = W
        d["y_%d" % n] = _quantize_mx(W, sb, f, axes=[ax], block_size=bs, round=rnd)
    json.dump([dict(fmt=c[0], bs=c[1], sb=c[2], round=c[3], axis=c[4]) for c in cases],
              open(os.path.join(HERE, "plain_cases.json"), "w"), indent=1)
    save("plain_mx.npz", **d)


def make_elemwise():
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4096, generator=g) * 3
    x[:8] = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), 1e-40, -1e-40, 3.3e38, 65504.0])
    d = {"x": x}
    for rnd in ("nearest", "floor", "even"):
        d["bf16_" + rnd] = _quantize_bfloat(x, 16, rnd)
        d["bf10_" + rnd] = _quantize_bfloat(x, 10, rnd)
        for f in ("fp8_e4m3", "fp8_e5m2", "fp6_e3m2", "fp6_e2m3", "fp4_e2m1", "int8", "int4", "int2"):
            eb, mb, _, mn, _ = _get_format_params(f)
            d["%s_%s_sat" % (f, rnd)] = _quantize_elemwise_core(x, mb, eb, mn, rnd, True, True)
            d["%s_%s_nosat_nodenorm" % (f, rnd)] = _quantize_elemwise_core(x, mb, eb, mn, rnd, False, False)
    save("elemwise.npz", **d)


def make_uniform():
    d = {}
    W = weights(48, 80, 600)
    d["w"] = W
    for n, (bits, perch, sym, mse) in enumerate([(4, True, False, False), (4, True, True, False), (3, False, True, False),
                                                 (2, True, False, True), (8, True, False, False)]):
        q = Quantizer()
        q.configure(bits, perchannel=perch, sym=sym, mse=mse)
        q.find_params(W, weight=True)
        d["scale_%d" % n] = q.scale
        d["zero_%d" % n] = q.zero
        d["q_%d" % n] = q.quantize(W)
    q = Quantizer()
    q.configure(2, perchannel=True, trits=True)
    q.find_params(W, weight=True)
    d["scale_trits"], d["zero_trits"], d["q_trits"] = q.scale, q.zero, q.quantize(W)
    x3 = torch.randn(2, 5, 80, generator=torch.Generator().manual_seed(601))
    q = Quantizer()
    q.configure(4, perchannel=True, sym=False)
    q.find_params(x3, weight=False)
    d["x3"], d["scale_x3"], d["zero_x3"], d["q_x3"] = x3, q.scale, q.zero, q.quantize(x3)
    save("uniform.npz", **d)


def make_gptq():
    d = {}
    cases = []
    for n, (rows, cols, ntok, blocksize, fi, fo, bs) in enumerate([
            (64, 96, 256, 128, "int2", "fp4", 16),       # single GPTQ block: no trailing GEMM
            (256, 160, 512, 64, "int2", "fp4", 16),      # 3 GPTQ blocks (64,64,32)
            (128, 128, 512, 128, "fp4_e2m1", "fp8_e4m3", 16),
            (96, 64, 256, 32, "fp6_e3m2", "fp6_e3m2", 32),   # fp6 inliers: the prune changes values
    ]):
        g = torch.Generator().manual_seed(700 + n)
        layer = torch.nn.Linear(cols, rows, bias=False)
        layer.weight.data = torch.randn(rows, cols, generator=g) * 0.02
        X = torch.randn(4, ntok // 4, cols, generator=g).half().float()
        X[..., :4] *= 6.0
        gq = ref_gptq.GPTQ(layer)
        gq.quantizer = MXQuantizer()
        gq.quantizer.configure(8, 8, fi, fo, axes=[0], block_size=bs)
        W0 = layer.weight.data.clone()
        for b in range(4):
            gq.add_batch(X[b:b + 1], None)
        H = gq.H.clone()
        # Hinv exactly as gptq.py:98-104 builds it
        Hd = H.clone()
        damp = 0.01 * torch.mean(torch.diag(Hd))
        Hd[torch.arange(cols), torch.arange(cols)] += damp
        Hinv = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hd)), upper=True)
        import io
        import contextlib
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            gq.fasterquant(blocksize=blocksize, percdamp=0.01)
        err = float([l for l in buf.getvalue().splitlines() if l.startswith("error")][0].split()[1])
        d["W_%d" % n], d["X_%d" % n], d["H_%d" % n], d["Hinv_%d" % n] = W0, X, H, Hinv
        d["Q_%d" % n] = layer.weight.data.clone()
        d["err_%d" % n] = np.float64(err)
        cases.append(dict(rows=rows, cols=cols, blocksize=blocksize, in_fmt=fi, out_fmt=fo, bs=bs))
    json.dump(cases, open(os.path.join(HERE, "gptq_cases.json"), "w"), indent=1)
    save("gptq.npz", **d)


def make_linear():
    import mx
    from mx.linear import linear
    from mx.specs import finalize_mx_specs
    d = {}
    cases = [dict(w_elem_format="fp4_e2m1", a_elem_format=None, block_size=32),
             dict(w_elem_format="fp6_e3m2", a_elem_format="fp6_e3m2", block_size=32, scale_bits=4, bfloat=16),
             dict(w_elem_format="fp4_e2m1", a_elem_format="fp8_e4m3", block_size=32, scale_bits=8),
             # not packable by the B200 path -> dense fallback: the spec of llm/opt_direct.py:98-105 (block 128),
             # a small block, and the weight / the activations left unquantized
             dict(w_elem_format="fp4", a_elem_format="int4", block_size=128, custom_cuda=False, quantize_backprop=False),
             dict(w_elem_format="fp4_e2m1", a_elem_format="fp8_e4m3", block_size=16, scale_bits=8),
             dict(w_elem_format=None, a_elem_format="fp8_e4m3", block_size=32, scale_bits=8),
             dict(w_elem_format="int8", a_elem_format=None, block_size=64, scale_bits=8, bfloat=16),
             # W4A4: both operands FP4 -> tcgen05 kind::mxf4 on the B200 path
             dict(w_elem_format="fp4_e2m1", a_elem_format="fp4_e2m1", block_size=32, scale_bits=8)]
    for n, spec in enumerate(cases):
        g = torch.Generator().manual_seed(800 + n)
        x = torch.randn(64, 256, generator=g)
        w = torch.randn(96, 256, generator=g) * 0.05
        b = torch.randn(96, generator=g) * 0.1
        y = linear(x, w, b, finalize_mx_specs(dict(spec)))
        d["x_%d" % n], d["w_%d" % n], d["b_%d" % n], d["y_%d" % n] = x, w, b, y
    json.dump(cases, open(os.path.join(HERE, "linear_cases.json"), "w"), indent=1)
    save("linear.npz", **d)


if __name__ == "__main__":
    make_outlier()
    make_mxops()
    make_plain()
    make_elemwise()
    make_uniform()
    make_gptq()
    make_linear()
