"""Packed MicroScopiQ weights: the artefact the reference never materialises.

    p = pack_mx_outlier(W, 8, 8, "fp4_e2m1", "fp8_e4m3", axes=[0], block_size=16)
    W_hat = p.dequantize()            # == quantize_mx_outlier_v1(W, ...) bit for bit

Planes (all CUDA tensors, layout in DESIGN.md):
  codes   uint8   inlier-width slot per element (2/4/8 bits), linear element order
  bitmap  int32   1 bit per element, set = outlier
  e_in    int8    inlier shared exponent per block  (-128 = NaN scale)
  e_out   int8    outlier shared exponent per block
  hi      uint8   per 1024-block segment, compacted high bits of the outlier codes (when wider)
  seg_cnt int32   outliers per segment
"""
import ctypes as C

import torch

from . import _native as N
from . import _ops

_p, _ll, _i, _f = C.c_void_p, C.c_longlong, C.c_int, C.c_float
N.register("msq_packed_build_hi_index", _i, [_p, _ll, _ll, _ll, _i, _p, _p])
N.register("msq_unpack_mx", _i, [_p, _p, _p, _p, _p, _p, _p, _ll, _ll, _ll, _i, _i, _i, _i, _f, _i, _i, _i, _f, _i, _p])


class PackedMX(object):
    """container for the packed planes + the metadata needed to decode them"""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def nbytes(self):
        """bytes a consumer has to read: hi counts only the compacted prefix of every segment"""
        n = self.codes.numel() + self.e_in.numel()
        if not self.plain:
            n += self.bitmap.numel() * 4 + self.e_out.numel()
            if self.hi is not None:
                n += (int(self.seg_cnt.sum().item()) * self.hi_w + 7) // 8
        return n

    def outlier_mask(self):
        """bool tensor of the original shape (decoded from the bitmap)"""
        words = self.bitmap.view(torch.int32)
        bits = (words.unsqueeze(-1) >> torch.arange(32, device=words.device, dtype=torch.int32)) & 1
        n = 1
        for s in self.shape:
            n *= s
        return bits.reshape(-1)[:n].reshape(self.shape).bool()

    def hi_index(self):
        """rank index over the bitmap (msq_packed_build_hi_index): stream offset of every 8th block inside its segment.
        Gives the row-block GEMM random access into the compacted hi stream; built once, cached (and stored by
        save_packed as a derived plane)."""
        idx = getattr(self, "hi_off8", None)
        if idx is None:
            pre, L, post = _ops.geometry(self.shape, self.axis)
            nblk = (L + self.block_size - 1) // self.block_size
            nb = pre * nblk * post
            idx = torch.empty((nb + 7) // 8, dtype=torch.int32, device=self.codes.device)
            rc = N.lib().msq_packed_build_hi_index(N.ptr(self.bitmap), pre, L, post, self.block_size, N.ptr(idx),
                                                   N.stream_ptr(idx.device))
            N.check(rc, "msq_packed_build_hi_index")
            self.hi_off8 = idx
        return idx

    def dequantize(self, dtype=torch.float32):
        pre, L, post = _ops.geometry(self.shape, self.axis)
        fi = _ops.fmt4(self.in_fmt)
        fo = _ops.fmt4(self.out_fmt if not self.plain else self.in_fmt)
        dev = self.codes.device
        out32 = torch.empty(self.shape, dtype=torch.float32, device=dev) if dtype == torch.float32 else None
        out16 = torch.empty(self.shape, dtype=torch.bfloat16, device=dev) if dtype == torch.bfloat16 else None
        if out32 is None and out16 is None:
            raise TypeError("dequantize to float32 or bfloat16")
        rc = N.lib().msq_unpack_mx(N.ptr(self.codes), N.ptr(self.e_in), N.ptr(self.e_out), N.ptr(self.bitmap),
                                   N.ptr(self.hi), N.ptr(out32), N.ptr(out16), pre, L, post, self.block_size,
                                   fi[0], fi[1], fi[2], fi[3], fo[0], fo[1], fo[2], fo[3], int(self.plain),
                                   N.stream_ptr(dev))
        N.check(rc, "msq_unpack_mx")
        return out32 if out32 is not None else out16


N.register("msq_quantize_mx_act", C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p])


def _alloc_planes(x, pre, L, post, bs, fi, fo, plain):
    L_ = N.lib()
    dev = x.device
    nblk = (L + bs - 1) // bs
    nb = pre * nblk * post
    n = x.numel()
    wi = L_.msq_packed_code_width(fi[0], fi[1])
    hi_w = 0 if plain else L_.msq_packed_hi_width(fi[0], fi[1], fo[0], fo[1])
    nseg = (nb + L_.msq_seg_blocks() - 1) // L_.msq_seg_blocks()
    stride = int(L_.msq_packed_hi_seg_stride(bs, hi_w))
    planes = dict(
        codes=torch.empty((n * wi + 7) // 8 + 16, dtype=torch.uint8, device=dev)[: (n * wi + 7) // 8],
        e_in=torch.empty(nb, dtype=torch.int8, device=dev),
        e_out=None if plain else torch.empty(nb, dtype=torch.int8, device=dev),
        bitmap=None if plain else torch.empty((n + 31) // 32, dtype=torch.int32, device=dev),
        hi=torch.empty(max(nseg * stride, 16), dtype=torch.uint8, device=dev) if hi_w > 0 else None,
        seg_cnt=None if plain else torch.empty(nseg, dtype=torch.int32, device=dev),
    )
    return planes, wi, hi_w, stride, nseg


def pack_mx_outlier(A, inlier_scale_bits, outlier_scale_bits, inlier_elem_format, outlier_elem_format, std_dev=2,
                    axes=None, block_size=16, round="nearest", flush_fp32_subnorms=False, variant=0, want_deq=False,
                    check=True):
    """quantize + pack in one pass (msq_quantize_mx_outlier_f32 with the packed outputs)."""
    x = _ops._as_f32(A)
    axis = _ops._single_axis(axes, x.ndim)
    pre, L, post = _ops.geometry(x.shape, axis)
    bs = int(block_size)
    fi, fo = _ops.fmt4(inlier_elem_format), _ops.fmt4(outlier_elem_format)
    pl, wi, hi_w, stride, nseg = _alloc_planes(x, pre, L, post, bs, fi, fo, False)
    deq = torch.empty_like(x) if want_deq else None
    st = torch.zeros(1, dtype=torch.int32, device=x.device)
    lo = hi = None
    if variant == 1:
        lo, hi = _ops.across_block_bounds(x, axis, bs, std_dev)
    path = C.c_int(0)
    rc = N.lib().msq_quantize_mx_outlier_f32(
        N.ptr(x), N.ptr(deq), pre, L, post, bs, int(inlier_scale_bits), int(outlier_scale_bits),
        fi[0], fi[1], fi[2], fi[3], fo[0], fo[1], fo[2], fo[3], float(std_dev), int(bool(flush_fp32_subnorms)),
        _ops._rmode(round), int(variant), N.ptr(lo), N.ptr(hi),
        N.ptr(pl["codes"]), N.ptr(pl["e_in"]), N.ptr(pl["e_out"]), N.ptr(pl["bitmap"]), N.ptr(pl["hi"]),
        N.ptr(pl["seg_cnt"]), None, None, N.ptr(st), C.byref(path), N.stream_ptr(x.device))
    N.check(rc, "msq_quantize_mx_outlier_f32(pack)")
    if check:
        _ops.raise_on_status(int(st.item()))
    p = PackedMX(shape=tuple(x.shape), axis=axis, block_size=bs, in_fmt=inlier_elem_format, out_fmt=outlier_elem_format,
                 in_scale_bits=inlier_scale_bits, out_scale_bits=outlier_scale_bits, code_width=wi, hi_w=hi_w,
                 hi_seg_stride=stride, nseg=nseg, plain=False, path=path.value, status=st, **pl)
    return (p, deq) if want_deq else p


#: activations blocked along their last axis go through the dedicated kernel (csrc/msq_quant_act.cu): the input is read
#: in its own dtype, rounding + codes by the hardware's narrow-format conversions; planes bit-identical to the generic path
ACT_KERNEL = True
_ACT_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _pack_mx_act(A, scale_bits, elem_format, flush):
    from .mx.formats import _get_format_params
    fp = tuple(_get_format_params(elem_format)[:4])
    if fp not in ((2, 3, 2, 6.0), (4, 5, 8, 448.0)):
        return None
    x = A.detach()
    N.require_cuda(x, "input")
    if x.dtype not in _ACT_DTYPES or x.dim() < 1 or x.shape[-1] % 32 or not (1 <= int(scale_bits) <= 8) or x.numel() == 0:
        return None
    x = x.contiguous()
    K = x.shape[-1]
    rows = x.numel() // K
    if x.data_ptr() % 32 or (K * x.element_size()) % 32:
        return None
    is4 = fp[0] == 2
    cb = K // 2 if is4 else K
    dev = x.device
    codes = torch.empty(rows * cb + 16, dtype=torch.uint8, device=dev)[: rows * cb]
    e = torch.empty(rows * (K // 32), dtype=torch.int8, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = N.lib().msq_quantize_mx_act(N.ptr(x), _ACT_DTYPES[x.dtype], K, rows, K, int(is4), int(scale_bits), int(bool(flush)),
                                     N.ptr(codes), cb, N.ptr(e), K // 32, N.ptr(st), N.stream_ptr(dev))
    if rc == -2:            # MSQ_ERR_UNSUPPORTED: the generic path takes it
        return None
    N.check(rc, "msq_quantize_mx_act")
    nseg = (e.numel() + N.lib().msq_seg_blocks() - 1) // N.lib().msq_seg_blocks()
    return PackedMX(shape=tuple(x.shape), axis=x.dim() - 1, block_size=32, in_fmt=elem_format, out_fmt=elem_format,
                    in_scale_bits=scale_bits, out_scale_bits=scale_bits, code_width=4 if is4 else 8, hi_w=0, hi_seg_stride=0,
                    nseg=nseg, plain=True, path=7, status=st, codes=codes, e_in=e, e_out=None, bitmap=None, hi=None,
                    seg_cnt=None)


def pack_mx(A, scale_bits, elem_format, axes=None, block_size=32, round="nearest", flush_fp32_subnorms=False, mode=1):
    """plain MX quantize + pack (codes + shared exponents); mode 1 = spec arithmetic (exact 2^e divisor)."""
    if (ACT_KERNEL and mode == 1 and int(block_size) == 32 and round == "nearest" and torch.is_tensor(A) and A.dim() >= 1
            and _ops._single_axis(axes, A.dim()) == A.dim() - 1):
        p = _pack_mx_act(A, scale_bits, elem_format, flush_fp32_subnorms)
        if p is not None:
            return p
    x = _ops._as_f32(A)
    axis = _ops._single_axis(axes, x.ndim)
    pre, L, post = _ops.geometry(x.shape, axis)
    bs = int(block_size)
    f = _ops.fmt4(elem_format)
    pl, wi, _, _, nseg = _alloc_planes(x, pre, L, post, bs, f, f, True)
    st = torch.zeros(1, dtype=torch.int32, device=x.device)
    path = C.c_int(0)
    rc = N.lib().msq_quantize_mx_f32(N.ptr(x), None, pre, L, post, bs, int(scale_bits), f[0], f[1], f[2], f[3],
                                     int(bool(flush_fp32_subnorms)), _ops._rmode(round), int(mode), None,
                                     N.ptr(pl["codes"]), N.ptr(pl["e_in"]), N.ptr(st), C.byref(path),
                                     N.stream_ptr(x.device))
    N.check(rc, "msq_quantize_mx_f32(pack)")
    return PackedMX(shape=tuple(x.shape), axis=axis, block_size=bs, in_fmt=elem_format, out_fmt=elem_format,
                    in_scale_bits=scale_bits, out_scale_bits=scale_bits, code_width=wi, hi_w=0, hi_seg_stride=0,
                    nseg=nseg, plain=True, path=path.value, status=st, **pl)


# ---------------------------------------------------------------------------------------------------------------
# On-disk format (replaces the reference's dead opt_pack3 / load_quant3, llm/opt.py:254-294)
#
#   bytes 0..7     magic  b"MSQPACK1"
#   bytes 8..15    little-endian uint64: length H of the JSON header
#   bytes 16..16+H JSON (utf-8): {"version": 1, "metadata": {...}, "tensors": {name: entry}}
#   then the raw planes, each starting on a 64-byte boundary (offsets in the header are relative to the first one)
#
# entry of a packed tensor: {"kind": "packed", "meta": {shape, axis, block_size, in_fmt, out_fmt, in_scale_bits,
#   out_scale_bits, code_width, hi_w, hi_seg_stride, nseg, plain}, "planes": {plane: {dtype, shape, offset, nbytes}}}
# entry of a dense tensor (biases, norms, embeddings ...): {"kind": "dense", "planes": {"data": {...}}}
# Only the used prefix layout is NOT compacted: the hi plane is stored with its segment stride, as in HBM, so a
# loaded tensor is usable by the kernels without any re-layout.
# ---------------------------------------------------------------------------------------------------------------
import json as _json
import struct as _struct

MAGIC = b"MSQPACK1"
FORMAT_VERSION = 1
_PLANES = ("codes", "bitmap", "e_in", "e_out", "hi", "seg_cnt", "hi_off8")
_META = ("shape", "axis", "block_size", "in_fmt", "out_fmt", "in_scale_bits", "out_scale_bits", "code_width", "hi_w",
         "hi_seg_stride", "nseg", "plain")
_DTYPES = {"uint8": torch.uint8, "int8": torch.int8, "int32": torch.int32, "float32": torch.float32,
           "float16": torch.float16, "bfloat16": torch.bfloat16, "int64": torch.int64}


def _fmt_name(f):
    return f if isinstance(f, str) or f is None else getattr(f, "name", str(f))


def _plane_entry(t, blobs, cursor):
    t = t.detach().contiguous().cpu()
    name = str(t.dtype).replace("torch.", "")
    if name not in _DTYPES:
        raise TypeError("cannot store dtype %s" % t.dtype)
    raw = t.view(torch.uint8).reshape(-1) if t.numel() else torch.empty(0, dtype=torch.uint8)
    if t.dtype in (torch.bfloat16,):
        raw = t.view(torch.int16).reshape(-1).view(torch.uint8)
    off = (cursor + 63) // 64 * 64
    blobs.append((off, raw))
    return {"dtype": name, "shape": list(t.shape), "offset": off, "nbytes": int(raw.numel())}, off + int(raw.numel())


def save_packed(path, tensors, metadata=None):
    """write {name: PackedMX | torch.Tensor} to one file; returns the number of bytes written"""
    entries, blobs, cursor = {}, [], 0
    for name, t in tensors.items():
        if isinstance(t, PackedMX):
            meta = {k: getattr(t, k) for k in _META}
            meta["shape"] = list(meta["shape"])
            meta["in_fmt"], meta["out_fmt"] = _fmt_name(meta["in_fmt"]), _fmt_name(meta["out_fmt"])
            planes = {}
            for pn in _PLANES:
                pl = getattr(t, pn, None)
                if pl is not None:
                    planes[pn], cursor = _plane_entry(pl, blobs, cursor)
            entries[name] = {"kind": "packed", "meta": meta, "planes": planes}
        else:
            e, cursor = _plane_entry(t, blobs, cursor)
            entries[name] = {"kind": "dense", "planes": {"data": e}}
    header = _json.dumps({"version": FORMAT_VERSION, "metadata": metadata or {}, "tensors": entries}).encode("utf-8")
    base = (16 + len(header) + 63) // 64 * 64
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(_struct.pack("<Q", len(header)))
        f.write(header)
        f.write(b"\0" * (base - 16 - len(header)))
        pos = 0
        for off, raw in blobs:
            f.write(b"\0" * (off - pos))
            f.write(raw.numpy().tobytes())
            pos = off + raw.numel()
    return base + cursor


def read_packed_header(path):
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("%s is not an MSQPACK1 file" % path)
        (hlen,) = _struct.unpack("<Q", f.read(8))
        header = _json.loads(f.read(hlen).decode("utf-8"))
    if header.get("version") != FORMAT_VERSION:
        raise ValueError("unsupported MSQPACK version %r" % (header.get("version"),))
    return header, (16 + hlen + 63) // 64 * 64


def load_packed(path, device="cuda", names=None):
    """read a file written by save_packed: {name: PackedMX | torch.Tensor} with the planes on `device`"""
    import numpy as np
    header, base = read_packed_header(path)
    mm = np.memmap(path, dtype=np.uint8, mode="r")

    def plane(e):
        raw = torch.from_numpy(np.array(mm[base + e["offset"]: base + e["offset"] + e["nbytes"]]))
        dt = _DTYPES[e["dtype"]]
        t = (raw.view(torch.int16).view(torch.bfloat16) if dt == torch.bfloat16 else raw.view(dt)) if raw.numel() else \
            torch.empty(0, dtype=dt)
        return t.reshape(e["shape"]).to(device)

    out = {}
    for name, e in header["tensors"].items():
        if names is not None and name not in names:
            continue
        if e["kind"] == "dense":
            out[name] = plane(e["planes"]["data"])
            continue
        kw = dict(e["meta"])
        kw["shape"] = tuple(kw["shape"])
        for pn in _PLANES:
            kw[pn] = plane(e["planes"][pn]) if pn in e["planes"] else None
        kw["path"], kw["status"] = 0, None
        out[name] = PackedMX(**kw)
    return out
