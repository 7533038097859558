"""tcgen05 GEMM / Hessian entry points of the C ABI (csrc/msq_gemm.cu)."""
import ctypes as C

import torch

from . import _native as N

_p, _ll, _i, _f = C.c_void_p, C.c_longlong, C.c_int, C.c_float
N.register("msq_gemm_nt_f32out", _i, [_p, _p, _p, _i, _i, _i, _ll, _ll, _ll, _f, _f, _i, _p])
N.register("msq_hessian_workspace_bytes", _ll, [_i, _i, _i])
N.register("msq_hessian_accumulate", _i, [_p, _i, _i, _ll, _i, _p, _ll, _f, _f, _p, _ll, _i, _p])

_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def gemm_nt(A, B, out=None, alpha=1.0, beta=0.0):
    """out[M,N] (fp32) = alpha * A[M,K] @ B[N,K].T + beta * out ; A, B fp16 or bf16, last dim contiguous"""
    N.require_cuda(A, "A")
    N.require_cuda(B, "B")
    assert A.dtype == B.dtype and A.dtype in (torch.float16, torch.bfloat16)
    assert A.dim() == 2 and B.dim() == 2 and A.shape[1] == B.shape[1]
    assert A.stride(1) == 1 and B.stride(1) == 1
    M, K = A.shape
    Nn = B.shape[0]
    if out is None:
        out = torch.empty(M, Nn, dtype=torch.float32, device=A.device)
        beta = 0.0
    assert out.dtype == torch.float32 and out.stride(1) == 1
    rc = N.lib().msq_gemm_nt_f32out(N.ptr(A), N.ptr(B), N.ptr(out), M, Nn, K, A.stride(0), B.stride(0), out.stride(0),
                                    float(alpha), float(beta), _DT[A.dtype], N.stream_ptr(A.device))
    N.check(rc, "msq_gemm_nt_f32out")
    return out


_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def hessian_accumulate(H, X, alpha, beta, symmetric=True):
    """H <- beta*H + alpha * X^T X  (X [T, K] tokens x features, fp32/fp16/bf16; H [K, K] fp32)"""
    N.require_cuda(H, "H")
    N.require_cuda(X, "X")
    assert X.dim() == 2 and X.stride(1) == 1 and X.dtype in _DT
    T, K = X.shape
    assert H.shape == (K, K) and H.dtype == torch.float32 and H.stride(1) == 1
    dt = _DT[X.dtype]
    nbytes = int(N.lib().msq_hessian_workspace_bytes(T, K, dt))
    ws = _workspace(nbytes, X.device)
    rc = N.lib().msq_hessian_accumulate(N.ptr(X), T, K, X.stride(0), dt, N.ptr(H), H.stride(0), float(alpha), float(beta),
                                        N.ptr(ws), nbytes, int(bool(symmetric)), N.stream_ptr(X.device))
    N.check(rc, "msq_hessian_accumulate")
    return H


N.register("msq_split3_planes_f32", _i, [_p, _ll, _i, _i, _p, _ll, _i, _p])
N.register("msq_gemm_split3_f32", _i, [_p, _ll, _i, _p, _ll, _i, _p, _ll, _i, _i, _i, _f, _f, _p])
N.register("msq_cholinv_workspace_bytes", _ll, [_i])
N.register("msq_cholinv_upper_f32", _i, [_p, _ll, _p, _ll, _i, _p, _ll, _p, _p])
N.register("msq_cholinv_ut_planes", _i, [_p, _i, C.POINTER(_p), C.POINTER(_ll), C.POINTER(_i)])


class Planes(object):
    """three bf16 planes whose sum is an fp32 matrix exactly (operand format of the split-plane tensor-core GEMM):
    plane p of element (r, c) at data[(p * plane_rows + r) * ld + c]"""

    def __init__(self, data, ld, plane_rows, rows, cols, keepalive=None, ptr=None):
        self.data, self.ld, self.plane_rows, self.rows, self.cols = data, ld, plane_rows, rows, cols
        self.keepalive = keepalive
        self._ptr = ptr

    @property
    def ptr(self):
        return C.c_void_p(self._ptr if self._ptr is not None else self.data.data_ptr())


def split3_planes(x):
    """fp32 [rows, cols] CUDA matrix -> Planes (msq_split3_planes_f32)"""
    N.require_cuda(x, "x")
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    rows, cols = x.shape
    ld = (cols + 63) // 64 * 64
    pr = (rows + 127) // 128 * 128
    data = torch.empty(3 * pr * ld, dtype=torch.bfloat16, device=x.device)
    rc = N.lib().msq_split3_planes_f32(N.ptr(x), x.stride(0), rows, cols, N.ptr(data), ld, pr, N.stream_ptr(x.device))
    N.check(rc, "msq_split3_planes_f32")
    return Planes(data, ld, pr, rows, cols)


def gemm_split3(A, B, out=None, alpha=1.0, beta=0.0):
    """out[M,N] (fp32) = alpha * A[M,K] @ B[N,K].T + beta * out with fp32-level accuracy on tcgen05.
    A, B: fp32 CUDA matrices or Planes (msq_gemm_split3_f32)."""
    pa = A if isinstance(A, Planes) else split3_planes(A)
    pb = B if isinstance(B, Planes) else split3_planes(B)
    assert pa.cols == pb.cols
    M, Nn, K = pa.rows, pb.rows, pa.cols
    dev = pa.data.device
    if out is None:
        out = torch.empty(M, (Nn + 3) // 4 * 4, dtype=torch.float32, device=dev)[:, :Nn]
        beta = 0.0
    assert out.dtype == torch.float32 and out.stride(1) == 1 and out.shape == (M, Nn)
    rc = N.lib().msq_gemm_split3_f32(pa.ptr, pa.ld, pa.plane_rows, pb.ptr, pb.ld, pb.plane_rows, N.ptr(out), out.stride(0),
                                     M, Nn, K, float(alpha), float(beta), N.stream_ptr(dev))
    N.check(rc, "msq_gemm_split3_f32")
    return out


class CholInv(object):
    """result of cholinv_upper: U (fp32 [n, n], upper, U^T U = H^-1), the bf16 planes of U^T that the routine leaves
    in its workspace (operand of the tensor-core trailing update) and the device info word"""

    def __init__(self, U, ut, info, ws):
        self.U, self.ut, self.info, self.ws = U, ut, info, ws


def cholinv_upper(H, check=True):
    """llm/gptq.py:101-103 (cholesky -> cholesky_inverse -> cholesky(upper)) as one native routine
    (msq_cholinv_upper_f32).  H: fp32 [n, n] SPD CUDA matrix, DESTROYED.  check=True reads the info word back
    (one host sync) and raises torch.linalg.LinAlgError like torch.linalg.cholesky does."""
    N.require_cuda(H, "H")
    assert H.dim() == 2 and H.shape[0] == H.shape[1] and H.dtype == torch.float32
    n = H.shape[0]
    if H.stride(1) != 1 or H.stride(0) % 4 or H.data_ptr() % 16:
        Hp = torch.zeros(n, (n + 3) // 4 * 4, dtype=torch.float32, device=H.device)
        Hp[:, :n] = H
        H = Hp[:, :n]
    dev = H.device
    U = torch.empty(n, (n + 3) // 4 * 4, dtype=torch.float32, device=dev)[:, :n]
    nbytes = int(N.lib().msq_cholinv_workspace_bytes(n))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = N.lib().msq_cholinv_upper_f32(N.ptr(H), H.stride(0), N.ptr(U), U.stride(0), n, N.ptr(ws), nbytes, N.ptr(info),
                                       N.stream_ptr(dev))
    N.check(rc, "msq_cholinv_upper_f32")
    pp, ld, pr = _p(), _ll(), _i()
    N.check(N.lib().msq_cholinv_ut_planes(N.ptr(ws), n, C.byref(pp), C.byref(ld), C.byref(pr)), "msq_cholinv_ut_planes")
    ut = Planes(ws, ld.value, pr.value, n, n, keepalive=ws, ptr=pp.value)
    if check:
        bad = int(info.item())
        if bad:
            raise torch.linalg.LinAlgError("cholinv_upper: the leading minor ending at %d is not positive-definite" % bad)
    return CholInv(U, ut, info, ws)


N.register("msq_gemm_packed_w_bf16", _i, [_p, _ll, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p])


#: activation rows at and above which packed_linear dequantizes the weight to bf16 first
FUSED_MAX_M = 512


def _bf16_terms(x):
    """split x into bf16 pieces whose sum is x exactly (bf16: 1, fp16: 2, fp32: 3 terms)"""
    if x.dtype == torch.bfloat16:
        return [x]
    xf = x.float()
    terms = []
    for _ in range(2 if x.dtype == torch.float16 else 3):
        t = xf.to(torch.bfloat16)
        terms.append(t)
        xf = xf - t.float()
    return terms


N.register("msq_gemm_packed_rowblock_bf16", _i, [_p, _ll, _p, _p, _p, _p, _p, _p, _ll, _p, _ll, _i, _i, _i, _i, _i, _i,
                                                  _i, _i, _i, _p])


def rowblock_ok(pw):
    """can this packed weight go through the row-block GEMM (the layout the quantizers emit: axes=[0])?"""
    return (len(pw.shape) == 2 and pw.axis == 0 and not pw.plain and pw.block_size in (16, 32)
            and pw.shape[0] % pw.block_size == 0 and pw.shape[1] % 64 == 0 and pw.code_width in (2, 4)
            and pw.hi_w in (0, 2, 4))


def _packed_linear_rowblock(x2, pw, out):
    from . import _ops
    Nn, K = pw.shape
    M = x2.shape[0]
    fi, fo = _ops.fmt4(pw.in_fmt), _ops.fmt4(pw.out_fmt)
    idx = pw.hi_index() if pw.hi_w else None
    for i, t in enumerate(_bf16_terms(x2)):
        t = t.contiguous()
        rc = N.lib().msq_gemm_packed_rowblock_bf16(N.ptr(t), t.stride(0), N.ptr(pw.codes), N.ptr(pw.bitmap), N.ptr(pw.e_in),
                                                   N.ptr(pw.e_out), N.ptr(pw.hi), N.ptr(idx), int(pw.hi_seg_stride),
                                                   N.ptr(out), out.stride(0), M, Nn, K, int(pw.block_size), fi[0], fi[1],
                                                   fo[0], fo[1], int(i > 0), N.stream_ptr(x2.device))
        N.check(rc, "msq_gemm_packed_rowblock_bf16")
    return out


N.register("msq_gemm_mxfp4w_tiled_folded", _i, [_p, _ll, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p])
N.register("msq_split_a16_e4m3", _i, [_p, _ll, _i, _i, _i, _i, _p, _p, _p])

#: 16-bit activations against a K-blocked FP4 weight below FUSED_MAX_M rows: two / three block-scaled E4M3 terms on the
#: block-scaled tensor-core path (the packed codes are the operand) instead of the producer-dequantizing bf16 kernel
SPLIT_A16 = True
#: ... and from FUSED_MAX_M rows as well when the weight has (next to) no outliers: the packed codes stay the only copy of
#: the weight (the bf16 route keeps 2 B per weight next to the 0.56-byte planes).  Measured 4096^3: 0.98 PFLOP/s against
#: 1.22 PFLOP/s for bf16 copy + kind::f16 GEMM; set False to prefer the copy.
LARGE_M_WITHOUT_BF16_COPY = True


class _SplitActs(object):
    """the stacked E4M3 terms of a 16-bit activation matrix, shaped like the PackedMX that packed_linear_mx8 takes"""
    block_size, code_width, in_fmt, plain = 32, 8, "fp8_e4m3", True

    def __init__(self, codes, e_in, shape):
        self.codes, self.e_in, self.shape = codes, e_in, shape


#: E4M3 terms per 16-bit activation: 2 hold bf16's 8 significant bits (exact within 2^12 of the block maximum, truncation
#: at 2^-19 of it below), fp16's 11 bits need 3
SPLIT_A16_TERMS = {torch.bfloat16: 2, torch.float16: 3}
#: measured (8192^2, bf16): M = 64: stacked 27.1 us / accumulated 31.2 us; M = 256: 49.7 / 41.6; M = 511: 90.7 / 62.1
SPLIT_A16_ACCUMULATE_MIN_M = 128


def split_a16(x2, terms=None):
    """x2 [M, K] bf16 / fp16 -> _SplitActs with terms * M rows (msq_split_a16_e4m3)"""
    M, K = x2.shape
    T = int(terms or SPLIT_A16_TERMS[x2.dtype])
    x2 = x2 if x2.stride(1) == 1 and x2.stride(0) % 8 == 0 and x2.data_ptr() % 16 == 0 else x2.contiguous()
    codes = torch.empty(T * M, K, dtype=torch.uint8, device=x2.device)
    e = torch.empty(T * M, K // 32, dtype=torch.int8, device=x2.device)
    rc = N.lib().msq_split_a16_e4m3(N.ptr(x2), x2.stride(0), _DT[x2.dtype], M, K, T, N.ptr(codes), N.ptr(e),
                                    N.stream_ptr(x2.device))
    N.check(rc, "msq_split_a16_e4m3")
    sa = _SplitActs(codes, e, (T * M, K))
    sa.terms = T
    return sa


#: the quantizers' own artefact (row blocks, FP8 outliers in a hi stream) below FUSED_MAX_M rows: re-encode the weight
#: ONCE, exactly, as K-blocked MX-FP8 (E4M3 bytes + UE8M0 scales, 1.03 B/weight next to the artefact's 0.94) and run the
#: block-scaled kernels on it, instead of the producer-bound row-block kernel
ROWBLOCK_VIA_MXFP8 = True


def rowblock_mxfp8_operand(pw):
    """exact K-blocked MX-FP8 form of ANY packed weight [N, K] (K % 32 == 0): (E4M3 bytes [N, K], UE8M0 scales [N, K/32]) with
    bytes * 2^(scale - 127) (+ pw._mx8_dense_resid: COO remainders of the few elements that are more than ~14 binades
    below their block's maximum) == pw.dequantize() bit for bit, or None when more than max(64, N) elements are not
    representable.  Built once (one device-to-host read for the exactness check), kept on the PackedMX as the operand of
    the block-scaled GEMMs."""
    hit = getattr(pw, "_mx8_dense", None)
    if hit is not None:
        return hit
    if getattr(pw, "_mx8_dense_failed", False):
        return None
    Nn, K = pw.shape
    if K % 32:
        return None
    w = pw.dequantize()                                              # fp32, exact values of the artefact
    blk = w.view(Nn, K // 32, 32)
    amax = blk.abs().amax(-1, keepdim=True)
    # exponent of the block maximum from its bits (frexp) and exact powers of two from bit patterns: the device log2 / exp2
    # are not exact at the integers.  The block maximum lands in [2^7, 2^8).
    e = (torch.frexp(amax.clamp_min(2.0 ** -100))[1] - 1 - 7).clamp_(-126, 120).to(torch.int32)
    p2 = lambda t: ((t + 127) << 23).view(torch.float32)
    q8 = (blk * p2(-e)).to(torch.float8_e4m3fn)
    back = q8.float() * p2(e)
    bad = (back != blk).view(Nn, K)
    nbad = int(bad.sum())                                            # the one device-to-host read
    if nbad > max(64, Nn):
        pw._mx8_dense_failed = True
        return None
    pw._mx8_dense_resid = None
    if nbad:
        # a handful of elements far below their block's maximum (measured: 7 of 16.7 M on a 4096^2 Gaussian weight, FP8
        # outliers ~2^18 under the block maximum): their remainders ride along as a sparse correction
        idx = bad.nonzero()
        dw = (w - back.view(Nn, K))[bad].contiguous()
        pw._mx8_dense_resid = (idx[:, 0].to(torch.int32).contiguous(), idx[:, 1].to(torch.int32).contiguous(), dw)
    wb = q8.view(torch.uint8).reshape(Nn, K).contiguous()
    wsf = (e + 127).to(torch.uint8).reshape(Nn, K // 32).contiguous()
    pw._mx8_dense = (wb, wsf)
    return pw._mx8_dense


def packed_linear(x, pw, bias=None):
    """y[..., N] (fp32) = x[..., K] @ dequant(pw)[N, K].T (+ bias) with pw a PackedMX blocked along K
    (block 32).  bf16 activations take one pass; fp16 / fp32 activations are split into exact bf16
    terms (2 / 3 accumulating passes) so that every product stays exact."""
    N.require_cuda(x, "x")
    from . import _ops
    if len(pw.shape) == 2 and pw.axis == 0:
        # the quantizers' own artefact (row blocks along the output features, hi stream): row-block GEMM, or -- for
        # geometries it does not cover -- unpack once to bf16 (exact) and the dense tcgen05 GEMM
        Nn, K = pw.shape
        lead = x.shape[:-1]
        x2 = x.reshape(-1, K)
        out = torch.empty(x2.shape[0], (Nn + 3) // 4 * 4, dtype=torch.float32, device=x.device)[:, :Nn]
        M = x2.shape[0]
        if (ROWBLOCK_VIA_MXFP8 and SPLIT_A16 and x2.dtype in SPLIT_A16_TERMS and M < FUSED_MAX_M and K % 128 == 0 and Nn % 4 == 0
                and rowblock_mxfp8_operand(pw) is not None):
            # 16-bit activations as block-scaled E4M3 terms x the exact MX-FP8 form of the artefact: block-scaled kernels
            sa = split_a16(x2)
            y3 = packed_linear_mx8(sa, pw, accumulate_terms=M >= SPLIT_A16_ACCUMULATE_MIN_M)
            folded = y3.shape[0] == M and sa.terms > 1
            if pw._mx8_dense_resid is not None:
                rn, rk, rdw = pw._mx8_dense_resid
                rc = N.lib().msq_sparse_outlier_correct(N.ptr(sa.codes), K, N.ptr(sa.e_in), N.ptr(rn), N.ptr(rk), N.ptr(rdw),
                                                        int(rk.numel()), N.ptr(y3), y3.stride(0), sa.shape[0], K,
                                                        M if folded else 0, N.stream_ptr(y3.device))
                N.check(rc, "msq_sparse_outlier_correct")
            out = y3 if folded else y3.view(sa.terms, M, Nn).sum(0)
        elif rowblock_ok(pw) and x2.shape[0] < FUSED_MAX_M:
            _packed_linear_rowblock(x2, pw, out)
        else:
            wb = getattr(pw, "_dense_bf16", None)
            if wb is None:
                wb = pw.dequantize(torch.bfloat16)
                if K % 8:
                    wb = torch.nn.functional.pad(wb, (0, 8 - K % 8))
                pw._dense_bf16 = wb if x2.shape[0] >= FUSED_MAX_M else None      # large-M callers reuse it
            for i, t in enumerate(_bf16_terms(x2)):
                t = t.contiguous() if K % 8 == 0 else torch.nn.functional.pad(t, (0, 8 - K % 8)).contiguous()
                gemm_nt(t, wb, out=out, alpha=1.0, beta=1.0 if i else 0.0)
        if bias is not None:
            out = out + bias.float()
        return out.reshape(*lead, Nn)
    if len(pw.shape) != 2 or pw.axis != 1 or pw.block_size != 32 or pw.hi_w != 0 or pw.shape[1] % 32:
        raise NotImplementedError("packed_linear needs a [N, K] tensor packed along K with block 32, K a multiple "
                                  "of 32 and equal code widths")
    Nn, K = pw.shape
    lead = x.shape[:-1]
    x2 = x.reshape(-1, K)
    M = x2.shape[0]
    out = torch.empty(M, Nn, dtype=torch.float32, device=x.device)
    fi, fo = _ops.fmt4(pw.in_fmt), _ops.fmt4(pw.out_fmt)
    split_ok = (SPLIT_A16 and x2.dtype in SPLIT_A16_TERMS and K % 128 == 0 and Nn % 4 == 0 and mx4_weight_ok(pw))
    if (M >= FUSED_MAX_M and Nn % 8 == 0
            and not (LARGE_M_WITHOUT_BF16_COPY and split_ok and getattr(pw, "_dense_bf16", None) is None
                     and mx8_tma_weight(pw, allow_copy=False, sparse_limit=MX8_TMA_SPARSE_BUDGET_LARGE_M / (float(M) * Nn))
                     is not None)):
        # compute-bound regime: dequantize the weight to bf16 (exact) ONCE per weight -- kept on the PackedMX
        # object, 2 B per weight next to the packed planes -- and run the dense GEMM
        wb = getattr(pw, "_dense_bf16", None)
        if wb is None:
            wb = pw._dense_bf16 = pw.dequantize(torch.bfloat16)
        for i, t in enumerate(_bf16_terms(x2)):
            gemm_nt(t.contiguous(), wb, out=out, alpha=1.0, beta=1.0 if i else 0.0)
        if bias is not None:
            out += bias.float()
        return out.reshape(*lead, Nn)
    if (split_ok
            and (SPLIT_A16_TERMS[x2.dtype] * M <= FP4_TMA_MAX_M or mx8_tma_weight(pw, allow_copy=False) is not None)):
        # W4A16 without dequantizing the weight: the activations become two / three block-scaled E4M3 terms stacked along
        # the token dimension, the packed codes feed the block-scaled tensor-core kernels, the slabs are added
        sa = split_a16(x2)
        if FP4_TMA and FP4_TILED and sa.shape[0] <= FP4_TMA_MAX_M:
            # decode sizes: the slabs are summed by the split-K atomics of the GEMM itself
            o_ptr, o_k, o_dw = mx4_sparse_outliers(pw)
            ct, st = mx4_tiled_weight(pw)
            out = torch.empty(M, Nn, dtype=torch.float32, device=x2.device)
            rc = N.lib().msq_gemm_mxfp4w_tiled_folded(N.ptr(sa.codes), K, N.ptr(sa.e_in), N.ptr(ct), N.ptr(st), N.ptr(o_ptr),
                                                      N.ptr(o_k), N.ptr(o_dw), N.ptr(out), out.stride(0), sa.shape[0], M, Nn, K,
                                                      N.stream_ptr(out.device))
            N.check(rc, "msq_gemm_mxfp4w_tiled_folded")
        else:
            # from SPLIT_A16_ACCUMULATE_MIN_M rows the slabs are accumulated inside the TMA-fed GEMM (K loop over (term, K
            # block): no [terms * M, N] result, no pass that adds the slabs); below, stacked slabs give more tiles to fill the SMs
            y3 = packed_linear_mx8(sa, pw, accumulate_terms=M >= SPLIT_A16_ACCUMULATE_MIN_M)
            out = y3 if y3.shape[0] == M else y3.view(sa.terms, M, Nn).sum(0)
        if bias is not None:
            out += bias.float()
        return out.reshape(*lead, Nn)
    for i, t in enumerate(_bf16_terms(x2)):
        t = t.contiguous()
        rc = N.lib().msq_gemm_packed_w_bf16(N.ptr(t), t.stride(0), N.ptr(pw.codes), N.ptr(pw.bitmap), N.ptr(pw.e_in),
                                            N.ptr(pw.e_out), N.ptr(out), out.stride(0), M, Nn, K, fi[0], fi[1], fo[0],
                                            fo[1], int(pw.plain), int(i > 0), N.stream_ptr(x.device))
        N.check(rc, "msq_gemm_packed_w_bf16")
    if bias is not None:
        out += bias.float()
    return out.reshape(*lead, Nn)


def dense_linear(x, w, bias=None):
    """y[..., N] (fp32) = x[..., K] @ w[N, K].T on tcgen05 for operands that have no packed form (the reference's
    `f_linear(qis_input, qis_weight)`, number_system/mx/linear.py:91, when the weight is not packable): both operands
    are split into three bf16 planes whose sum is the operand exactly and contracted by the split-plane GEMM
    (fp32-level accuracy; csrc/msq_gemm_split3.cu)."""
    N.require_cuda(x, "x")
    N.require_cuda(w, "w")
    K = x.shape[-1]
    assert w.dim() == 2 and w.shape[1] == K
    lead = x.shape[:-1]
    out = gemm_split3(x.reshape(-1, K).float().contiguous(), w.float().contiguous())
    if bias is not None:
        out = out + bias.float()
    return out.reshape(*lead, w.shape[0])


N.register("msq_gemm_mxfp8_blockscaled", _i, [_p, _ll, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _i, _p])
N.register("msq_widen_fp4_to_mxfp8", _i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p])
N.register("msq_gemm_mxfp8_dense", _i, [_p, _ll, _p, _p, _ll, _p, _p, _ll, _i, _i, _i, _i, _p])

#: activation rows at and above which the weight is widened once and streamed through TMA
MX8_DENSE_MIN_M = 512
#: decode sizes: feed the packed 4-bit codes to the tensor core directly (TMA 16U4_ALIGN16B) when the layout allows.
#: Measured (4096^2, kernel only): 9.6 / 12.5 us at M = 1 / 16 against 15.5 / 15.8 us for the producer-widen mode; the
#: per-token cost of this mode is 4.5x that of the TMA-fed E4M3 mode (cause not established), so it is used up to M = 32.
FP4_TMA = True
FP4_TMA_MAX_M = 32


def mx8_weight_ok(pw):
    """can this packed weight go through the native block-scaled FP8 path?  FP4 (E2M1) inliers and
    outliers, K blocks of 32, and every outlier block's exponent relative to the block scale,
    e_out - 2 e_in, within the range where q * 2^(e_out - 2 e_in) is still an exact E4M3 value.
    (One device->host read; the answer is cached on the PackedMX object.)"""
    from .mx.formats import _get_format_params
    ok = getattr(pw, "_mx8_ok", None)
    if ok is not None:
        return ok
    ok = (len(pw.shape) == 2 and pw.axis == 1 and pw.block_size == 32 and pw.hi_w == 0 and pw.code_width == 4
          and _get_format_params(pw.in_fmt)[:2] == (2, 3) and (pw.plain or _get_format_params(pw.out_fmt)[:2] == (2, 3)))
    if ok and not pw.plain:
        has_out = pw.bitmap != 0                       # block 32 along K == one bitmap word per block
        d = (pw.e_out.to(torch.int32) - 2 * pw.e_in.to(torch.int32))[has_out]
        ok = bool(d.numel() == 0 or (int(d.min()) >= -6 and int(d.max()) <= 6))
    pw._mx8_ok = ok
    return ok


N.register("msq_gemm_mxfp4w_blockscaled", _i, [_p, _ll, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p])


N.register("msq_tile_fp4_weight", _i, [_p, _p, _p, _p, _i, _i, _p])
N.register("msq_gemm_mxfp4w_tiled_blockscaled", _i, [_p, _ll, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p])

#: decode sizes read the weight from a tile-contiguous copy of the codes + scales (built once per weight, +0.53 B/weight
#: next to the planes; set False to stream the row-major planes themselves)
FP4_TILED = True


def mx4_tiled_weight(pw):
    """(codes_tiled, sf_tiled) of a K-blocked FP4 weight (msq_tile_fp4_weight), cached on the PackedMX"""
    t = getattr(pw, "_mx4_tiled", None)
    if t is None:
        Nn, K = pw.shape
        chunks = ((Nn + 127) // 128) * (K // 128)
        ct = torch.empty(chunks * 8192, dtype=torch.uint8, device=pw.codes.device)
        st = torch.empty(chunks * 512, dtype=torch.uint8, device=pw.codes.device)
        rc = N.lib().msq_tile_fp4_weight(N.ptr(pw.codes), N.ptr(pw.e_in), N.ptr(ct), N.ptr(st), Nn, K,
                                         N.stream_ptr(pw.codes.device))
        N.check(rc, "msq_tile_fp4_weight")
        t = pw._mx4_tiled = (ct, st)
    return t


def mx4_weight_ok(pw):
    """can the packed 4-bit codes of this weight feed the tensor core directly (FP4-by-TMA mode)?  FP4 (E2M1) inliers and
    outliers, K blocks of 32, K a multiple of 128 -- any exponents: the outliers are a sparse epilogue correction."""
    from .mx.formats import _get_format_params
    return (len(pw.shape) == 2 and pw.axis == 1 and pw.block_size == 32 and pw.hi_w == 0 and pw.code_width == 4
            and pw.shape[1] % 128 == 0 and _get_format_params(pw.in_fmt)[:2] == (2, 3)
            and (pw.plain or _get_format_params(pw.out_fmt)[:2] == (2, 3)) and pw.codes.data_ptr() % 32 == 0)


def mx4_sparse_outliers(pw):
    """CSR of the outlier corrections of a K-blocked FP4 weight (built once, cached on the PackedMX):
    o_dw = true value - the value the tensor core computes from the slot (code * 2^e_in)."""
    sp = getattr(pw, "_mx4_sparse", None)
    if sp is not None:
        return sp
    if pw.plain:
        sp = (None, None, None)
    else:
        from .packed import PackedMX
        Nn, K = pw.shape
        w_true = pw.dequantize()
        as_plain = PackedMX(**{**pw.__dict__, "plain": True})
        w_hw = as_plain.dequantize()
        mask = pw.outlier_mask()
        idx = mask.nonzero()                                          # row-major order: sorted by feature, then k
        dw = (w_true - w_hw)[mask].contiguous()
        o_ptr = torch.zeros(Nn + 1, dtype=torch.int32, device=w_true.device)
        if idx.numel():
            o_ptr[1:] = torch.cumsum(torch.bincount(idx[:, 0], minlength=Nn), 0).to(torch.int32)
        sp = (o_ptr, idx[:, 1].to(torch.int32).contiguous(), dw.float())
        pw._mx4_sparse_rows = idx[:, 0].to(torch.int32).contiguous()
        if not idx.numel():
            sp = (None, None, None)
    pw._mx4_sparse = sp
    return sp


def widen_weight_mx8(pw):
    """packed FP4(+outliers) weight -> dense MX-FP8 operand (E4M3 bytes [N, K], UE8M0 scales [N, K/32]),
    outliers folded in exactly as the fused kernel does (msq_widen_fp4_to_mxfp8)."""
    Nn, K = pw.shape
    dev = pw.codes.device
    wb = torch.empty(Nn, K, dtype=torch.uint8, device=dev)
    wsf = torch.empty(Nn, K // 32, dtype=torch.uint8, device=dev)
    rc = N.lib().msq_widen_fp4_to_mxfp8(N.ptr(pw.codes), N.ptr(pw.bitmap), N.ptr(pw.e_in), N.ptr(pw.e_out), N.ptr(wb),
                                        N.ptr(wsf), Nn, K, int(pw.plain), N.stream_ptr(dev))
    N.check(rc, "msq_widen_fp4_to_mxfp8")
    return wb, wsf


N.register("msq_sparse_outlier_correct", _i, [_p, _ll, _p, _p, _p, _p, _i, _p, _ll, _i, _i, _i, _p])
N.register("msq_tile_scales_ue8m0", _i, [_p, _i, _i, _i, _i, _p, _p])
N.register("msq_gemm_mxfp8_tma_tile_tokens", _i, [_i, _i, _i])
N.register("msq_gemm_mxfp8_tma_workspace_bytes", _ll, [_i, _i, _i])
N.register("msq_gemm_mxfp8_tma", _i, [_p, _ll, _p, _p, _ll, _i, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _i, _p, _ll, _p])

N.register("msq_sparse_outlier_correct_fp4", _i, [_p, _ll, _p, _p, _p, _p, _i, _p, _ll, _i, _i, _i, _p])
N.register("msq_gemm_mx_tma_k_slices", _i, [_i, _i, _i, _i])
N.register("msq_gemm_mxfp8_tma_bf16_workspace_bytes", _ll, [_i, _i, _i])
N.register("msq_gemm_mxfp8_tma_bf16", _i, [_p, _ll, _p, _p, _ll, _i, _p, _p, _ll, _i, _i, _i, _p, _ll, _p])
N.register("msq_gemm_mxfp4_tma_bf16_workspace_bytes", _ll, [_i, _i, _i])
N.register("msq_gemm_mxfp4_tma_bf16", _i, [_p, _ll, _p, _p, _ll, _p, _p, _ll, _i, _i, _i, _p, _ll, _p])
N.register("msq_gemm_mxfp4_tma_tile_tokens", _i, [_i, _i, _i])
N.register("msq_gemm_mxfp4_tma_workspace_bytes", _ll, [_i, _i, _i])
N.register("msq_gemm_mxfp4_tma", _i, [_p, _ll, _p, _p, _ll, _p, _p, _ll, _i, _i, _i, _i, _p, _ll, _p])


#: W4A4 (both operands FP4) runs on kind::mxf4 (msq_gemm_mxfp4_tma); False routes it like any other pairing
MX4_TMA = True


def mx4_pair_ok(pw, M=1):
    """can this weight be the FP4 operand of the kind::mxf4 kernel for M tokens?  K-blocked FP4 codes (outliers in FP4 as
    well: they become a sparse correction behind the GEMM while outliers x tokens stays within the budget of the FP8
    pairing), K % 256 == 0"""
    if not (MX4_TMA and mx4_weight_ok(pw) and pw.shape[1] % 256 == 0):
        return False
    if pw.plain:
        return True
    o_k = mx4_sparse_outliers(pw)[1]
    return o_k is None or o_k.numel() * float(M) <= MX8_TMA_SPARSE_BUDGET


def packed_linear_mx4(xq, pw, bias=None, assume_clean=False, out_dtype=torch.float32):
    """W4A4: y = dq(xq) @ dq(pw).T with BOTH operands as packed FP4 codes on tcgen05 kind::mxf4.block_scale (K = 64 per
    instruction, twice the rate of kind::mxf8f6f4).  xq: FP4 (E2M1) activations packed along K with block 32 and no
    outliers (pack_mx, or pack_mx_outlier with an empty bitmap -- checked here unless assume_clean); pw: an FP4 weight
    packed along K with block 32, no outliers, K % 256 == 0 (mx4_pair_ok).  Replaces F.linear on the two fake-quantized
    operands of number_system/mx/linear.py:66-91 for a_elem_format = w_elem_format = fp4_e2m1."""
    from . import _ops
    Nn, K = pw.shape
    if (xq.shape[-1] != K or xq.block_size != 32 or xq.code_width != 4 or xq.hi_w != 0 or xq.axis != len(xq.shape) - 1
            or _ops.fmt4(xq.in_fmt)[:2] != (2, 3)):
        raise NotImplementedError("activations must be MX-FP4 (E2M1) packed along K with block 32")
    if not xq.plain and not assume_clean and bool((xq.bitmap != 0).any()):
        raise NotImplementedError("activations with outliers of their own cannot feed kind::mxf4")
    M = 1
    for s_ in xq.shape[:-1]:
        M *= s_
    if not mx4_pair_ok(pw, M):
        raise NotImplementedError("weight must be an MX-FP4 weight packed along K with block 32 (few FP4 outliers), K % 256 == 0")
    wsf_t = getattr(pw, "_mx4_sf_tiled", None)
    if wsf_t is None:
        wsf_t = pw._mx4_sf_tiled = tile_scales(pw.e_in.reshape(Nn, K // 32))
    has_outliers = (not pw.plain) and mx4_sparse_outliers(pw)[1] is not None and mx4_sparse_outliers(pw)[1].numel() > 0
    if out_dtype == torch.bfloat16 and bias is None and not has_outliers:
        # the result leaves the epilogue as bf16 (rounded once from the fp32 accumulator): no conversion pass behind the GEMM
        out = torch.empty(M, Nn, dtype=torch.bfloat16, device=pw.codes.device)
        nbytes = int(N.lib().msq_gemm_mxfp4_tma_bf16_workspace_bytes(M, Nn, K))
        ws = _workspace(nbytes, out.device)
        rc = N.lib().msq_gemm_mxfp4_tma_bf16(N.ptr(xq.codes), K // 2, N.ptr(xq.e_in), N.ptr(pw.codes), K // 2, N.ptr(wsf_t),
                                             N.ptr(out), out.stride(0), M, Nn, K, N.ptr(ws), nbytes, N.stream_ptr(out.device))
        N.check(rc, "msq_gemm_mxfp4_tma_bf16")
        return out.reshape(*xq.shape[:-1], Nn)
    out = torch.empty(M, Nn, dtype=torch.float32, device=pw.codes.device)
    nbytes = int(N.lib().msq_gemm_mxfp4_tma_workspace_bytes(M, Nn, K))
    ws = _workspace(nbytes, out.device)
    rc = N.lib().msq_gemm_mxfp4_tma(N.ptr(xq.codes), K // 2, N.ptr(xq.e_in), N.ptr(pw.codes), K // 2, N.ptr(wsf_t), N.ptr(out),
                                    out.stride(0), M, Nn, K, 0, N.ptr(ws), nbytes, N.stream_ptr(out.device))
    N.check(rc, "msq_gemm_mxfp4_tma")
    if not pw.plain:
        o_ptr, o_k, o_dw = mx4_sparse_outliers(pw)
        if o_k is not None and o_k.numel():
            # the tensor core read every outlier slot as q * 2^e_in: add x[m, k] * (true value - that) per outlier and token
            rc = N.lib().msq_sparse_outlier_correct_fp4(N.ptr(xq.codes), K // 2, N.ptr(xq.e_in), N.ptr(pw._mx4_sparse_rows), N.ptr(o_k),
                                                        N.ptr(o_dw), int(o_k.numel()), N.ptr(out), out.stride(0), M, K, 0,
                                                        N.stream_ptr(out.device))
            N.check(rc, "msq_sparse_outlier_correct_fp4")
    if bias is not None:
        out += bias.float()
    return out.reshape(*xq.shape[:-1], Nn)


#: compute-bound sizes (M >= MX8_DENSE_MIN_M): fully TMA-fed kernel (csrc/msq_gemm_mx8_tma.cu)
MX8_TMA = True
#: ... with the packed 4-bit codes themselves as the operand when the weight has no outliers (half the L2 -> SMEM bytes)
MX8_TMA_FP4 = True


def tile_scales(scales, rows_per_tile=128):
    """[rows, K/32] int8 shared exponents (or uint8 UE8M0 bytes) -> tcgen05.cp chunk layout (msq_tile_scales_ue8m0)"""
    rows, kb32 = scales.shape
    K = kb32 * 32
    chunks = (rows_per_tile + 127) // 128
    out = torch.empty(((rows + rows_per_tile - 1) // rows_per_tile) * (K // 128) * chunks * 512, dtype=torch.uint8,
                      device=scales.device)
    rc = N.lib().msq_tile_scales_ue8m0(N.ptr(scales), int(scales.dtype == torch.int8), rows, K, rows_per_tile, N.ptr(out),
                                       N.stream_ptr(scales.device))
    N.check(rc, "msq_tile_scales_ue8m0")
    return out


#: a weight WITH outliers still feeds its packed codes to the TMA-fed kernel; the outliers are a sparse correction
#: C[m, n_e] += x[m, k_e] * dw_e applied by a pass of its own behind the GEMM (msq_sparse_outlier_correct: one thread per
#: (outlier, token), every SM works on it) while outliers x tokens stays below this budget; beyond it the widened E4M3
#: copy (compute-bound sizes) or the producer-widen kernel.  MX8_TMA_FUSED_SPARSE = True applies the correction inside the
#: GEMM's epilogue instead (four warps per CTA gather: latency-bound, measured 8192^2, 0.85 outliers per row, M = 256:
#: 63.8 us against 22.4 us for the weight without outliers).
MX8_TMA_SPARSE_BUDGET = 4.0e6
MX8_TMA_SPARSE_BUDGET_LARGE_M = 2.5e5        # from MX8_DENSE_MIN_M rows the widened copy is the alternative: a few microseconds
MX8_TMA_FUSED_SPARSE = False
MX8_TMA_SPARSE_MAX_PER_ROW = 1.0            # used when the token count is not known to the caller


def mx8_tma_weight(pw, allow_copy=True, sparse_limit=None):
    """operand of the TMA-fed kernel, built once per weight and kept on the PackedMX:
    (bytes, row pitch, is_fp4, tiled scales, o_ptr, o_k, o_dw) -- the packed codes themselves for an FP4 weight without
    or with few outliers (nothing is copied but the scales; outliers = CSR correction in the epilogue), else, when
    allow_copy, the widened E4M3 copy with the outliers folded in (1 byte per weight; needs mx8_weight_ok; shared with
    cache_dense=True); None when neither applies"""
    Nn, K = pw.shape
    if sparse_limit is None:
        sparse_limit = MX8_TMA_SPARSE_MAX_PER_ROW
    if MX8_TMA_FP4 and mx4_weight_ok(pw) and getattr(pw, "_mx8_dense", None) is None:
        op = getattr(pw, "_mx8_tma_fp4", None)
        if op is None:
            sp = mx4_sparse_outliers(pw)
            op = pw._mx8_tma_fp4 = (pw.codes, K // 2, 1, tile_scales(pw.e_in.reshape(Nn, K // 32))) + tuple(sp)
        nnz = 0 if op[5] is None else op[5].numel()
        if nnz <= sparse_limit * Nn or not mx8_weight_ok(pw):
            return op                  # few outliers -- or no other kernel can take this weight's exponents
        if not allow_copy:
            # below the compute-bound sizes no copy is made: the sparse correction up to 4x the budget, beyond that the
            # producer-widen kernel (whose cost does not depend on the number of outliers)
            return op if nnz <= 4 * sparse_limit * Nn else None
    dense = getattr(pw, "_mx8_dense", None)
    if dense is None:
        if not allow_copy or not mx8_weight_ok(pw):
            return None
        dense = pw._mx8_dense = widen_weight_mx8(pw)
    sf_t = getattr(pw, "_mx8_dense_sf_tiled", None)
    if sf_t is None or sf_t[0] is not dense[1]:
        sf_t = pw._mx8_dense_sf_tiled = (dense[1], tile_scales(dense[1]))
    return (dense[0], K, 0, sf_t[1], None, None, None)


def packed_linear_mx8(xq, pw, bias=None, cache_dense=False, accumulate_terms=False, out_dtype=torch.float32):
    """y = dq(xq) @ dq(pw).T on tcgen05 kind::mxf8f6f4.block_scale.  xq: PackedMX of the activations
    ([..., K], FP8 E4M3, block 32 along the last axis, no outliers); pw: packed FP4 weight (mx8_weight_ok).
    Routing: M <= 32 -> the packed codes by TMA + sparse outlier epilogue (msq_gemm_mxfp4w_tiled_blockscaled);
    M > 32 -> the fully TMA-fed kernel (msq_gemm_mxfp8_tma): a plain FP4 weight feeds its packed codes, a weight with
    outliers its widened E4M3 copy (1 byte per element next to the 0.56-byte planes), which is built and kept on `pw`
    from M >= MX8_DENSE_MIN_M (compute-bound) or when cache_dense=True; otherwise the producer-widen kernel
    (msq_gemm_mxfp8_blockscaled) reads the packed planes."""
    from . import _ops
    K = pw.shape[1]
    if xq.shape[-1] != K or xq.block_size != 32 or xq.code_width != 8 or _ops.fmt4(xq.in_fmt)[:2] != (4, 5):
        raise NotImplementedError("activations must be MX-FP8 (E4M3) packed along K with block 32")
    M = 1
    for s_ in xq.shape[:-1]:
        M *= s_
    Nn = pw.shape[0]
    dense = getattr(pw, "_mx8_dense", None)
    terms = int(getattr(xq, "terms", 1)) if accumulate_terms else 1          # stacked slabs accumulated inside the GEMM
    if terms > 1:
        # -> [M, N]; when the TMA-fed kernel cannot take this weight the slabs come back stacked ([terms * M, N])
        M //= terms
        if not (MX8_TMA and M * terms > FP4_TMA_MAX_M and K % 128 == 0 and Nn % 4 == 0 and mx8_tma_weight(
                pw, allow_copy=cache_dense or dense is not None or M >= MX8_DENSE_MIN_M,
                sparse_limit=-1.0 if cache_dense else
                (MX8_TMA_SPARSE_BUDGET if M < MX8_DENSE_MIN_M else MX8_TMA_SPARSE_BUDGET_LARGE_M) / (float(M) * Nn)) is not None):
            return packed_linear_mx8(xq, pw, bias=None, cache_dense=cache_dense)
    if MX8_TMA and M * terms > FP4_TMA_MAX_M and K % 128 == 0 and Nn % 4 == 0:
        # everything above the decode sizes: both operands and both scale streams by TMA, double-buffered accumulators.
        # A plain FP4 weight feeds its packed codes; a weight with outliers its widened E4M3 copy -- built (and kept) only
        # in the compute-bound regime or on request, below that the producer-widen kernel reads the packed planes
        op = mx8_tma_weight(pw, allow_copy=cache_dense or dense is not None or M >= MX8_DENSE_MIN_M,
                            sparse_limit=-1.0 if cache_dense else
                            (MX8_TMA_SPARSE_BUDGET if M < MX8_DENSE_MIN_M else MX8_TMA_SPARSE_BUDGET_LARGE_M) / (float(M) * Nn))
        if op is not None and out_dtype == torch.bfloat16 and bias is None and terms == 1 and op[4] is None:
            # no outliers to correct, one term: the accumulator leaves the epilogue as bf16 (no conversion pass behind the GEMM)
            w, ldw, is_fp4, wsf_t = op[:4]
            out = torch.empty(M, Nn, dtype=torch.bfloat16, device=pw.codes.device)
            nbytes = int(N.lib().msq_gemm_mxfp8_tma_bf16_workspace_bytes(M, Nn, K))
            ws = _workspace(nbytes, out.device)
            rc = N.lib().msq_gemm_mxfp8_tma_bf16(N.ptr(xq.codes), K, N.ptr(xq.e_in), N.ptr(w), ldw, is_fp4, N.ptr(wsf_t), N.ptr(out),
                                                 out.stride(0), M, Nn, K, N.ptr(ws), nbytes, N.stream_ptr(out.device))
            N.check(rc, "msq_gemm_mxfp8_tma_bf16")
            return out.reshape(*xq.shape[:-1], Nn)
        if op is not None:
            out = torch.empty(M, Nn, dtype=torch.float32, device=pw.codes.device)
            w, ldw, is_fp4, wsf_t, o_ptr, o_k, o_dw = op
            behind = o_ptr is not None and (terms > 1 or not MX8_TMA_FUSED_SPARSE)
            nbytes = int(N.lib().msq_gemm_mxfp8_tma_workspace_bytes(M, Nn, K))
            ws = _workspace(nbytes, out.device)
            rc = N.lib().msq_gemm_mxfp8_tma(N.ptr(xq.codes), K, N.ptr(xq.e_in), N.ptr(w), ldw, is_fp4, N.ptr(wsf_t),
                                            N.ptr(None if behind else o_ptr), N.ptr(None if behind else o_k),
                                            N.ptr(None if behind else o_dw), N.ptr(out), out.stride(0), M, Nn, K, terms, 0,
                                            N.ptr(ws), nbytes, N.stream_ptr(out.device))
            N.check(rc, "msq_gemm_mxfp8_tma")
            if behind:
                rc = N.lib().msq_sparse_outlier_correct(N.ptr(xq.codes), K, N.ptr(xq.e_in), N.ptr(pw._mx4_sparse_rows), N.ptr(o_k),
                                                        N.ptr(o_dw), int(o_k.numel()), N.ptr(out), out.stride(0), M * terms, K,
                                                        M if terms > 1 else 0, N.stream_ptr(out.device))
                N.check(rc, "msq_sparse_outlier_correct")
            if bias is not None:
                out += bias.float()
            return out if terms > 1 else out.reshape(*xq.shape[:-1], Nn)
    if FP4_TMA and not cache_dense and M <= FP4_TMA_MAX_M and mx4_weight_ok(pw):
        # decode sizes: the packed 4-bit codes are the tensor-core operand (TMA-expanded), outliers = sparse epilogue
        o_ptr, o_k, o_dw = mx4_sparse_outliers(pw)
        out = torch.empty(M, Nn, dtype=torch.float32, device=pw.codes.device)
        if FP4_TILED:
            ct, st = mx4_tiled_weight(pw)
            rc = N.lib().msq_gemm_mxfp4w_tiled_blockscaled(N.ptr(xq.codes), K, N.ptr(xq.e_in), N.ptr(ct), N.ptr(st),
                                                           N.ptr(o_ptr), N.ptr(o_k), N.ptr(o_dw), N.ptr(out), out.stride(0), M,
                                                           Nn, K, 0, N.stream_ptr(out.device))
        else:
            rc = N.lib().msq_gemm_mxfp4w_blockscaled(N.ptr(xq.codes), K, N.ptr(xq.e_in), N.ptr(pw.codes), N.ptr(pw.e_in),
                                                     N.ptr(o_ptr), N.ptr(o_k), N.ptr(o_dw), N.ptr(out), out.stride(0), M, Nn, K,
                                                     0, N.stream_ptr(out.device))
        N.check(rc, "msq_gemm_mxfp4w_blockscaled")
        if bias is not None:
            out += bias.float()
        return out.reshape(*xq.shape[:-1], Nn)
    if dense is None and not mx8_weight_ok(pw):
        raise NotImplementedError("weight is not eligible for the block-scaled FP8 path")
    out = torch.empty(M, Nn, dtype=torch.float32, device=pw.codes.device)
    if K % 128 == 0 and (dense is not None or cache_dense or M >= MX8_DENSE_MIN_M):
        # widen the FP4 weight to E4M3 bytes (once per call in the compute-bound regime, once per weight when
        # cached), then both operands stream through TMA
        if dense is None:
            dense = widen_weight_mx8(pw)
            if cache_dense:
                pw._mx8_dense = dense
        wb, wsf = dense
        rc = N.lib().msq_gemm_mxfp8_dense(N.ptr(xq.codes), K, N.ptr(xq.e_in), N.ptr(wb), K, N.ptr(wsf), N.ptr(out),
                                          out.stride(0), M, Nn, K, 0, N.stream_ptr(out.device))
        N.check(rc, "msq_gemm_mxfp8_dense")
        if bias is not None:
            out += bias.float()
        return out.reshape(*xq.shape[:-1], Nn)
    rc = N.lib().msq_gemm_mxfp8_blockscaled(N.ptr(xq.codes), K, N.ptr(xq.e_in), N.ptr(pw.codes), N.ptr(pw.bitmap),
                                            N.ptr(pw.e_in), N.ptr(pw.e_out), N.ptr(out), out.stride(0), M, Nn, K,
                                            int(pw.plain), 0, N.stream_ptr(out.device))
    N.check(rc, "msq_gemm_mxfp8_blockscaled")
    if bias is not None:
        out += bias.float()
    return out.reshape(*xq.shape[:-1], Nn)
