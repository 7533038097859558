"""GPTQ solver on B200 -- mirror of the reference's llm/gptq.py (class GPTQ: add_batch,
fasterquant, free; attributes layer, dev, rows, columns, H, nsamples, quantizer).

add_batch   -> tcgen05 Hessian accumulation (csrc/msq_gemm.cu), optionally batched over calls
fasterquant -> Cholesky trio through torch.linalg (cuSOLVER), then per 128-column block the fused
               column-loop kernel + fp32 trailing update (csrc/msq_gptq.cu).  When the fused path's
               "prune is a no-op" check fails, or for quantizer settings it does not cover, the
               column-synchronous generic path below reproduces gptq.py:116-160 op by op on the GPU.
"""
import ctypes as C
import math
import time

import torch
import torch.nn as nn

try:
    import transformers
    _Conv1D = transformers.Conv1D
except Exception:  # pragma: no cover
    _Conv1D = ()

from .. import _native as N
from .. import _ops, _tc
from ..utils.quant import *  # noqa: F401,F403  (the reference re-exports utils.quant from gptq)
from ..utils.quant import MXQuantizer, Quantizer, quantize, quantize_mx_outlier_hessian

DEBUG = False

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

_p, _ll, _i, _f = C.c_void_p, C.c_longlong, C.c_int, C.c_float
N.register("msq_gptq_block", _i, [_p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _i,
                                  _i, _i, _i, _f, _i, _i, _i, _f, _f, _i, _i, _p])
N.register("msq_gptq_trailing_update", _i, [_p, _ll, _p, _p, _ll, _i, _i, _i, _i, _i, _p])
N.register("msq_gptq_check_prune", _i, [_p, _p, _i, _p, _p])
N.register("msq_rank1_sub_f32", _i, [_p, _ll, _p, _p, _i, _i, _p])
N.register("msq_gptq_solve_workspace_bytes", _ll, [_i])
N.register("msq_gptq_solve_f32", _i, [_p, _ll, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _p, _ll, _p, _p, _p, _p, _p,
                                      _i, _i, _i, _i, _i, _i, _f, _i, _i, _i, _f, _f, _i, _i, _p, _p, _p, _p, _p])
N.register("msq_pack_from_codes", _i, [_p, _p, _ll, _ll, _ll, _i, _i, _i, _i, _i, _p, _p, _p, _p])


class _Shared(object):
    """Hessian state that several GPTQ objects may share: linears fed by the SAME activations (q/k/v, gate/up) have
    identical Hessians, so one accumulation and one factorization serve all of them (identical inputs => identical
    bits; llm/llama.py:139-159 builds and factors each of them separately)."""

    def __init__(self, columns, dev):
        self.H = torch.zeros((columns, columns), device=dev)
        self.nsamples = 0
        self.pending, self.pending_n = [], 0
        self.users = 1
        self.factor = {}            # (percdamp, actorder, method) -> dict(Hinv, ut, dead, perm, invperm)


class GPTQ:

    #: number of add_batch calls whose activations are buffered before one Hessian GEMM
    #: (1 = accumulate immediately, exactly like gptq.py:53-58)
    hessian_batch = 1
    #: the reference calls torch.cuda.empty_cache() in free() (gptq.py:191).  Kept on by default: measured on a
    #: LLaMA-2-7B-shaped run, skipping it made the solves slower (36.5 s vs 19.7 s over 224 linears), not faster.
    release_cache_on_free = True
    #: where the tie order of the prune's torch.topk (gptq.py:150) comes from: "cuda" (default) or "cpu"
    prune_topk_device = "cuda"
    #: print 'time' / 'error' like the reference (gptq.py:174-175)
    verbose = True
    #: how the upper Cholesky factor of H^-1 (gptq.py:101-103) is obtained.
    #:   "native"    (default) msq_cholinv_upper_f32: one reverse Cholesky H = R R^T + one triangular inverse, blocked,
    #:               every panel product on tcgen05 (fp32-accurate split-plane GEMM); no library call.  Same matrix as
    #:               the reference's three calls up to fp32 rounding (tests/test_gpu_cholinv.py states the bound).
    #:   "reference" torch.linalg.cholesky -> torch.cholesky_inverse -> torch.linalg.cholesky(upper) (cuSOLVER)
    #:   "flip"      torch: one Cholesky of the index-reversed matrix + one triangular solve (the algebra of "native")
    hinv_method = "native"
    #: trailing update W[:, i2:] -= Err1 Hinv[i1:i2, i2:] (gptq.py:165) on tcgen05 (True) or in fp32 SIMT (False)
    trailing_tc = True
    #: keep the packed artefact of the result (self.packed: PackedMX whose dequantize() is the new weight bit for bit,
    #: built from the column kernel's own codes) -- what save_packed / MXLinear.from_packed / the row-block GEMM consume.
    #: Available on the fused path without act-order (None otherwise).
    emit_packed = False

    def __init__(self, layer, prune=True):
        #: prune=False is the reference's vlm/gptq.py:130-144 variant: the column quantizer is quantize_mx_outlier_v1
        #: and no inlier is zeroed afterwards (llm/gptq.py:146-153 is skipped), so rows never interact
        self.prune = prune
        self.layer = layer
        self.dev = self.layer.weight.device
        W = layer.weight.data.clone()
        if isinstance(self.layer, nn.Conv2d):
            W = W.flatten(1)
        if isinstance(self.layer, _Conv1D):
            W = W.t()
        self.rows = W.shape[0]
        self.columns = W.shape[1]
        self._share = _Shared(self.columns, self.dev)
        self.packed = None
        self._emitted = None
        self.last_path = None
        self.last_error = None

    # -- Hessian ---------------------------------------------------------------------------
    @property
    def H(self):
        self._flush()
        return self._share.H

    @H.setter
    def H(self, value):
        self._share.pending, self._share.pending_n = [], 0
        self._share.H = value

    @H.deleter
    def H(self):
        self._share.pending, self._share.pending_n = [], 0
        self._share.H = None

    @property
    def nsamples(self):
        return self._share.nsamples

    @nsamples.setter
    def nsamples(self, value):
        self._share.nsamples = value

    def share_hessian_with(self, owner):
        """use `owner`'s Hessian (and its factorization) instead of an own one: for linears that receive the same
        input tensor.  After this call only the owner's add_batch accumulates; this object's add_batch is a no-op."""
        if owner is self or self._share is owner._share:
            return
        assert owner.columns == self.columns, "shared Hessians need the same number of input features"
        self._share.users -= 1
        self._share = owner._share
        self._share.users += 1
        self._hessian_owner = owner

    def add_batch(self, inp, out):
        if getattr(self, "_hessian_owner", None) is not None:
            return                                    # the owner accumulates for the whole group
        if DEBUG:
            self.inp1 = inp
            self.out1 = out
        if len(inp.shape) == 2:
            inp = inp.unsqueeze(0)
        b = inp.shape[0]
        if isinstance(self.layer, nn.Linear) or isinstance(self.layer, _Conv1D):
            if len(inp.shape) == 3:
                inp = inp.reshape((-1, inp.shape[-1]))
            x = inp                                   # [tokens, features]; the reference transposes (gptq.py:42)
        elif isinstance(self.layer, nn.Conv2d):
            unfold = nn.Unfold(self.layer.kernel_size, dilation=self.layer.dilation, padding=self.layer.padding,
                               stride=self.layer.stride)
            x = unfold(inp).permute([1, 0, 2]).flatten(1).t()
        else:
            raise TypeError("unsupported layer type %s" % type(self.layer))
        N.require_cuda(x, "inp")
        if x.dtype not in (torch.float16, torch.bfloat16, torch.float32):
            x = x.float()
        sh = self._share
        sh.pending.append((x.detach().contiguous(), b))
        sh.pending_n += 1
        if sh.pending_n >= self.hessian_batch:
            self._flush()

    def _flush(self):
        """H <- H * n/(n+b) + (2/(n+b)) X^T X   (gptq.py:53-58) for the buffered calls"""
        sh = self._share
        if not sh.pending:
            return
        xs = [x for x, _ in sh.pending]
        b = sum(bb for _, bb in sh.pending)
        X = xs[0] if len(xs) == 1 else torch.cat(xs, 0)
        n_new = sh.nsamples + b
        _tc.hessian_accumulate(sh.H, X, alpha=2.0 / n_new, beta=sh.nsamples / n_new)
        sh.nsamples = n_new
        sh.pending, sh.pending_n = [], 0

    # -- solver ----------------------------------------------------------------------------
    def _fused_ok(self):
        q = self.quantizer
        if not isinstance(q, MXQuantizer) or q.inlier_elem_format is None:
            return False
        axes = q.axes if isinstance(q.axes, (list, tuple)) else [q.axes]
        if list(axes) not in ([0], [-2]) or q.block_size not in (8, 16, 32):
            return False
        if q.shared_exp_method != "max" or q.round not in ("nearest", "floor", "even"):
            return False
        return _ops.fmt4(q.inlier_elem_format)[0] <= 5 and _ops.fmt4(q.outlier_elem_format)[0] <= 5

    def fasterquant(self, blocksize=128, percdamp=.01, groupsize=-1, actorder=False, static_groups=False):
        st = self._fq_launch(blocksize, percdamp, groupsize, actorder, static_groups)
        self._fq_finish(st)
        torch.cuda.synchronize()                           # gptq.py:173

    def _fq_launch(self, blocksize=128, percdamp=.01, groupsize=-1, actorder=False, static_groups=False):
        """first half of fasterquant: everything up to the queued fused solve (no host synchronisation), on the
        current stream.  _fq_finish completes it.  The layer-sequential driver launches the independent linears of a
        layer on separate streams this way (the column kernel is latency bound and fills a fraction of the SMs)."""
        W = self.layer.weight.data.clone()
        if isinstance(self.layer, nn.Conv2d):
            W = W.flatten(1)
        if isinstance(self.layer, _Conv1D):
            W = W.t()
        W = W.float().contiguous()
        N.require_cuda(W, "layer.weight")

        tick = time.time()

        if not self.quantizer.ready():
            self.quantizer.find_params(W, weight=True)

        fac = self._factorize(percdamp, actorder)
        dead, perm = fac["dead"], fac["perm"]
        W.masked_fill_(dead.unsqueeze(0), 0)               # W[:, dead] = 0

        groups = None
        if static_groups:
            import copy
            groups = []
            for i in range(0, self.columns, groupsize):
                quantizer = copy.deepcopy(self.quantizer)
                quantizer.find_params(W[:, i:(i + groupsize)], weight=True)
                groups.append(quantizer)

        if actorder:
            W = W[:, perm].contiguous()

        raw = None
        if self._fused_ok() and blocksize == 128:
            raw = self._solve_fused_raw(W.clone(), fac["Hinv"], fac["ut"], emit=self.emit_packed)
        return dict(tick=tick, W=W, fac=fac, raw=raw, groups=groups,
                    args=(blocksize, groupsize, actorder, static_groups))

    def _fq_finish(self, st):
        """second half of fasterquant: read the flags of the fused solve (one host sync), fall back to the
        column-synchronous path if the prune was not a no-op, write the weight back"""
        W, fac = st["W"], st["fac"]
        blocksize, groupsize, actorder, static_groups = st["args"]
        perm, invperm, Hinv = fac["perm"], fac["invperm"], fac["Hinv"]
        Q = None
        if st["raw"] is not None:
            Q, loss = self._solve_fused_check(st["raw"], W.shape, fac["info"])
            self.last_path = "fused"
        if Q is None:
            if fac["info"] is not None and int(fac["info"].item()):
                raise torch.linalg.LinAlgError("GPTQ: the damped Hessian is not positive-definite (leading minor %d)"
                                               % int(fac["info"].item()))
            Q, loss = self._solve_generic(W, Hinv, blocksize, groupsize, actorder, static_groups, st["groups"], perm)
            self.last_path = "generic"

        self.last_error = loss
        if self.verbose:
            print('time %.2f' % (time.time() - st["tick"]))
            print('error', loss)

        if actorder:
            Q = Q[:, invperm]
            self.packed = None                             # the artefact is in permuted column order
        if isinstance(self.layer, _Conv1D):
            Q = Q.t()
        self.layer.weight.data = Q.reshape(self.layer.weight.shape).to(self.layer.weight.data.dtype)
        if DEBUG:
            print(torch.sum((self.layer(self.inp1) - self.out1) ** 2))

    def _factorize(self, percdamp, actorder):
        """gptq.py:78-103: dead columns, act-order permutation, damping and the Cholesky trio.  The result is cached on
        the shared Hessian state, so the members of a share group factor once."""
        sh = self._share
        key = (float(percdamp), bool(actorder), self.hinv_method)
        fac = sh.factor.get(key)
        if fac is not None:
            return fac
        H = self.H
        if H is None:
            raise RuntimeError("GPTQ.fasterquant: the Hessian has already been consumed with different settings")
        del self.H
        dead = torch.diag(H) == 0
        H.diagonal().masked_fill_(dead, 1)                 # H[dead, dead] = 1 without the host sync of mask indexing
        perm = invperm = None
        if actorder:
            perm = torch.argsort(torch.diag(H), descending=True)
            H = H[perm][:, perm]
            invperm = torch.argsort(perm)
        damp = percdamp * torch.mean(torch.diag(H))
        diag = torch.arange(self.columns, device=self.dev)
        H[diag, diag] += damp
        ut = None
        if self.hinv_method == "native":
            ci = _tc.cholinv_upper(H, check=False)
            Hinv, ut = ci.U, ci.ut
            fac = dict(info=ci.info)
        else:
            if self.hinv_method == "flip":
                R = torch.linalg.cholesky(H.flip(0, 1)).flip(0, 1)          # upper, H = R R^T
                H = torch.linalg.solve_triangular(R, torch.eye(self.columns, device=self.dev, dtype=R.dtype), upper=True)
            else:
                H = torch.linalg.cholesky(H)
                H = torch.cholesky_inverse(H)
                H = torch.linalg.cholesky(H, upper=True)
            Hinv = H.contiguous()
            fac = dict(info=None)
        fac.update(dead=dead, perm=perm, invperm=invperm, Hinv=Hinv, ut=ut)
        sh.factor[key] = fac
        return fac

    def _solve_fused_raw(self, W, Hinv, ut=None, emit=False):
        """queue gptq.py:106-165 for the rows in W (msq_gptq_solve_f32); returns device tensors, no host sync:
        (Q, loss[1] fp64, counters int32 [2*cols + 3] = zeros[cols] | kcnt[cols] | status | violation | info).
        emit=True also keeps the quantizer's own artefact (code bytes, bitmap, block exponents) in self._emitted."""
        q = self.quantizer
        dev = W.device
        rows, cols = W.shape
        fi, fo = _ops.fmt4(q.inlier_elem_format), _ops.fmt4(q.outlier_elem_format)
        Q = torch.empty_like(W)
        em = None
        if emit and cols % 32 == 0 and rows % int(q.block_size) == 0:
            nbk = (rows // int(q.block_size)) * cols
            em = dict(codes8=torch.empty(rows, cols, dtype=torch.uint8, device=dev),
                      bitmap=torch.empty(rows * cols // 32, dtype=torch.int32, device=dev),
                      e_in=torch.empty(nbk, dtype=torch.int8, device=dev), e_out=torch.empty(nbk, dtype=torch.int8, device=dev))
        self._emitted = em
        L, s = N.lib(), N.stream_ptr(dev)
        nws = int(L.msq_gptq_solve_workspace_bytes(rows))
        ws = _tc._workspace(nws, dev)
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        counters = torch.zeros(2 * cols + 3, dtype=torch.int32, device=dev)
        zeros, kcnt, status, viol = counters[:cols], counters[cols:2 * cols], counters[2 * cols:2 * cols + 1], counters[2 * cols + 1:2 * cols + 2]
        if ut is None and self.trailing_tc and cols % 8 == 0:
            ut = _tc.split3_planes(Hinv.t().contiguous())      # Hinv from a library route: build the operand planes
        if not self.trailing_tc or cols % 8:
            ut = None
        rc = L.msq_gptq_solve_f32(N.ptr(W), W.stride(0), N.ptr(Q), Q.stride(0), N.ptr(Hinv), Hinv.stride(0),
                                  ut.ptr if ut is not None else None, ut.ld if ut is not None else 0,
                                  ut.plane_rows if ut is not None else 0, rows, cols, N.ptr(ws), nws, N.ptr(loss),
                                  N.ptr(zeros), N.ptr(kcnt), N.ptr(status), N.ptr(viol), int(q.block_size),
                                  int(q.inlier_scale_bits), int(q.outlier_scale_bits), fi[0], fi[1], fi[2], fi[3],
                                  fo[0], fo[1], fo[2], fo[3], float(q.std_dev), int(bool(q.flush_fp32_subnorms)),
                                  _ops._rmode(q.round), N.ptr(em["codes8"]) if em else None,
                                  N.ptr(em["bitmap"]) if em else None, N.ptr(em["e_in"]) if em else None,
                                  N.ptr(em["e_out"]) if em else None, s)
        N.check(rc, "msq_gptq_solve_f32")
        return Q, loss, counters

    def _packed_from_emitted(self, rows, cols):
        """PackedMX of the GPTQ result from the column kernel's own codes (msq_pack_from_codes): dequantize() == Q"""
        from ..packed import PackedMX, _alloc_planes
        em, q = self._emitted, self.quantizer
        if em is None:
            return None
        fi, fo = _ops.fmt4(q.inlier_elem_format), _ops.fmt4(q.outlier_elem_format)
        bs = int(q.block_size)
        ref = em["codes8"]
        pl, wi, hi_w, stride, nseg = _alloc_planes(ref, 1, rows, cols, bs, fi, fo, False)
        rc = N.lib().msq_pack_from_codes(N.ptr(em["codes8"]), N.ptr(em["bitmap"]), 1, rows, cols, bs, fi[0], fi[1], fo[0],
                                         fo[1], N.ptr(pl["codes"]), N.ptr(pl["hi"]), N.ptr(pl["seg_cnt"]),
                                         N.stream_ptr(ref.device))
        N.check(rc, "msq_pack_from_codes")
        pl["bitmap"], pl["e_in"], pl["e_out"] = em["bitmap"], em["e_in"], em["e_out"]
        self._emitted = None
        return PackedMX(shape=(rows, cols), axis=0, block_size=bs, in_fmt=q.inlier_elem_format,
                        out_fmt=q.outlier_elem_format, in_scale_bits=q.inlier_scale_bits,
                        out_scale_bits=q.outlier_scale_bits, code_width=wi, hi_w=hi_w, hi_seg_stride=stride, nseg=nseg,
                        plain=False, path=4, status=None, **pl)

    def _solve_fused_check(self, raw, shape, info=None):
        """host side of the fused solve: one read of (status, violation, info); (Q, loss) or (None, None)"""
        Q, loss, counters = raw
        rows, cols = shape
        if info is not None:                          # pivot check of the native factorization rides on the same read
            counters[2 * cols + 2:] = info
        flags = counters[2 * cols:].tolist()          # one host sync per linear
        if flags[2]:
            raise torch.linalg.LinAlgError("GPTQ: the damped Hessian is not positive-definite (leading minor %d)" % flags[2])
        _ops.raise_on_status(flags[0])
        if flags[1] != 0 and self.prune:
            self._emitted = None
            return None, None
        self.packed = self._packed_from_emitted(rows, cols)
        return Q, float(loss.item())

    def _solve_fused(self, W, Hinv, ut=None, info=None):
        """gptq.py:106-165 as ONE queued call (msq_gptq_solve_f32): per 128-column block the two-warp column kernel and
        the trailing update; returns (Q, loss) or (None, None) when the prune would have changed values (then the
        generic path must run)."""
        raw = self._solve_fused_raw(W, Hinv, ut, emit=self.emit_packed)
        return self._solve_fused_check(raw, W.shape, info)

    def _solve_generic(self, W, Hinv, blocksize, groupsize, actorder, static_groups, groups, perm):
        """column-synchronous path: gptq.py:106-165 statement by statement, on the GPU"""
        Q = torch.zeros_like(W)
        loss = torch.zeros((), dtype=torch.float64, device=W.device)
        L, s = N.lib(), N.stream_ptr(W.device)
        for i1 in range(0, self.columns, blocksize):
            i2 = min(i1 + blocksize, self.columns)
            count = i2 - i1
            W1 = W[:, i1:i2].clone()
            Q1 = torch.zeros_like(W1)
            Err1 = torch.zeros_like(W1)
            Losses1 = torch.zeros_like(W1)
            Hinv1 = Hinv[i1:i2, i1:i2].contiguous()
            for i in range(count):
                w = W1[:, i]
                d = Hinv1[i, i]
                if groupsize != -1:
                    if not static_groups:
                        if (i1 + i) % groupsize == 0:
                            self.quantizer.find_params(W[:, (i1 + i):(i1 + i + groupsize)], weight=True)
                    else:
                        idx = i1 + i
                        if actorder:
                            idx = perm[idx]
                        self.quantizer = groups[idx // groupsize]
                qz = self.quantizer
                if isinstance(qz, MXQuantizer):
                    res = quantize_mx_outlier_hessian(w.unsqueeze(1).contiguous(), qz.inlier_scale_bits,
                                                      qz.outlier_scale_bits, qz.inlier_elem_format,
                                                      qz.outlier_elem_format, qz.shared_exp_method, qz.std_dev, qz.axes,
                                                      qz.block_size, qz.round, qz.flush_fp32_subnorms, qz.custom_cuda)
                    if isinstance(res, tuple):
                        q, num_outliers_per_block = res
                        q = q.flatten()
                        importance = (q ** 2) / d ** 2
                        num_outliers = int(num_outliers_per_block.sum().to(torch.int16)) if self.prune else 0
                        if num_outliers > 0:
                            # gptq.py:150: torch.topk leaves the order of tied entries unspecified, and on the coarse
                            # MX grid most of a column's q^2/d^2 values are tied.  The default is the CUDA tie order
                            # (what the reference itself gets on a GPU); prune_topk_device = "cpu" reproduces a
                            # CPU run of the reference (the committed fixtures) entry for entry.
                            if self.prune_topk_device == "cpu":
                                least = torch.topk(importance.cpu(), num_outliers, largest=False).indices.to(q.device)
                            else:
                                least = torch.topk(importance, num_outliers, largest=False).indices
                            q[least] = 0
                    else:
                        q = res.flatten()
                else:   # plain GPTQ quantizer (upstream GPTQ behaviour; the reference's fork cannot reach it)
                    q = quantize(w.unsqueeze(1), qz.scale, qz.zero, qz.maxq).flatten()
                Q1[:, i] = q
                Losses1[:, i] = (w - q) ** 2 / d ** 2
                err1 = ((w - q) / d).contiguous()
                rc = L.msq_rank1_sub_f32(N.ptr(W1[:, i:]), W1.stride(0), N.ptr(err1), N.ptr(Hinv1[i, i:]), self.rows,
                                         count - i, s)
                N.check(rc, "msq_rank1_sub_f32")
                Err1[:, i] = err1
            Q[:, i1:i2] = Q1
            loss += (Losses1 / 2).double().sum()
            if i2 < self.columns:
                E = torch.zeros(self.rows, 128, dtype=torch.float32, device=W.device) if count <= 128 else None
                if E is not None:
                    E[:, :count] = Err1
                    rc = L.msq_gptq_trailing_update(N.ptr(W), W.stride(0), N.ptr(E), N.ptr(Hinv), Hinv.stride(0),
                                                    self.rows, i1, count, i2, self.columns, s)
                    N.check(rc, "msq_gptq_trailing_update")
                else:   # blocksize > 128: chunk the K dimension of the update
                    for c0 in range(0, count, 128):
                        c1 = min(c0 + 128, count)
                        E = torch.zeros(self.rows, 128, dtype=torch.float32, device=W.device)
                        E[:, :c1 - c0] = Err1[:, c0:c1]
                        rc = L.msq_gptq_trailing_update(N.ptr(W), W.stride(0), N.ptr(E), N.ptr(Hinv), Hinv.stride(0),
                                                        self.rows, i1 + c0, c1 - c0, i2, self.columns, s)
                        N.check(rc, "msq_gptq_trailing_update")
        return Q, float(loss.item())

    def free(self):
        if DEBUG:
            self.inp1 = None
            self.out1 = None
        sh = self._share
        sh.users -= 1
        if sh.users <= 0:
            sh.H = None
            sh.pending, sh.pending_n = [], 0
            sh.factor = {}
        self._share = _Shared.__new__(_Shared)
        self._share.H, self._share.nsamples, self._share.pending, self._share.pending_n = None, 0, [], 0
        self._share.users, self._share.factor = 1, {}
        self.Losses = None
        self.Trace = None
        if self.release_cache_on_free:
            torch.cuda.empty_cache()
