"""Multi-GPU plumbing for the quantization path (one process per GPU, torch.distributed).

The path shards at three levels; only the first needs a collective:
  * calibration samples  -> per-rank Hessians, ONE sum-allreduce of the K x K fp32 matrix per linear
  * linears / row blocks -> independent GPTQ solves, results broadcast / all-gathered
  * packed GEMM          -> output features split across ranks, no collective

Everything here is backend-agnostic host logic (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(n, rank, world):
    """contiguous [lo, hi) share of n items for `rank` (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def combine_hessians(H_local, n_local, group=None, upper_only=True):
    """Each rank holds the running-mean Hessian of ITS samples, H_g = (2/n_g) sum x x^T
    (llm/gptq.py:53-58).  Returns (H, N) with H = (2/N) sum over all samples:
        H = sum_g (n_g / N) H_g          -- one sum-allreduce.
    upper_only packs the upper triangle (H is symmetric): ~half the bytes on the wire."""
    if not is_dist():
        return H_local, n_local
    dev = H_local.device
    n = torch.tensor([float(n_local)], dtype=torch.float64, device=dev)
    dist.all_reduce(n, group=group)
    N = int(round(float(n.item())))
    if N == 0:
        return H_local, 0
    w = float(n_local) / N
    K = H_local.shape[0]
    if upper_only:
        iu = torch.triu_indices(K, K, device=dev)
        buf = (H_local[iu[0], iu[1]] * w).contiguous()
        dist.all_reduce(buf, group=group)
        H = torch.empty_like(H_local)
        H[iu[0], iu[1]] = buf
        H[iu[1], iu[0]] = buf
    else:
        H = (H_local * w).contiguous()
        dist.all_reduce(H, group=group)
    return H, N


def allreduce_gptq_hessian(gptq, group=None):
    """in-place version for a GPTQ object: after this every rank holds the global H / nsamples"""
    H, N = combine_hessians(gptq.H, gptq.nsamples, group=group)
    gptq.H = H
    gptq.nsamples = N
    return gptq


def assign_linears(shapes, world):
    """Greedy longest-processing-time assignment of independent linears to ranks.
    shapes: {name: (rows, cols)}; cost model rows*cols^2 + cols^3 (update + Cholesky trio).
    Returns {name: rank}."""
    cost = {k: r * c * c + c ** 3 for k, (r, c) in shapes.items()}
    load = [0] * world
    out = {}
    for name in sorted(cost, key=lambda k: (-cost[k], k)):
        r = min(range(world), key=lambda i: (load[i], i))
        out[name] = r
        load[r] += cost[name]
    return out


def row_shards(rows, block_size, world):
    """row ranges for sharding ONE linear: boundaries are multiples of block_size^2 so that MX blocks
    and the `outlier_pos[::block_size]` count (utils/quant.py:66) stay rank-local.  The top-k prune
    (llm/gptq.py:146-153) couples all rows of a column; row sharding is therefore exact only when the
    prune is a no-op (checked by the fused solver) -- otherwise shard by linear instead."""
    unit = block_size * block_size
    nunits = (rows + unit - 1) // unit
    out = []
    for r in range(world):
        lo, hi = shard_range(nunits, r, world)
        out.append((min(lo * unit, rows), min(hi * unit, rows)))
    return out


def broadcast_weight(layer, src, group=None):
    """after rank `src` quantized `layer`, give every rank the result"""
    if is_dist():
        dist.broadcast(layer.weight.data, src=src, group=group)


def all_gather_rows(Q_local, row_ranges, rows, group=None):
    """assemble a row-sharded quantized weight on every rank"""
    if not is_dist():
        return Q_local
    world = dist.get_world_size(group)
    cols = Q_local.shape[1]
    maxr = max(hi - lo for lo, hi in row_ranges)
    pad = torch.zeros(maxr, cols, dtype=Q_local.dtype, device=Q_local.device)
    pad[: Q_local.shape[0]] = Q_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.empty(rows, cols, dtype=Q_local.dtype, device=Q_local.device)
    for (lo, hi), p in zip(row_ranges, parts):
        out[lo:hi] = p[: hi - lo]
    return out


def fasterquant_row_sharded(gq, percdamp=.01, group=None, solve_fn=None):
    """GPTQ.fasterquant (llm/gptq.py:60-193, blocksize 128, no act-order / groups) with the ROWS of one linear sharded over
    the ranks.  Every rank holds the same weight and the same (allreduced) Hessian, factors it
    (identical inputs => identical Hinv on every rank; the factorization is small next to the solve of a 70B-sized
    linear), runs the fused column loop on ITS row range -- boundaries are multiples of block_size^2, so MX blocks and
    the `outlier_pos[::block_size]` count stay rank-local -- and the quantized rows are all-gathered.
    The only coupling between rows is the top-k prune (llm/gptq.py:146-153): it is a no-op when every column holds at
    least k exact zeros, which needs the per-column counts of ALL rows: one int32 sum-allreduce of [2 * cols]
    counters.  If the check fails (or the quantizer has no fused path) every rank runs the unsharded solver instead.
    solve_fn(W_rows, fac) -> (Q_rows, loss[1], counters[2*cols+3]) replaces the CUDA solve in the CPU (gloo) tests.
    Returns the path taken: "row-sharded" or "replicated"."""
    import torch.nn as nn
    world = dist.get_world_size(group) if is_dist() else 1
    rank = dist.get_rank(group) if is_dist() else 0
    q = gq.quantizer
    bs = int(getattr(q, "block_size", 16) or 16)
    fused = solve_fn is not None or gq._fused_ok()
    if world == 1 or not fused:
        gq.fasterquant(percdamp=percdamp)
        return "replicated"
    W = gq.layer.weight.data.clone()
    if isinstance(gq.layer, nn.Conv2d):
        W = W.flatten(1)
    W = W.float().contiguous()
    fac = gq._factorize(percdamp, False)
    W.masked_fill_(fac["dead"].unsqueeze(0), 0)
    ranges = row_shards(gq.rows, bs, world)
    lo, hi = ranges[rank]
    cols = gq.columns
    if hi > lo:
        Wl = W[lo:hi].contiguous()
        Ql, loss, counters = solve_fn(Wl, fac) if solve_fn is not None else gq._solve_fused_raw(Wl, fac["Hinv"], fac["ut"])
    else:
        Ql = W[0:0]
        loss = torch.zeros(1, dtype=torch.float64, device=W.device)
        counters = torch.zeros(2 * cols + 3, dtype=torch.int32, device=W.device)
    if fac.get("info") is not None:
        counters[2 * cols + 2:] = fac["info"]
    dist.all_reduce(counters, group=group)            # zeros / k per column over all rows; status / info: any rank
    dist.all_reduce(loss, group=group)
    zeros, kcnt = counters[:cols], counters[cols:2 * cols]
    flags = counters[2 * cols:].tolist()
    if flags[2]:
        raise torch.linalg.LinAlgError("GPTQ: the damped Hessian is not positive-definite")
    from . import _ops
    _ops.raise_on_status(1 if (flags[0] & 1) else 0)
    violated = bool((zeros < kcnt.to(torch.int16).to(torch.int32)).any()) and getattr(gq, "prune", True)
    if violated:
        gq._share.factor[(float(percdamp), False, gq.hinv_method)] = fac       # reuse the factorization
        gq.fasterquant(percdamp=percdamp)
        return "replicated"
    Q = all_gather_rows(Ql, ranges, gq.rows, group=group)
    gq.last_error = float(loss.item())
    gq.last_path = "row-sharded"
    gq.layer.weight.data = Q.reshape(gq.layer.weight.shape).to(gq.layer.weight.data.dtype)
    return "row-sharded"
