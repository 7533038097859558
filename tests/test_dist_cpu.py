"""CPU tier, world_size 2, gloo: the host-side multi-GPU logic (Hessian combine, sharding maps,
row gather).  The per-rank Hessians are produced by the oracle, so no GPU is needed."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import msq_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from msq_b200 import dist as D
    torch.manual_seed(0)
    K, T, nsamp = 48, 32, 7                       # 7 samples over 2 ranks: uneven split on purpose
    X = torch.randn(nsamp, T, K)
    lo, hi = D.shard_range(nsamp, rank, world)
    H, n = torch.zeros(K, K), 0
    for s in range(lo, hi):
        H, n = O.hessian_add_batch(H, n, X[s:s + 1])
    Hg, N = D.combine_hessians(H, n)
    Hg2, _ = D.combine_hessians(H, n, upper_only=False)
    Href, nref = torch.zeros(K, K), 0
    for s in range(nsamp):
        Href, nref = O.hessian_add_batch(Href, nref, X[s:s + 1])
    ok = N == nsamp and float((Hg - Href).abs().max()) < 1e-5 * float(Href.abs().max())
    ok = ok and float((Hg2 - Hg).abs().max()) < 1e-6 * float(Href.abs().max()) and torch.equal(Hg, Hg.t())
    # row-sharded result gather
    rows, cols, bs = 1000, 8, 4
    rr = D.row_shards(rows, bs, world)
    full = torch.arange(rows * cols, dtype=torch.float32).reshape(rows, cols)
    got = D.all_gather_rows(full[rr[rank][0]:rr[rank][1]].clone(), rr, rows)
    ok = ok and torch.equal(got, full)
    # broadcast of a quantized layer from its owner
    layer = torch.nn.Linear(4, 4, bias=False)
    layer.weight.data.fill_(float(rank + 1))
    D.broadcast_weight(layer, src=1)
    ok = ok and bool((layer.weight.data == 2.0).all())
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def test_sharding_maps():
    from msq_b200 import dist as D
    assert [D.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    shapes = {"q": (4096, 4096), "k": (4096, 4096), "v": (4096, 4096), "o": (4096, 4096), "gate": (11008, 4096),
              "up": (11008, 4096), "down": (4096, 11008)}
    a = D.assign_linears(shapes, 8)
    assert sorted(set(a.values())) == list(range(7))          # 7 linears on 7 of 8 ranks
    a2 = D.assign_linears(shapes, 2)
    cost = lambda k: shapes[k][0] * shapes[k][1] ** 2 + shapes[k][1] ** 3
    loads = [sum(cost(k) for k in a2 if a2[k] == r) for r in range(2)]
    assert max(loads) / sum(loads) < 0.8
    rs = D.row_shards(4096, 16, 8)
    assert rs[0] == (0, 512) and rs[-1][1] == 4096 and all(lo % 256 == 0 for lo, _ in rs)
    rs = D.row_shards(1000, 16, 3)                            # ragged: last shard shorter, none negative
    assert rs[-1][1] == 1000 and all(hi >= lo for lo, hi in rs)


class _OracleGPTQ(object):
    """stand-in for msq_b200.GPTQ on the CPU tier: same attributes, factorization and solves done by the oracle"""

    def __init__(self, layer, H, cfg):
        from types import SimpleNamespace
        self.layer, self.rows, self.columns = layer, layer.weight.shape[0], layer.weight.shape[1]
        self.quantizer = SimpleNamespace(block_size=cfg["block_size"])
        self.H, self.cfg, self.prune, self.hinv_method = H, cfg, True, "oracle"
        self._share = SimpleNamespace(factor={})
        self.last_path = self.last_error = None
        self.replicated_calls = 0

    def _fused_ok(self):
        return True

    def _factorize(self, percdamp, actorder):
        W0 = self.layer.weight.data.clone().float()
        Hinv, _ = O.hessian_inverse_factor(self.H.clone(), W0, percdamp)
        return dict(Hinv=Hinv, ut=None, info=None, dead=torch.diag(self.H) == 0, perm=None, invperm=None)

    def fasterquant(self, percdamp=.01):
        self.replicated_calls += 1


def _rowshard_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from msq_b200 import dist as D
    g = torch.Generator().manual_seed(5)
    rows, cols = 768, 160                                     # 3 units of 16^2 rows over 2 ranks: uneven on purpose
    W = torch.randn(rows, cols, generator=g) * 0.02
    X = torch.randn(4, 64, cols, generator=g)
    H, n = torch.zeros(cols, cols), 0
    for b in range(4):
        H, n = O.hessian_add_batch(H, n, X[b:b + 1])
    cfg = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format="int2", outlier_elem_format="fp4", block_size=16)
    layer = torch.nn.Linear(cols, rows, bias=False)
    layer.weight.data = W.clone()
    gq = _OracleGPTQ(layer, H, cfg)

    def solve(Wl, fac):
        Q, loss, info = O.fasterquant(Wl, fac["Hinv"], cfg, 128)
        cnt = torch.zeros(2 * cols + 3, dtype=torch.int32)
        cnt[:cols] = (Q == 0).sum(0).to(torch.int32)
        cnt[cols:2 * cols] = torch.tensor(info["k"], dtype=torch.int32)
        return Q, torch.tensor([loss], dtype=torch.float64), cnt
    path = D.fasterquant_row_sharded(gq, solve_fn=solve)
    Hinv, _ = O.hessian_inverse_factor(H.clone(), W.clone(), .01)
    Qref, loss_ref, info = O.fasterquant(W, Hinv, cfg, 128)
    ok = path == "row-sharded" and info["columns_with_zeros_lt_k"] == 0 and gq.replicated_calls == 0
    ok = ok and torch.equal(layer.weight.data, Qref) and abs(gq.last_error - loss_ref) <= 1e-9 * loss_ref
    # a quantizer whose prune is NOT a no-op must fall back to the replicated solver on every rank
    layer2 = torch.nn.Linear(cols, rows, bias=False)
    layer2.weight.data = W.clone()
    gq2 = _OracleGPTQ(layer2, H, cfg)

    def solve_bad(Wl, fac):
        Q, loss, cnt = solve(Wl, fac)
        if rank == 1:
            cnt[3] = 0                                        # column 3: no zeros on this rank's rows ...
            cnt[cols + 3] = 30000                             # ... and a k nobody can satisfy
        return Q, loss, cnt
    path2 = D.fasterquant_row_sharded(gq2, solve_fn=solve_bad)
    ok = ok and path2 == "replicated" and gq2.replicated_calls == 1
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_sharded_fasterquant_two_rank_gloo():
    """rows of one linear over 2 ranks == the unsharded oracle solve, bit for bit; counters are reduced before the
    prune check; a violated check sends every rank to the replicated solver"""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rowshard_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def _mover_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from msq_b200.llm.sequential import LayerMover
    torch.manual_seed(0)                          # the same host model on every rank

    def mk():
        return torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.LayerNorm(53), torch.nn.Linear(53, 11, bias=False)).half()
    layers = [mk(), mk()]
    want = [{n: p.detach().clone() for n, p in l.named_parameters()} for l in layers]

    def gather(out, mine):                        # gloo: all_gather on a list of views
        parts = list(out.chunk(world))
        dist.all_gather(parts, mine.clone())
    mv = LayerMover(layers, torch.device("cpu"), enabled=False, shard=True, gather=gather)
    ok = mv.world == world
    # the two halves of the sharded route on CPU tensors (source and destination are both the host here)
    for i, layer in enumerate(layers):
        named = list(layer.named_parameters())
        offs, chunk = LayerMover.shard_plan([p.numel() * p.element_size() for _, p in named], world)
        ok = ok and chunk % 256 == 0 and chunk * world >= offs[-1] + named[-1][1].numel() * 2
        stage = torch.zeros(chunk, dtype=torch.uint8)
        LayerMover.stage_shard([p.data for _, p in named], offs, rank * chunk, (rank + 1) * chunk, stage)
        buf = torch.empty(chunk * world, dtype=torch.uint8)
        gather(buf, stage)
        for (n, p), off in zip(named, offs):
            nb = p.numel() * p.element_size()
            got = buf[off: off + nb].view(p.dtype).reshape(p.shape)
            ok = ok and torch.equal(got, want[i][n])
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_layer_upload_two_rank_gloo():
    """LayerMover(shard=True): each rank stages 1/world of a layer's bytes, an all-gather assembles every parameter"""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_mover_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)
