"""GPU tier, >= 2 GPUs (skipped on a single-GPU box): calibration samples split over ranks, per-rank
tcgen05 Hessians, ONE NCCL sum-allreduce (dist.combine_hessians) == the single-GPU Hessian."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import msq_b200
    from msq_b200 import dist as D
    msq_b200.GPTQ.verbose = False
    K, T, nsamp = 1024, 512, 6
    g = torch.Generator().manual_seed(0)
    X = torch.randn(nsamp, T, K, generator=g).half()
    layer = torch.nn.Linear(K, 256, bias=False).cuda().half()
    gq = msq_b200.GPTQ(layer)
    lo, hi = D.shard_range(nsamp, rank, world)
    for s in range(lo, hi):
        gq.add_batch(X[s:s + 1].cuda(), None)
    D.allreduce_gptq_hessian(gq)
    ref = msq_b200.GPTQ(layer)
    for s in range(nsamp):
        ref.add_batch(X[s:s + 1].cuda(), None)
    err = float((gq.H - ref.H).abs().max() / ref.H.abs().max())
    ret[rank] = (gq.nsamples == nsamp) and err < 1e-5 and bool(torch.equal(gq.H, gq.H.t()))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_hessian_allreduce_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)
