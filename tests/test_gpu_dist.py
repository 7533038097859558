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


def _seq_worker(rank, world, port, ret):
    import torch.distributed as dist
    import transformers
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from msq_b200.llm import sequential as S
    torch.manual_seed(0)
    cfg = transformers.LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=4, vocab_size=512, max_position_embeddings=256)
    model = transformers.LlamaForCausalLM(cfg).half().eval()
    model.seqlen = 128
    before = {n: p.detach().clone() for n, p in model.named_parameters() if "proj" in n}
    g = torch.Generator().manual_seed(1)
    calib = [torch.randint(0, 512, (1, 128), generator=g) for _ in range(6)]
    qcfg = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format="fp4_e2m1", outlier_elem_format="fp8_e4m3",
                axes=[0], block_size=16)
    stats = {}
    S.sequential_quantize(model, calib, dev, qcfg=qcfg, stats=stats)
    ok = True
    rel = []
    for n, p in model.named_parameters():
        if "proj" not in n:
            continue
        w = p.detach().to(dev).float()
        ok &= bool(torch.isfinite(w).all())
        lo, hi = w.clone(), w.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok &= bool(torch.equal(lo, hi))                      # every rank holds the same quantized weight
        b = before[n].to(dev).float()
        rel.append(float((w - b).norm() / b.norm()))
    ok &= 0.0 < max(rel) < 0.5                               # quantized (changed) but close to the original
    # the linears of a layer were spread over the ranks, and every calibration sample was used exactly once
    tot = torch.tensor([float(stats["linears"])], device=dev)
    dist.all_reduce(tot)
    ok &= int(tot.item()) == 2 * 7
    ret[rank] = ok
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sequential_driver_data_parallel():
    """llm/llama.py:62-173 under torch.distributed: samples sharded, Hessians allreduced, solves spread over ranks"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_seq_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def _rowshard_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import msq_b200
    from msq_b200 import dist as D
    g = torch.Generator().manual_seed(11)
    rows, cols = 1536, 640
    W = torch.randn(rows, cols, generator=g) * 0.02
    X = torch.randn(4, 256, cols, generator=g).half()

    def make():
        layer = torch.nn.Linear(cols, rows, bias=False).to(dev)
        layer.weight.data = W.clone().to(dev)
        gq = msq_b200.GPTQ(layer)
        gq.verbose = False
        gq.quantizer = msq_b200.MXQuantizer()
        gq.quantizer.configure(8, 8, "int2", "fp4", axes=[0], block_size=16)
        return layer, gq
    layer, gq = make()
    lo, hi = D.shard_range(4, rank, world)
    for s in range(lo, hi):
        gq.add_batch(X[s:s + 1].to(dev), None)
    D.allreduce_gptq_hessian(gq)
    Hsave = gq.H.clone()
    path = D.fasterquant_row_sharded(gq)
    # reference: the unsharded solver on the same (allreduced) Hessian
    layer2, gq2 = make()
    gq2.H, gq2.nsamples = Hsave, 4
    gq2.fasterquant()
    ok = path == "row-sharded" and gq2.last_path == "fused"
    ok &= bool(torch.equal(layer.weight.data, layer2.weight.data))          # rows never interact: bit-identical
    ok &= abs(gq.last_error - gq2.last_error) <= 1e-9 * gq2.last_error
    ret[rank] = ok
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_sharded_fasterquant_nccl():
    """SURVEY 8e row 2 on two GPUs: rows of one linear over the ranks + counter allreduce + all-gather == unsharded"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_rowshard_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def test_entry_points_follow_the_tensor_device():
    """a tensor on cuda:1 while cuda:0 is current must run on cuda:1 (device guard inside the C ABI), and
    the shared-memory opt-in of the big kernels must work on a second device of the same process"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import msq_b200
    from msq_b200 import _tc
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(0)
    W = torch.randn(256, 128, generator=g) * 0.02
    y0 = msq_b200.quantize_mx_outlier_v1(W.to("cuda:0"), 8, 8, "int2", "fp4", "max", 2, [0], 16)
    y1 = msq_b200.quantize_mx_outlier_v1(W.to("cuda:1"), 8, 8, "int2", "fp4", "max", 2, [0], 16)
    assert y1.device.index == 1 and torch.equal(y0.cpu(), y1.cpu())
    A = torch.randn(300, 200, generator=g)
    c0 = _tc.gemm_split3(A.to("cuda:0"), A.to("cuda:0"))
    c1 = _tc.gemm_split3(A.to("cuda:1"), A.to("cuda:1"))
    assert c1.device.index == 1 and torch.equal(c0.cpu(), c1.cpu())
    assert torch.cuda.current_device() == 0
