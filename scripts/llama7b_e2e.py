"""End-to-end GPTQ quantization of a random-init LLaMA-2-7B-shaped model with synthetic calibration
tokens (128 x 2048), through msq_b200.llm.sequential (the llama_sequential mirror).  Prints one JSON line.
    python scripts/llama7b_e2e.py [--layers 32] [--nsamples 128]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import transformers

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--nsamples", type=int, default=128)
ap.add_argument("--seqlen", type=int, default=2048)
ap.add_argument("--in-fmt", default="fp4_e2m1")
ap.add_argument("--out-fmt", default="fp8_e4m3")
ap.add_argument("--nearest", action="store_true")
a = ap.parse_args()

from msq_b200.llm import sequential as S
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
t0 = time.time()
cfg = transformers.LlamaConfig(hidden_size=4096, intermediate_size=11008, num_hidden_layers=a.layers, num_attention_heads=32,
                               num_key_value_heads=32, vocab_size=32000, max_position_embeddings=4096)
skip = lambda *x, **k: None
saved = (torch.nn.init.kaiming_uniform_, torch.nn.init.uniform_, torch.nn.init.normal_)
torch.nn.init.kaiming_uniform_ = torch.nn.init.uniform_ = torch.nn.init.normal_ = skip
torch.set_default_dtype(torch.float16)
with transformers.modeling_utils.no_init_weights() if hasattr(transformers.modeling_utils, "no_init_weights") else torch.no_grad():
    model = transformers.LlamaForCausalLM(cfg)
torch.set_default_dtype(torch.float32)
torch.nn.init.kaiming_uniform_, torch.nn.init.uniform_, torch.nn.init.normal_ = saved
g = torch.Generator(device=dev).manual_seed(0)
with torch.no_grad():
    for n, p in model.named_parameters():          # random-init weights of the architecture (no checkpoint available)
        if p.dim() >= 2:
            p.data = (torch.randn(p.shape, device=dev, generator=g) * 0.02).half().cpu()
        else:
            p.data = torch.ones_like(p.data)
model.eval()
model.seqlen = a.seqlen
t_build = time.time() - t0
gc = torch.Generator().manual_seed(3)
calib = [torch.randint(0, 32000, (1, a.seqlen), generator=gc) for _ in range(a.nsamples)]
qcfg = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format=a.in_fmt, outlier_elem_format=a.out_fmt,
            axes=[0], block_size=16)
stats = {}
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.time()
S.sequential_quantize(model, calib, dev, qcfg=qcfg, nearest=a.nearest, stats=stats)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t_q = time.time() - t0
if world > 1:
    # every rank must end with the same quantized model
    w = model.model.layers[-1].mlp.down_proj.weight.data.to(dev).float()
    chk = torch.stack([w.sum(), w.abs().sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    stats["ranks_agree"] = bool(torch.equal(lo, hi))
    tt = torch.tensor([t_q], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_q = float(tt)
    dist.destroy_process_group()
if rank == 0:
  print(json.dumps({"n_gpus": world, "model": "LLaMA-2-7B shape, random init, %d layers" % a.layers, "nsamples": a.nsamples, "seqlen": a.seqlen,
                  "quantizer": "%s/%s block 16 axes=[0]" % (a.in_fmt, a.out_fmt), "mode": "rtn" if a.nearest else "gptq",
                  "quantize_s": t_q, "build_model_s": t_build, **stats}))
