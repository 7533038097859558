"""Perplexity fixture: a tiny random-init LLaMA quantized by the REFERENCE (its GPTQ class llm/gptq.py, its
MXQuantizer utils/quant.py and its RTN call quantize_mx_outlier_v1, all imported in place from /root/reference,
CPU) in the layer-sequential order of llm/llama.py:62-173, evaluated with llama_eval's perplexity arithmetic
(llm/llama.py:266-282) on synthetic tokens.  Run in the authoring container only:

    python tests/golden/make_golden_ppl.py

The reference's own llama_sequential cannot be called directly: it passes a fixed set of keyword arguments to the
decoder layers that the installed transformers no longer accepts; the loop below is that function's
body with the layer call made version-independent.  Writes tests/golden/ppl_tiny_llama.npz (model weights, token ids,
reference perplexities)."""
import contextlib
import copy
import io
import os
import sys
import warnings

import numpy as np
import torch
import transformers

warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [REF + "/number_system", REF, REF + "/llm"]
torch.cuda.synchronize = lambda *a, **k: None          # llm/gptq.py:173 calls it unconditionally

from utils.quant import MXQuantizer, quantize_mx_outlier_v1   # noqa: E402  (the reference's)
import gptq as ref_gptq                                        # noqa: E402  (the reference's)

CFG = dict(hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=4,
           vocab_size=256, max_position_embeddings=128)
SEQLEN, NCALIB, NTEST = 32, 16, 8
#: the reference's hard-coded quantizer, llm/llama.py:130-137
QARGS = (8, 8, "int2", "fp4")
QKW = dict(axes=[0], block_size=16)


def find_linears(module, prefix=""):
    out = {}
    for n, m in module.named_children():
        name = prefix + "." + n if prefix else n
        if isinstance(m, torch.nn.Linear):
            out[name] = m
        else:
            out.update(find_linears(m, name))
    return out


class Stop(Exception):
    pass


@torch.no_grad()
def layer_inputs(model, batches):
    store = {"x": [], "a": (), "k": {}}

    class Catcher(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, *a, **k):
            store["x"].append(x)
            store["a"], store["k"] = a, k
            raise Stop
    layers = model.model.layers
    layers[0] = Catcher(layers[0])
    model.config.use_cache = False
    for b in batches:
        try:
            model(b)
        except Stop:
            pass
    layers[0] = layers[0].m
    return torch.cat(store["x"], 0), store["a"], store["k"]


def run(layer, x, a, k):
    o = layer(x, *a, **k)
    return o[0] if isinstance(o, (tuple, list)) else o


@torch.no_grad()
def reference_gptq(model, calib):
    """llama_sequential (llm/llama.py:62-173) with the reference's classes"""
    inps, a, k = layer_inputs(model, calib)
    outs = torch.zeros_like(inps)
    for layer in model.model.layers:
        subset = find_linears(layer)
        g = {}
        for name, lin in subset.items():
            g[name] = ref_gptq.GPTQ(lin)
            g[name].quantizer = MXQuantizer()
            g[name].quantizer.configure(*QARGS, **QKW)
        hooks = [lin.register_forward_hook((lambda nm: lambda _, i, o: g[nm].add_batch(i[0].data, o.data))(name))
                 for name, lin in subset.items()]
        for j in range(inps.shape[0]):
            outs[j] = run(layer, inps[j].unsqueeze(0), a, k)
        for h in hooks:
            h.remove()
        for name in subset:
            with contextlib.redirect_stdout(io.StringIO()):
                g[name].fasterquant(percdamp=.01, groupsize=-1, actorder=False, static_groups=False)
            g[name].free()
        for j in range(inps.shape[0]):
            outs[j] = run(layer, inps[j].unsqueeze(0), a, k)
        inps, outs = outs, inps


@torch.no_grad()
def reference_rtn(model):
    """the --nearest branch, llm/llama.py:226-253"""
    for layer in model.model.layers:
        for lin in find_linears(layer).values():
            lin.weight.data = quantize_mx_outlier_v1(lin.weight.data, *QARGS, "max", 2, [0], 16)


@torch.no_grad()
def perplexity(model, ids):
    """llm/llama.py:266-282"""
    n = ids.numel() // SEQLEN
    nll = []
    for i in range(n):
        b = ids[:, i * SEQLEN:(i + 1) * SEQLEN]
        logits = model(b).logits
        loss = torch.nn.CrossEntropyLoss()(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), b[:, 1:].reshape(-1))
        nll.append(loss.float() * SEQLEN)
    return float(torch.exp(torch.stack(nll).sum() / (n * SEQLEN)))


def main():
    torch.manual_seed(0)
    base = transformers.LlamaForCausalLM(transformers.LlamaConfig(**CFG)).float().eval()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():      # heavier-tailed weights so that the outlier path matters
        for n, p in base.named_parameters():
            if "proj" in n:
                idx = torch.randperm(p.numel(), generator=g)[: max(1, p.numel() // 200)]
                p.view(-1)[idx] *= 8
        # peaky output distribution + test text generated greedily by the fp32 model itself: its perplexity is low
        # and moves visibly when the decoder linears are perturbed (uniformly random tokens give ~vocab_size whatever
        # the weights are)
        base.lm_head.weight.mul_(40.0)
    def generate(n):
        seqs = []
        for _ in range(n):
            ids = torch.randint(0, CFG["vocab_size"], (1, 4), generator=g)
            with torch.no_grad():
                while ids.shape[1] < SEQLEN:
                    ids = torch.cat([ids, base(ids).logits[:, -1].argmax(-1, keepdim=True)], 1)
            seqs.append(ids)
        return seqs
    calib = generate(NCALIB // 2) + [torch.randint(0, CFG["vocab_size"], (1, SEQLEN), generator=g) for _ in range(NCALIB // 2)]
    test = torch.cat(generate(NTEST), 1)
    ppl_fp = perplexity(base, test)
    m_g = copy.deepcopy(base)
    reference_gptq(m_g, calib)
    ppl_g = perplexity(m_g, test)
    m_r = copy.deepcopy(base)
    reference_rtn(m_r)
    ppl_r = perplexity(m_r, test)
    print("reference perplexities: fp32 %.4f  gptq %.4f  rtn %.4f" % (ppl_fp, ppl_g, ppl_r))
    out = {"sd__" + k: v.detach().numpy() for k, v in base.state_dict().items()}
    out.update(calib=torch.cat(calib, 0).numpy(), test=test.numpy(), ppl_fp=np.float64(ppl_fp), ppl_gptq=np.float64(ppl_g),
               ppl_rtn=np.float64(ppl_r))
    for li, layer in enumerate(m_g.model.layers):          # every reference-quantized linear
        for name, lin in find_linears(layer).items():
            out["gptq__%d.%s" % (li, name)] = lin.weight.detach().numpy()
    for li, layer in enumerate(m_r.model.layers):
        for name, lin in find_linears(layer).items():
            out["rtn__%d.%s" % (li, name)] = lin.weight.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "ppl_tiny_llama.npz"), **out)
    print("wrote ppl_tiny_llama.npz", os.path.getsize(os.path.join(HERE, "ppl_tiny_llama.npz")), "bytes")


if __name__ == "__main__":
    main()
