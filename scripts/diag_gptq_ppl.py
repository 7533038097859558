"""diagnostic: layer-0 down_proj (K=176: two GPTQ blocks) vs the reference fixture"""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch, transformers
import msq_b200, msq_oracle as O
from msq_b200.llm import sequential as S
z = np.load(os.path.join(ROOT, "tests/golden/ppl_tiny_llama.npz"))
cfg = transformers.LlamaConfig(hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                               num_key_value_heads=4, vocab_size=256, max_position_embeddings=128)
m = transformers.LlamaForCausalLM(cfg).float().eval()
m.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")})
dev = torch.device("cuda:0")
msq_b200.GPTQ.verbose = False
calib = [torch.from_numpy(z["calib"][i:i + 1]) for i in range(z["calib"].shape[0])]
inps, a, k = S.capture_layer_inputs(m, calib, dev)
layer = copy.deepcopy(m.model.layers[0]).to(dev)
lin = layer.mlp.down_proj
X = []
h = lin.register_forward_hook(lambda _, i, o: X.append(i[0].data.clone()))
for j in range(inps.shape[0]):
    S._run_layer(layer, inps[j].unsqueeze(0), a, k)
h.remove()
K = 176
ref_q = torch.from_numpy(z["gptq__0.mlp.down_proj"])
Hc = torch.zeros(K, K); n = 0
for x in X:
    xx = x.reshape(-1, K).t().cpu().float()
    Hc *= n / (n + 1); n += 1
    xx = (2 / n) ** 0.5 * xx
    Hc += xx @ xx.t()
def run(Hinject=None):
    l2 = copy.deepcopy(lin)
    g = msq_b200.GPTQ(l2); g.hessian_batch = 1
    g.quantizer = msq_b200.MXQuantizer(); g.quantizer.configure(8, 8, "int2", "fp4", axes=[0], block_size=16)
    for x in X: g.add_batch(x, None)
    Hours = g.H.clone()
    if Hinject is not None: g.H = Hinject.to(dev).clone()
    g.fasterquant()
    return l2.weight.data.cpu(), Hours.cpu(), g.last_path, g.last_error
q1, Hours, path, err1 = run()
print("H rel err ours vs cpu fp32:", float((Hours - Hc).abs().max() / Hc.abs().max()), "path", path, "loss", err1)
print("ours (own H)      vs ref: differing", float((q1 != ref_q).float().mean()))
q2, _, path2, err2 = run(Hc)
print("ours (cpu H)      vs ref: differing", float((q2 != ref_q).float().mean()), path2, "loss", err2)
Hd = Hc.clone(); Hd[torch.arange(K), torch.arange(K)] += 0.01 * torch.mean(torch.diag(Hd))
Hinv = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hd)), upper=True)
cfgq = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format="int2", outlier_elem_format="fp4", block_size=16)
Qo, loss_o, info = O.fasterquant(lin.weight.data.cpu().float(), Hinv, cfgq, 128)
print("oracle (cpu H)    vs ref: differing", float((Qo != ref_q).float().mean()), "zeros<k cols", info["columns_with_zeros_lt_k"], "loss", loss_o)
print("ours (cpu H) vs oracle  : differing", float((q2 != Qo).float().mean()))
fd = lambda A, B: (int((A != B).any(0).float().argmax()) if (A != B).any() else None)
print("first differing column ours-vs-oracle:", fd(q2, Qo), "| ours-vs-ref:", fd(q2, ref_q), "| oracle-vs-ref:", fd(Qo, ref_q))
print("dead columns (H diag == 0):", int((torch.diag(Hc) == 0).sum()), " min/max diag", float(torch.diag(Hc).min()), float(torch.diag(Hc).max()))
