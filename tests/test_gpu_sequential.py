"""GPU tier: the layer-sequential driver on random-init HF models.
The tolerance clause of the north star ("validated against the reference's perplexity on synthetic
calibration data") is checked here: quantizing with the CUDA path and with the oracle (the
reference's arithmetic on the CPU) must give the same perplexity on synthetic tokens."""
import copy

import pytest
import torch

import msq_oracle as O

pytestmark = pytest.mark.gpu
transformers = pytest.importorskip("transformers")


def tiny_llama(seed=0):
    cfg = transformers.LlamaConfig(hidden_size=128, intermediate_size=352, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=4, vocab_size=512, max_position_embeddings=256)
    torch.manual_seed(seed)
    m = transformers.LlamaForCausalLM(cfg).float().eval()
    m.seqlen = 64
    return m


def tiny_opt(seed=0):
    cfg = transformers.OPTConfig(hidden_size=128, ffn_dim=256, num_hidden_layers=2, num_attention_heads=4, vocab_size=512,
                                 word_embed_proj_dim=128, max_position_embeddings=256)
    torch.manual_seed(seed)
    m = transformers.OPTForCausalLM(cfg).float().eval()
    m.seqlen = 64
    return m


@pytest.mark.timeout(600)
@pytest.mark.parametrize("make", [tiny_llama, tiny_opt])
def test_gptq_and_rtn_drivers(make):
    from msq_b200.llm import sequential as S
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    calib = [torch.randint(0, 512, (1, 64), generator=g) for _ in range(8)]
    test_ids = torch.randint(0, 512, (1, 64 * 4), generator=g)
    qcfg = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format="fp4_e2m1", outlier_elem_format="fp8_e4m3",
                axes=[0], block_size=16)
    base = make()
    ppl_fp = S.perplexity(copy.deepcopy(base), test_ids, dev, 64)

    # RTN: CUDA path vs the reference arithmetic (oracle) applied to the same model
    m_rtn = copy.deepcopy(base)
    S.sequential_quantize(m_rtn, None, dev, qcfg=qcfg, nearest=True)
    m_ref = copy.deepcopy(base)
    layers, _, _, _ = S.decoder_parts(m_ref)
    for layer in layers:
        for name, lin in S.find_layers(layer).items():
            lin.weight.data = O.quantize_mx_outlier(lin.weight.data.float(), 8, 8, "fp4_e2m1", "fp8_e4m3", 2, 0, 16)
    for (n1, p1), (n2, p2) in zip(m_rtn.named_parameters(), m_ref.named_parameters()):
        assert torch.equal(p1.cpu(), p2.cpu()), n1                    # bit-exact weights => identical model
    ppl_rtn = S.perplexity(m_rtn, test_ids, dev, 64)
    ppl_ref = S.perplexity(m_ref, test_ids, dev, 64)
    assert abs(ppl_rtn - ppl_ref) <= 1e-6 * ppl_ref

    # GPTQ driver: runs, uses the hot path for every linear, stays finite and close to the fp model
    m_g = copy.deepcopy(base)
    stats = {}
    q = S.sequential_quantize(m_g, calib, dev, qcfg=qcfg, stats=stats)
    n_lin = sum(len(S.find_layers(l)) for l in S.decoder_parts(m_g)[0])
    assert len(q) == n_lin == stats["linears"]
    ppl_g = S.perplexity(m_g, test_ids, dev, 64)
    assert ppl_g == ppl_g and ppl_g < 4 * ppl_fp
    # weights changed and live on the MX grid: re-quantizing a GPTQ'd weight column block is stable in value range
    w0 = dict(base.named_parameters())
    changed = [n for n, p in m_g.named_parameters() if "proj" in n or "fc" in n]
    assert any(not torch.equal(dict(m_g.named_parameters())[n].cpu(), w0[n].cpu()) for n in changed)


@pytest.mark.timeout(600)
def test_driver_options_do_not_change_the_result():
    """overlapped layer moves (background thread + side stream), side streams for the independent solves and the
    packed-artefact collection are scheduling choices: the quantized model must be bit-identical with and without"""
    from msq_b200.llm import sequential as S
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    calib = [torch.randint(0, 512, (1, 64), generator=g) for _ in range(8)]
    qcfg = dict(inlier_scale_bits=8, outlier_scale_bits=8, inlier_elem_format="int2", outlier_elem_format="fp4",
                axes=[0], block_size=16)
    base = tiny_llama(7)
    outs = []
    packed = {}
    for kw in (dict(overlap_moves=False, solve_streams_n=0), dict(overlap_moves=True, solve_streams_n=3),
               dict(overlap_moves=True, solve_streams_n=2, packed_out=packed), dict(overlap_moves=False, share_hessians=False)):
        m = copy.deepcopy(base)
        S.sequential_quantize(m, calib, dev, qcfg=qcfg, **kw)
        outs.append({n: p.detach().cpu().clone() for n, p in m.named_parameters()})
        assert all(not p.is_cuda for p in m.parameters())              # every layer is back on the host
    for o in outs[1:]:
        for n in outs[0]:
            assert torch.equal(outs[0][n], o[n]), n
    # the collected artefacts decode to the weights the driver wrote
    assert len(packed) > 0
    for name, pw in packed.items():
        w = outs[2]["model." + name + ".weight"]
        assert torch.equal(pw.dequantize().cpu(), w.float()), name


def _fixture_model():
    import numpy as np
    import os
    from util import GOLD
    z = np.load(os.path.join(GOLD, "ppl_tiny_llama.npz"))
    cfg = transformers.LlamaConfig(hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=4, vocab_size=256, max_position_embeddings=128)
    m = transformers.LlamaForCausalLM(cfg).float().eval()
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}
    m.load_state_dict(sd)
    m.seqlen = 32
    return m, z


@pytest.mark.timeout(600)
def test_perplexity_against_the_reference_fixture():
    """north star: "tolerance validated against the reference's perplexity on synthetic calibration data".
    tests/golden/ppl_tiny_llama.npz holds a tiny LLaMA, synthetic tokens, and the perplexities obtained by running the
    REFERENCE's GPTQ / MXQuantizer / quantize_mx_outlier_v1 on it (tests/golden/make_golden_ppl.py, int2 inliers / fp4
    outliers / block 16 = llm/llama.py:130-137).  fp32 9.82, reference GPTQ 25.09, reference RTN 32.39."""
    from msq_b200.llm import sequential as S
    dev = torch.device("cuda:0")
    base, z = _fixture_model()
    calib = [torch.from_numpy(z["calib"][i:i + 1]) for i in range(z["calib"].shape[0])]
    test_ids = torch.from_numpy(z["test"])
    ppl_fp = S.perplexity(copy.deepcopy(base), test_ids, dev, 32)
    assert abs(ppl_fp - float(z["ppl_fp"])) <= 1e-3 * float(z["ppl_fp"])          # same model, same arithmetic

    # RTN: bit-exact weights, so the perplexity differs only by the fp32 forward's summation order (CPU vs GPU)
    m_r = copy.deepcopy(base)
    S.sequential_quantize(m_r, None, dev, qcfg=S.REFERENCE_QCFG, nearest=True)
    for li, layer in enumerate(m_r.model.layers):
        for name, lin in S.find_layers(layer).items():
            assert torch.equal(lin.weight.detach().cpu(), torch.from_numpy(z["rtn__%d.%s" % (li, name)])), (li, name)
    ppl_r = S.perplexity(m_r, test_ids, dev, 32)
    assert abs(ppl_r - float(z["ppl_rtn"])) <= 1e-3 * float(z["ppl_rtn"]), (ppl_r, float(z["ppl_rtn"]))

    def gptq_run(topk_device):
        import msq_b200
        old = msq_b200.GPTQ.prune_topk_device
        msq_b200.GPTQ.prune_topk_device = topk_device
        try:
            m_g = copy.deepcopy(base)
            S.sequential_quantize(m_g, calib, dev, qcfg=S.REFERENCE_QCFG, hessian_batch=1)
        finally:
            msq_b200.GPTQ.prune_topk_device = old
        diff = {}
        for li, layer in enumerate(m_g.model.layers):
            for name, lin in S.find_layers(layer).items():
                ref = torch.from_numpy(z["gptq__%d.%s" % (li, name)])
                diff["%d.%s" % (li, name)] = float((lin.weight.detach().cpu() != ref).float().mean())
        return S.perplexity(m_g, test_ids, dev, 32), diff

    # GPTQ, tie order of the prune's torch.topk taken from the CPU like in the fixture's (CPU) run of the reference:
    # the six single-block linears of layer 0 are bit-identical to the reference's, later ones see inputs that differ in
    # the last bits (tensor-core Hessian, GPU forward) and stay within 2 % differing entries; perplexity within 2 %.
    ppl_c, diff_c = gptq_run("cpu")
    # GPTQ as shipped (CUDA tie order = what the reference itself gets when it runs on a GPU): same algorithm, other
    # arbitrary choice among tied prune candidates; the perplexity must stay in the reference's range (better than RTN).
    ppl_g, diff_g = gptq_run("cuda")
    print("perplexity fp32 %.4f | RTN ours %.4f ref %.4f | GPTQ ref %.4f, ours cpu-ties %.4f (max differing %.4f), "
          "ours cuda-ties %.4f (max differing %.4f)" % (ppl_fp, ppl_r, float(z["ppl_rtn"]), float(z["ppl_gptq"]), ppl_c,
                                                        max(diff_c.values()), ppl_g, max(diff_g.values())))
    for name in ("0.self_attn.q_proj", "0.self_attn.k_proj", "0.self_attn.v_proj", "0.self_attn.o_proj", "0.mlp.gate_proj",
                 "0.mlp.up_proj"):
        assert diff_c[name] == 0.0 and diff_g[name] == 0.0, (name, diff_c[name], diff_g[name])
    assert max(diff_c.values()) <= 0.02, diff_c
    assert abs(ppl_c - float(z["ppl_gptq"])) <= 0.02 * float(z["ppl_gptq"]), (ppl_c, float(z["ppl_gptq"]))
    assert ppl_fp < ppl_g < ppl_r and abs(ppl_g - float(z["ppl_gptq"])) <= 0.25 * float(z["ppl_gptq"])
