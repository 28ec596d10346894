"""the ORCHESTRATION of xlmr.KernelFeatureExtractor (weight concatenation, mask, eps, pooler,
plan columns) checked on the CPU by substituting plain-torch stand-ins with the kernel Functions' contracts; the kernels
themselves are covered by the GPU suites. Result must equal the Hugging Face module it wraps."""
import math

import torch

from _util import pkg, rel_err


def _cell():
    from transformers import XLMRobertaConfig, XLMRobertaModel
    cfg = XLMRobertaConfig(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                           max_position_embeddings=64, type_vocab_size=1, pad_token_id=1)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    return XLMRobertaModel(cfg, add_pooling_layer=True).eval()


def test_text_encoder_wiring_matches_huggingface(monkeypatch):
    xl, Fn, ops = pkg("xlmr"), pkg("functional"), pkg("ops")

    def linear(x, w, b, act="none", engine=0):
        y = torch.nn.functional.linear(x, w, b)
        return torch.tanh(y) if act == "tanh" else y

    def mask_additive(mask, n):
        return (1.0 - mask[:, :n].float()) * -10000.0

    def folded_attention(plan, tensors, mask_add=None, bias=None):
        (slot, qc, Lq, _, _), = plan.roles["q"]
        (_, kc, Lk, _, _), = plan.roles["k"]
        (_, vc, _, _, _), = plan.roles["v"]
        t = tensors[slot]
        HD = plan.heads * plan.dh
        def heads(col, L):
            return t[:, col:col + HD].reshape(plan.NP, L, plan.heads, plan.dh).permute(0, 2, 1, 3)
        s = heads(qc, Lq) @ heads(kc, Lk).transpose(-1, -2) / math.sqrt(plan.dh)
        s = s + mask_add[torch.arange(plan.NP) // plan.mask_div][:, None, None, :Lk]
        return (torch.softmax(s, -1) @ heads(vc, Lk)).permute(0, 2, 1, 3).reshape(plan.NP * Lq, HD)

    def layer_tail(a, res, res_idx, res_inv, params, engine=0, drop1=None, drop2=None, out_every=0, eps=1e-12):
        wo, bo, g1, b1, w1, bi1, w2, bi2, g2, b2 = params
        ln = lambda s, g, b: torch.nn.functional.layer_norm(s, (s.shape[-1],), g, b, eps)
        x1 = ln(torch.nn.functional.linear(a, wo, bo) + res, g1, b1)
        h = torch.nn.functional.linear(x1, w1, bi1)
        h = h * 0.5 * (1.0 + torch.erf(h / math.sqrt(2.0)))
        return ln(torch.nn.functional.linear(h, w2, bi2) + x1, g2, b2)

    monkeypatch.setattr(Fn, "linear", linear)
    monkeypatch.setattr(Fn, "folded_attention", folded_attention)
    monkeypatch.setattr(Fn, "layer_tail", layer_tail)
    monkeypatch.setattr(ops, "mask_additive", mask_additive)

    cell = _cell()
    enc = xl.KernelFeatureExtractor(cell=cell).eval()
    g = torch.Generator().manual_seed(1)
    B, L = 4, 20
    ids = torch.randint(3, 200, (B, L), generator=g)
    mask = torch.ones(B, L, dtype=torch.int64)
    mask[1, 13:] = 0
    ids = torch.where(mask == 1, ids, torch.ones_like(ids))
    tt = torch.zeros_like(ids)

    class FakeCuda(torch.Tensor):
        pass
    # force the kernel branch on CPU tensors: is_cuda is only consulted on input_ids
    monkeypatch.setattr(type(ids), "is_cuda", property(lambda self: True), raising=False)
    try:
        with torch.no_grad():
            seq, pooled, att = enc(ids, tt, mask)
    finally:
        monkeypatch.undo()
    with torch.no_grad():
        ref = cell(input_ids=ids, token_type_ids=tt, attention_mask=mask)
    valid = mask.bool()
    assert att is None
    assert rel_err(seq[valid], ref[0][valid]) < 1e-5
    assert rel_err(pooled, ref[1]) < 1e-5
