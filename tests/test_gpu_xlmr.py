"""The kernel-backed XLM-R encoder against the stock Hugging Face module it wraps, same
weights and inputs, eval() mode; fp32 (CUDA-core engines) at 1e-4 on states and gradients, bf16 at the bf16 bar."""
import pytest
import torch

from _util import pkg, rel_err

pytestmark = pytest.mark.gpu


def _cell(hidden=128, heads=2, inter=256, layers=2, vocab=200):
    from transformers import XLMRobertaConfig, XLMRobertaModel
    cfg = XLMRobertaConfig(vocab_size=vocab, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                           intermediate_size=inter, max_position_embeddings=64, type_vocab_size=1, pad_token_id=1)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    return XLMRobertaModel(cfg, add_pooling_layer=True)


@pytest.mark.parametrize("dtype,tol", [(None, 1e-4), (torch.bfloat16, 3e-2)])
def test_kernel_text_encoder_matches_huggingface(dtype, tol):
    xl = pkg("xlmr")
    cell = _cell().cuda().eval()
    enc = xl.KernelFeatureExtractor(cell=cell).eval()
    enc.compute_dtype = dtype
    enc.engine = pkg("_lib").ENGINE_SIMT if dtype is None else 0
    g = torch.Generator().manual_seed(1)
    B, L = 5, 24
    ids = torch.randint(3, 200, (B, L), generator=g).cuda()
    mask = torch.ones(B, L, dtype=torch.int64)
    mask[1, 17:] = 0
    mask[3, 9:] = 0
    ids = torch.where(mask.cuda() == 1, ids, torch.ones_like(ids))               # pad_token_id = 1 where masked
    tt = torch.zeros_like(ids)
    ref = cell(input_ids=ids, token_type_ids=tt, attention_mask=mask.cuda())
    want_seq, want_pool = ref[0], ref[1]
    w = torch.randn(want_seq.shape, generator=g).cuda() * mask.cuda().unsqueeze(-1)  # padded positions carry no loss
    (want_seq * w).sum().backward()
    g_ref = {k: v.grad.clone() for k, v in cell.named_parameters() if v.grad is not None}
    cell.zero_grad()
    if dtype is None:
        pkg("ops").set_attn_engine(pkg("_lib").ENGINE_SIMT)
    try:
        seq, pooled, att = enc(ids, tt, mask.cuda())
        (seq.float() * w).sum().backward()
    finally:
        pkg("ops").set_attn_engine(0)
    valid = mask.cuda().bool()
    assert att is None
    assert rel_err(seq.float()[valid], want_seq[valid]) < tol
    assert rel_err(pooled.float(), want_pool) < tol
    gmax = max(float(v.abs().max()) for v in g_ref.values())
    for k, v in cell.named_parameters():
        if k in g_ref and "pooler" not in k:
            err = float((v.grad - g_ref[k]).abs().max()) / max(float(g_ref[k].abs().max()), 1e-3 * gmax)
            assert err < (tol if dtype is None else 0.15), (k, err)
