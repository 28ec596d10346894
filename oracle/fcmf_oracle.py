"""ORACLE -- test infrastructure, not product code.

A CPU, fp32, functional restatement of the reference's FCMF fusion path
(sonbui25/Multimodal-Aspect-Category-Sentiment-Analysis), exactly as the reference
executes it: one Python pass per aspect (run_multimodal_fcmf.py:462-478), one Python
pass per image (fcmf_pretraining.py:47-125), every row computed, nothing hoisted.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module. The product package never does.

Parity pin: the reference has no tests and no golden vectors, so
this restatement is pinned against outputs of the reference code itself, imported from
/root/reference by ``oracle/make_golden.py`` and committed under ``tests/golden``
(``tests/test_oracle_golden.py`` checks them on every CPU run).

All arithmetic is eval-mode (dropout = identity): parity is only definable there
(SURVEY.md section 8(c)). Each function cites the reference file:line it restates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

NEG_MASK = -10000.0           # fcmf_pretraining.py:56,100,136
LN_EPS = 1e-12                # mm_modeling.py:159,272,320


# ------------------------------------------------------------------ primitives
def erf_gelu(x: Tensor) -> Tensor:
    """mm_modeling.py:10-15 -- x * 0.5 * (1 + erf(x / sqrt(2)))."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def tf_layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float = LN_EPS) -> Tensor:
    """mm_modeling.py:166-171 -- biased variance, epsilon inside the square root."""
    mu = x.mean(-1, keepdim=True)
    var = (x - mu).pow(2).mean(-1, keepdim=True)
    return weight * ((x - mu) / torch.sqrt(var + eps)) + bias


def affine(x: Tensor, p: Params, name: str) -> Tensor:
    """nn.Linear: y = x W^T + b."""
    return torch.nn.functional.linear(x, p[name + ".weight"], p[name + ".bias"])


def extended_mask(added_attention_mask: Tensor, n: int, dtype=torch.float32) -> Tensor:
    """fcmf_pretraining.py:53-56 / 97-100 / 133-136 -- (1 - m[:, :n]) * -10000 as [B,1,1,n].

    The ROI and fusion masks stay int64 in the reference (``.to(dtype=self.dtype)`` is a no-op,
    lines 99 and 135) and are promoted when added to the scores; values are identical."""
    m = added_attention_mask[:, :n].unsqueeze(1).unsqueeze(2).to(dtype)
    return (1.0 - m) * NEG_MASK


def split_heads(x: Tensor, heads: int) -> Tensor:
    """mm_modeling.py:188-191 -- [B,T,H] -> [B,heads,T,dh]."""
    b, t, h = x.shape
    return x.view(b, t, heads, h // heads).permute(0, 2, 1, 3)


def multihead_attention(q_in: Tensor, kv_in: Tensor, add_mask: Tensor, p: Params, prefix: str, heads: int) -> Tensor:
    """BertSelfAttention.forward (mm_modeling.py:193-219) when q_in is kv_in, BertCoAttention.forward
    (mm_modeling.py:240-266) otherwise. Scale is applied BEFORE the mask add (lines 204-206)."""
    q = split_heads(affine(q_in, p, prefix + ".query"), heads)
    k = split_heads(affine(kv_in, p, prefix + ".key"), heads)
    v = split_heads(affine(kv_in, p, prefix + ".value"), heads)
    dh = q.shape[-1]
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)
    scores = scores + add_mask
    probs = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
    return ctx.view(ctx.shape[0], ctx.shape[1], -1)


def attention_block(q_in: Tensor, kv_in: Tensor, add_mask: Tensor, p: Params, prefix: str, heads: int) -> Tensor:
    """BertAttention / BertCrossAttention (mm_modeling.py:283-303): attention then
    BertSelfOutput (mm_modeling.py:276-280): LN(dense(ctx) + input)."""
    ctx = multihead_attention(q_in, kv_in, add_mask, p, prefix + ".self", heads)
    dense = affine(ctx, p, prefix + ".output.dense")
    return tf_layer_norm(dense + q_in, p[prefix + ".output.LayerNorm.weight"], p[prefix + ".output.LayerNorm.bias"])


def encoder_layer(q_in: Tensor, kv_in: Tensor, add_mask: Tensor, p: Params, prefix: str, heads: int) -> Tensor:
    """BertLayer (mm_modeling.py:338-342) / BertCrossAttentionLayer (mm_modeling.py:351-355):
    attention block -> BertIntermediate (311-314) -> BertOutput (324-328)."""
    a = attention_block(q_in, kv_in, add_mask, p, prefix + ".attention", heads)
    inter = erf_gelu(affine(a, p, prefix + ".intermediate.dense"))
    out = affine(inter, p, prefix + ".output.dense")
    return tf_layer_norm(out + a, p[prefix + ".output.LayerNorm.weight"], p[prefix + ".output.LayerNorm.bias"])


def first_token_pooler(x: Tensor, p: Params, prefix: str) -> Tensor:
    """BertPooler.forward (mm_modeling.py:425-431): tanh(dense(x[:, 0]))."""
    return torch.tanh(affine(x[:, 0], p, prefix + ".dense"))


# ------------------------------------------------------------------ geometric ROI attention
def box_relational_embedding(boxes: Tensor, dim_g: int = 64, wave_len: float = 1000.0) -> Tensor:
    """BoxMultiHeadedAttention.BoxRelationalEmbedding (roi_modeling.py:79-138).

    boxes: [B, NR, 4] as (x_min, x_max, y_min, y_max) (line 95), float64 in the reference data
    path. Output [B, NR, NR, 64] in the dtype of ``boxes``; entry (i, j) describes box i relative
    to box j. The frequency table is built in float32 (``torch.arange(dim_g / 8)`` is a float32
    arange, lines 123-125) and only then promoted against the float64 positions."""
    b = boxes.shape[0]
    x_min, x_max, y_min, y_max = torch.chunk(boxes, 4, dim=-1)
    cx, cy = (x_min + x_max) * 0.5, (y_min + y_max) * 0.5
    w, h = (x_max - x_min) + 1.0, (y_max - y_min) + 1.0
    dx = torch.log(torch.clamp(torch.abs((cx - cx.view(b, 1, -1)) / w), min=1e-3))
    dy = torch.log(torch.clamp(torch.abs((cy - cy.view(b, 1, -1)) / h), min=1e-3))
    dw = torch.log(w / w.view(b, 1, -1))
    dh = torch.log(h / h.view(b, 1, -1))
    pos = torch.stack((dx, dy, dw, dh), dim=-1)                       # [B,NR,NR,4]
    freq = torch.arange(dim_g / 8)                                    # float32 0..7
    freq = 1.0 / torch.pow(wave_len, freq / (dim_g / 8))              # float32
    arg = (100.0 * pos).unsqueeze(-1) * freq.view(1, 1, 1, 1, -1)     # [B,NR,NR,4,8]
    arg = arg.reshape(b, pos.shape[1], pos.shape[2], -1)              # [B,NR,NR,32]
    return torch.cat((torch.sin(arg), torch.cos(arg)), dim=-1)


def box_multihead_attention(x: Tensor, boxes: Tensor, p: Params, prefix: str, heads: int = 8) -> Tensor:
    """BoxMultiHeadedAttention.forward with q = k = v = x and mask=None
    (roi_modeling.py:140-180) + box_attention (roi_modeling.py:14-47)."""
    b, nr, hid = x.shape
    dk = hid // heads
    emb = box_relational_embedding(boxes).to(x.dtype)                 # line 149 cast
    q, k, v = [affine(x, p, f"{prefix}.linears.{i}").view(b, nr, heads, dk).transpose(1, 2) for i in range(3)]
    flat = emb.view(-1, emb.shape[-1])
    geo = [affine(flat, p, f"{prefix}.WGs.{i}").view(b, 1, nr, nr) for i in range(heads)]
    geo = torch.relu(torch.cat(geo, dim=1))                           # lines 160-162
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)     # roi_modeling.py:29-30
    w_mn = torch.softmax(torch.log(torch.clamp(geo, min=1e-6)) + scores, dim=-1)   # lines 40-41
    out = torch.matmul(w_mn, v).transpose(1, 2).contiguous().view(b, nr, hid)
    return affine(out, p, f"{prefix}.linears.3")


# ------------------------------------------------------------------ the fusion encoder
def fusion_encoder(sequence_output: Tensor, visual_embeds_att: Tensor, roi_embeds_att: Tensor,
                   roi_coors: Tensor, added_attention_mask: Tensor, p: Params,
                   heads: int, num_imgs: int, num_roi: int, pre: str = "encoder.") -> Tensor:
    """FCMFEncoder.forward after the text encoder (fcmf_pretraining.py:42-141).

    sequence_output [B,L,H] stands for ``self.bert(...)[0]`` (line 41). Returns [B, 1+2*NI, H]."""
    seq_len = sequence_output.shape[1]
    h_list: List[Tensor] = []
    r_list: List[Tensor] = []
    for i in range(num_imgs):
        # A. image-guided attention (lines 49-56, 84-92)
        img = affine(visual_embeds_att[:, i, :], p, pre + "vismap2text")
        img_mask = extended_mask(added_attention_mask, 49, img.dtype)
        t2i = encoder_layer(sequence_output, img, img_mask, p, pre + "text2img_attention.layer.0", heads)
        h_list.append(first_token_pooler(t2i, p, pre + "text2img_pooler").unsqueeze(1))
        # D. geometric ROI-aware attention (lines 97-123)
        roi_mask = extended_mask(added_attention_mask, seq_len + num_roi, sequence_output.dtype)
        roi = affine(roi_embeds_att[:, i, :], p, pre + "roimap2text")
        rel = box_multihead_attention(roi, roi_coors[:, i, :], p, pre + "box_head")
        text_roi = torch.cat((sequence_output, rel), dim=1)
        mm = encoder_layer(text_roi, text_roi, roi_mask, p, pre + "mm_attention.layer.0", heads)
        r_list.append(first_token_pooler(mm, p, pre + "text2roi_pooler").unsqueeze(1))
    fusion = torch.cat([sequence_output[:, 0:1, :]] + h_list + r_list, dim=1)       # lines 127-131
    fuse_mask = extended_mask(added_attention_mask, 1 + 2 * num_imgs, fusion.dtype)  # lines 133-136
    return encoder_layer(fusion, fusion, fuse_mask, p, pre + "mm_attention.layer.0", heads)   # line 139


def classifier_head(fused: Tensor, p: Params) -> Tensor:
    """FCMF.forward tail (fcmf_multimodal.py:48-50): text_pooler -> dropout(eval) -> classifier."""
    return affine(first_token_pooler(fused, p, "text_pooler"), p, "classifier")


def fcmf_logits(sequence_output: Tensor, visual_embeds_att: Tensor, roi_embeds_att: Tensor, roi_coors: Tensor,
                added_attention_mask: Tensor, p: Params, heads: int, num_imgs: int, num_roi: int) -> Tensor:
    """FCMF.forward with the text encoder stubbed (fcmf_multimodal.py:39-51)."""
    fused = fusion_encoder(sequence_output, visual_embeds_att, roi_embeds_att, roi_coors,
                           added_attention_mask, p, heads, num_imgs, num_roi)
    return classifier_head(fused, p)


def aspect_loop(seq_all: Tensor, visual_embeds_att: Tensor, roi_embeds_att: Tensor, roi_coors: Tensor,
                added_mask_all: Tensor, labels: Tensor, p: Params, heads: int, num_imgs: int, num_roi: int
                ) -> Tuple[Tensor, Tensor]:
    """The training-step body (run_multimodal_fcmf.py:462-475): for each aspect call the model on the
    [:, a] slices with the SAME visual tensors, loss = sum over aspects of the batch-mean cross entropy.

    seq_all [B,A,L,H], added_mask_all [B,A,Lm], labels [B,A]. Returns (logits [B,A,C], loss)."""
    total = None
    outs = []
    for a in range(seq_all.shape[1]):
        logits = fcmf_logits(seq_all[:, a], visual_embeds_att, roi_embeds_att, roi_coors,
                             added_mask_all[:, a], p, heads, num_imgs, num_roi)
        loss = torch.nn.functional.cross_entropy(logits, labels[:, a])
        total = loss if total is None else total + loss
        outs.append(logits)
    return torch.stack(outs, dim=1), total


# ------------------------------------------------------------------ IAOG decoder
def head_attention(k_in: Tensor, q_in: Tensor, w_kx: Tensor, w_qx: Tensor, proj_w: Tensor, proj_b: Tensor,
                   causal: bool, valid_len: Tensor | None = None) -> Tensor:
    """Attention.forward, 'scaled_dot_product' (mm_modeling.py:66-132): per-head weight tensors
    [nh, H, dh]; the projected KEYS are also the values (line 129); a 2-D ``memory_len`` selects a
    tril(q_len x k_len) mask on self- AND cross-attention (lines 115-118), a 1-D one a length mask;
    masked_fill(-1e4) (line 124); heads are concatenated on the feature axis (line 130)."""
    nh = w_kx.shape[0]
    # Entry n of the reference's flattened batch pairs INPUT  k.repeat(nh,1,1)[n]      = batch  n % B   (head-major)
    #                                             with WEIGHT w_kx.repeat(B,1,1)[n]   = head   n % nh  (batch-major)
    # (mm_modeling.py:79-85), and torch.split/cat (line 130) puts entry n = c*B + b into output slot c of batch b.
    # So slot c of batch b is projected with head weights (c*B + b) % nh -- a quirk that is part of the contract.
    B = k_in.shape[0]
    slot_head = (torch.arange(nh).view(nh, 1) * B + torch.arange(B).view(1, B)) % nh          # [slot c, batch b]
    kx = torch.einsum("bke,cbed->cbkd", k_in, w_kx[slot_head])
    qx = torch.einsum("bqe,cbed->cbqd", q_in, w_qx[slot_head])
    score = torch.matmul(qx, kx.transpose(-1, -2)) / math.sqrt(w_kx.shape[-1])
    q_len, k_len = score.shape[-2], score.shape[-1]
    if causal:
        keep = torch.tril(torch.ones(q_len, k_len, dtype=torch.bool, device=score.device))
        score = score.masked_fill(~keep, -1e4)
    elif valid_len is not None:
        keep = torch.arange(k_len, device=score.device).unsqueeze(0) < valid_len.unsqueeze(1)   # [B,k]
        score = score.masked_fill(~keep.view(1, -1, 1, k_len), -1e4)
    prob = torch.softmax(score, dim=-1)
    out = torch.matmul(prob, kx)                                   # [nh,B,q,dh]
    out = out.permute(1, 2, 0, 3).reshape(q_in.shape[0], q_len, nh * w_kx.shape[-1])
    return torch.nn.functional.linear(out, proj_w, proj_b)


def sinusoid_table(max_pos: int, hidden: int) -> Tensor:
    """PositionalEncoding.__init__ (mm_modeling.py:619-626)."""
    pe = torch.zeros(max_pos, hidden)
    x = torch.arange(max_pos, dtype=torch.float32).reshape(-1, 1) / torch.pow(
        10000, torch.arange(0, hidden, 2, dtype=torch.float32) / hidden)
    pe[:, 0::2] = torch.sin(x)
    pe[:, 1::2] = torch.cos(x)
    return pe


def iaog_decoder(dec_x: Tensor, enc_out: Tensor, p: Params, num_blocks: int, pre: str = "decoder.") -> Tensor:
    """IAOGDecoder.forward in training mode (mm_modeling.py:649-662) over TransformerDecoderBlock.forward
    (mm_modeling.py:585-613). Both attentions receive a 2-D mask argument in training
    (dec_valid_lens [B,T] at line 601; combined_mask [B,15] at lines 607-610) so both are tril-masked."""
    emb = p[pre + "embedding.weight"]
    hid = emb.shape[1]
    x = emb[dec_x] * math.sqrt(hid) + sinusoid_table(512, hid)[: dec_x.shape[1]].to(emb.dtype)
    for i in range(num_blocks):
        b = f"{pre}blks.block{i}."
        x2 = head_attention(x, x, p[b + "attention1.w_kx"], p[b + "attention1.w_qx"],
                            p[b + "attention1.proj.weight"], p[b + "attention1.proj.bias"], causal=True)
        y = tf_layer_norm(x2 + x, p[b + "addnorm1.ln.weight"], p[b + "addnorm1.ln.bias"])
        y2 = head_attention(enc_out, y, p[b + "attention2.w_kx"], p[b + "attention2.w_qx"],
                            p[b + "attention2.proj.weight"], p[b + "attention2.proj.bias"], causal=True)
        z = tf_layer_norm(y2 + y, p[b + "addnorm2.ln.weight"], p[b + "addnorm2.ln.bias"])
        f = affine(erf_gelu(affine(z, p, b + "ffn.dense1")), p, b + "ffn.dense2")
        x = tf_layer_norm(f + z, p[b + "add_norm3.ln.weight"], p[b + "add_norm3.ln.bias"])
    return torch.nn.functional.linear(x, p[pre + "dense.weight"], p[pre + "dense.bias"])
