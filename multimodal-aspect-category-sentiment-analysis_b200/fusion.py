"""Folded, hoisted orchestration of the FCMF fusion encoder (FCMFEncoder.forward after the text encoder,
fcmf_pretraining.py:42-141) on top of the kernel Functions in functional.py.

What changes relative to the reference's execution order (results are identical, tests/ prove it):
  * the aspect loop (run_multimodal_fcmf.py:464-475) and the image loop (fcmf_pretraining.py:47) are folded
    into the launch dimension: problem p = (b*A + a)*NI + i;
  * aspect-independent work runs once per SAMPLE: vismap2text, roimap2text, box_head, the K/V projections of the
    image patches and the Q/K/V projections of the ROI rows under mm_attention;
  * image-independent work runs once per (sample, aspect): the text->image query projection and the Q/K/V
    projections of the text rows under mm_attention;
  * torch.cat((sequence_output, relative_roi)) (line 114) is never materialised for the attention: the kernel
    reads two row segments;
  * rows="live" additionally drops the rows no output depends on (every per-image branch ends in BertPooler,
    which keeps token 0 -- mm_modeling.py:425-431): only query row 0 goes through attention/LN/FFN.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import functional as Fn
from . import ops
from ._lib import ENGINE_AUTO, mark

Tensor = torch.Tensor
PATCH_MASK = 49                    # fcmf_pretraining.py:53

# Dropout sites of the folded path (train() mode). One site = one nn.Dropout module call of the reference, folded over
# aspects and images: the mask of a site is keep(site seed, row, column) with the row convention noted per site.
#   *_attn : attention probabilities (mm_modeling.py:213, 260; roi_modeling.py:42-43)   row = (problem*heads + head)*Lq + query
#   *_out1 : BertSelfOutput hidden dropout (mm_modeling.py:278)                          row = row of the layer-tail input
#   *_out2 : BertOutput hidden dropout (mm_modeling.py:326)                              row = row of the layer-tail input
#   head   : FCMF.dropout on the pooled [CLS] (fcmf_multimodal.py:49)                    row = b*A + a
# box_attn is drawn once per (sample, image) because the ROI self-attention is hoisted out of the aspect loop (the
# reference redraws it for each of the A aspect passes); every other site has one independent draw per reference call.
DROP_SITES = {"box_attn": 1, "t2i_attn": 2, "t2i_out1": 3, "t2i_out2": 4, "mm_attn": 5, "mm_out1": 6, "mm_out2": 7,
              "fus_attn": 8, "fus_out1": 9, "fus_out2": 10, "head": 11}


class _Index:
    """int32 device index tables of one (B, A, L, NI, NR, rows) configuration."""

    def __init__(self, B: int, A: int, L: int, NI: int, NR: int, live: bool, device):
        i32 = dict(dtype=torch.int32, device=device)
        BA, NP = B * A, B * A * NI
        p = torch.arange(NP, **i32)
        self.p2ba = (p // NI).contiguous()
        self.p2bi = ((p // (A * NI)) * NI + p % NI).contiguous()
        ba = torch.arange(BA, **i32).view(BA, 1)
        self.ba2p = (ba * NI + torch.arange(NI, **i32).view(1, NI)).contiguous()                     # [BA, NI]
        b = torch.arange(B, **i32).view(B, 1, 1)
        i = torch.arange(NI, **i32).view(1, NI, 1)
        a = torch.arange(A, **i32).view(1, 1, A)
        self.bi2p = (((b * A + a) * NI + i).reshape(B * NI, A)).contiguous()                       # [B*NI, A]

        Lq = 1 if live else L
        # text->image tail: row m = p*Lq + l takes residual row ba*L + l
        l = torch.arange(Lq, **i32).view(1, Lq)
        self.t2i_res_idx = (self.p2ba.view(NP, 1) * L + l).reshape(-1).contiguous()
        inv = torch.full((BA, L, NI), -1, **i32)
        inv[:, :Lq, :] = (self.ba2p.view(BA, 1, NI) * Lq + l.view(1, Lq, 1))
        self.t2i_res_inv = inv.reshape(BA * L, NI).contiguous()

        # text+ROI tail
        if live:
            self.roi_res_idx = self.t2i_res_idx
            self.roi_res_inv = self.t2i_res_inv
        else:
            S = L + NR
            s = torch.arange(S, **i32).view(1, S)
            text_row = self.p2ba.view(NP, 1) * L + s
            roi_row = BA * L + self.p2bi.view(NP, 1) * NR + (s - L)
            self.roi_res_idx = torch.where(s < L, text_row, roi_row).reshape(-1).contiguous()
            G = max(NI, A)
            inv_t = torch.full((BA, L, G), -1, **i32)
            inv_t[:, :, :NI] = self.ba2p.view(BA, 1, NI) * S + torch.arange(L, **i32).view(1, L, 1)
            inv_r = torch.full((B * NI, NR, G), -1, **i32)
            inv_r[:, :, :A] = self.bi2p.view(B * NI, 1, A) * S + L + torch.arange(NR, **i32).view(1, NR, 1)
            self.roi_res_inv = torch.cat((inv_t.reshape(BA * L, G), inv_r.reshape(B * NI * NR, G)), 0).contiguous()


_INDEX_CACHE: Dict[Tuple, _Index] = {}


def _index(B, A, L, NI, NR, live, device) -> _Index:
    key = (B, A, L, NI, NR, live, str(device))
    if key not in _INDEX_CACHE:
        if len(_INDEX_CACHE) > 64:
            _INDEX_CACHE.clear()
        _INDEX_CACHE[key] = _Index(B, A, L, NI, NR, live, device)
    return _INDEX_CACHE[key]


def _cat_wb(mods):
    return torch.cat([m.weight for m in mods], 0), torch.cat([m.bias for m in mods], 0)


def _tail_params(layer):
    """(Wo, bo, ln1.w, ln1.b, W1, b1, W2, b2, ln2.w, ln2.b) of a BertLayer / BertCrossAttentionLayer."""
    ao, it, out = layer.attention.output, layer.intermediate, layer.output
    return (ao.dense.weight, ao.dense.bias, ao.LayerNorm.weight, ao.LayerNorm.bias, it.dense.weight, it.dense.bias,
            out.dense.weight, out.dense.bias, out.LayerNorm.weight, out.LayerNorm.bias)


def fused_forward(enc, sequence_output: Tensor, visual_embeds_att: Tensor, roi_embeds_att: Tensor, roi_coors: Tensor,
                  added_attention_mask: Tensor, aspects: int = 1, rows: str = "full", engine: int = ENGINE_AUTO,
                  compute_dtype: Optional[torch.dtype] = None, step_seed: Optional[int] = None) -> Tensor:
    """sequence_output [B*A, L, H] (row ba = b*A + a), visual tensors [B, ...], added_attention_mask [B*A, Lm].
    Returns the fused sequence [B*A, 1 + 2*NI, H] of fcmf_pretraining.py:139-141 in the compute dtype.
    In train() mode the reference's dropouts are applied inside the kernels (DROP_SITES); step_seed fixes the masks
    (default: a fresh draw from torch's CPU generator)."""
    if rows not in ("full", "live"):
        raise ValueError(f"rows must be 'full' or 'live', got {rows!r}")
    live = rows == "live"
    BA, L, H = sequence_output.shape
    A = aspects
    if BA % A != 0:
        raise RuntimeError(f"folded batch {BA} is not a multiple of the aspect count {A}")
    B = BA // A
    NI, NR = enc.num_imgs, enc.num_roi
    if visual_embeds_att.shape[0] != B or visual_embeds_att.shape[1] < NI:
        raise RuntimeError(f"visual_embeds_att {tuple(visual_embeds_att.shape)} does not provide {NI} images for {B} samples")
    if roi_embeds_att.shape[0] != B or roi_embeds_att.shape[1] < NI or roi_embeds_att.shape[2] != NR:
        raise RuntimeError(f"roi_embeds_att {tuple(roi_embeds_att.shape)} does not match num_imgs={NI}, num_roi={NR}")
    P, Dv = visual_embeds_att.shape[2], visual_embeds_att.shape[3]
    if P != PATCH_MASK:
        raise RuntimeError(f"the reference masks exactly {PATCH_MASK} patches (fcmf_pretraining.py:53); got {P}")
    S, F = L + NR, 1 + 2 * NI
    dt = compute_dtype or sequence_output.dtype
    dev = sequence_output.device
    nh = enc.text2img_attention.layer[0].attention.self.num_attention_heads
    dh = H // nh
    ix = _index(B, A, L, NI, NR, live, dev)
    mask_add = ops.mask_additive(added_attention_mask.reshape(BA, -1), max(P, S, F))      # [BA, Lmm] fp32

    t2i = enc.text2img_attention.layer[0]
    mm = enc.mm_attention.layer[0]
    if enc.training and step_seed is None:
        step_seed = Fn.new_step_seed()
    if not enc.training:
        step_seed = None

    def drop(site: str, module) -> Optional[ops.Drop]:
        return Fn.site_drop(step_seed, DROP_SITES[site], float(getattr(module, "dropout_p", 0.0)))

    seq2 = sequence_output.to(dt).reshape(BA * L, H)
    vis2 = visual_embeds_att[:, :NI].to(dt).reshape(B * NI * P, Dv)
    roi2 = roi_embeds_att[:, :NI].to(dt).reshape(B * NI * NR, Dv)

    # ---- once per sample: what the text->image branch needs ---------------------------------------------
    # Issue order = reverse of the order in which backward finishes the gradients (autograd runs nodes in reverse creation order):
    # every hoisted projection, and every torch.cat of weights, is created right before the branch that consumes it, so that its
    # weight gradients are final when that branch's backward ends and their all-reduce bucket can start under the remaining backward
    # (at N = 8 the mm_attention and ROI-side buckets used to become final after the last kernel of the step).
    mark("fusion: per-sample image projections")
    patches = Fn.linear(vis2, enc.vismap2text.weight, enc.vismap2text.bias, engine=engine)            # [B*NI*P, H]
    w_kv, b_kv = _cat_wb((t2i.attention.self.key, t2i.attention.self.value))
    kv_p = Fn.linear(patches, w_kv, b_kv, engine=engine)                                                # [B*NI*P, 2H]

    # ---- once per (sample, aspect) ----------------------------------------------------------------------
    mark("fusion: per-(sample, aspect) projections")
    tq = t2i.attention.self.query
    if live:
        cls_rows = sequence_output.to(dt)[:, 0, :]                                                      # [BA, H] strided
        q_t = Fn.linear(cls_rows, tq.weight, tq.bias, engine=engine)                                    # [BA, H]
    else:
        q_t = Fn.linear(seq2, tq.weight, tq.bias, engine=engine)                                        # [BA*L, H]
    Lq = 1 if live else L
    NP = BA * NI

    # ---- text -> image branch, all (sample, aspect, image) problems in one launch ------------------------
    mark("fusion: text->image branch")
    plan1 = Fn.AttnPlan(NP, nh, dh, mask_div=NI, drop=drop("t2i_attn", t2i.attention.self)).add("q", 0, 0, Lq, ix.p2ba, ix.ba2p) \
        .add("k", 1, 0, P, ix.p2bi, ix.bi2p).add("v", 1, H, P, ix.p2bi, ix.bi2p)
    ctx1 = Fn.folded_attention(plan1, (q_t, kv_p), mask_add, None)                                      # [NP*Lq, H]
    # every row is computed (rows="full"), but BertPooler reads token 0 only: the tail hands out that row of each problem
    y1 = Fn.layer_tail(ctx1, seq2, ix.t2i_res_idx, ix.t2i_res_inv, _tail_params(t2i), engine=engine,
                       drop1=drop("t2i_out1", t2i.attention.output), drop2=drop("t2i_out2", t2i.output), out_every=Lq)
    h_img = Fn.linear(y1, enc.text2img_pooler.dense.weight, enc.text2img_pooler.dense.bias,
                      act="tanh", engine=engine)                                                        # [NP, H]

    # ---- once per sample: ROI side (box attention -> relative_roi), then once per (sample, aspect): Q/K/V of the text rows --------
    mark("fusion: per-sample ROI projections + box attention")
    roi_p = Fn.linear(roi2, enc.roimap2text.weight, enc.roimap2text.bias, engine=engine)              # [B*NI*NR, H]
    box = enc.box_head
    w_b, b_b = _cat_wb(box.linears[:3])
    qkv_b = Fn.linear(roi_p, w_b, b_b, engine=engine)                                                   # [B*NI*NR, 3H]
    wg_w = torch.cat([g.weight for g in box.WGs], 0)                                                    # [8, 64]
    wg_b = torch.cat([g.bias for g in box.WGs], 0)                                                      # [8]
    geo = Fn.box_geometry(roi_coors[:, :NI].reshape(B * NI, NR, 4), wg_w, wg_b)                         # [B*NI, 8, NR, NR]
    dkb = H // box.h
    plan_b = Fn.AttnPlan(B * NI, box.h, dkb, drop=drop("box_attn", box)).add("q", 0, 0, NR, None, None).add("k", 0, H, NR, None, None) \
        .add("v", 0, 2 * H, NR, None, None)
    ctx_b = Fn.folded_attention(plan_b, (qkv_b,), None, geo)                                            # [B*NI*NR, H]
    rel = Fn.linear(ctx_b, box.linears[3].weight, box.linears[3].bias, engine=engine)                   # relative_roi
    w_mm, b_mm = _cat_wb((mm.attention.self.query, mm.attention.self.key, mm.attention.self.value))

    # ---- text + ROI branch ---------------------------------------------------------------------------------
    mark("fusion: text+ROI branch")
    if live:
        kv_t = Fn.linear(seq2, w_mm[H:], b_mm[H:], engine=engine)                                       # [BA*L, 2H]
        q0_mm = Fn.linear(cls_rows, w_mm[:H], b_mm[:H], engine=engine)                                  # [BA, H]
        kv_r = Fn.linear(rel, w_mm[H:], b_mm[H:], engine=engine)                                        # [B*NI*NR, 2H]
    else:
        qkv_t = Fn.linear(seq2, w_mm, b_mm, engine=engine)                                              # [BA*L, 3H]
        qkv_r = Fn.linear(rel, w_mm, b_mm, engine=engine)                                               # [B*NI*NR, 3H]
    if live:
        plan2 = Fn.AttnPlan(NP, nh, dh, mask_div=NI, drop=drop("mm_attn", mm.attention.self)).add("q", 0, 0, 1, ix.p2ba, ix.ba2p) \
            .add("k", 1, 0, L, ix.p2ba, ix.ba2p).add("k", 2, 0, NR, ix.p2bi, ix.bi2p) \
            .add("v", 1, H, L, ix.p2ba, ix.ba2p).add("v", 2, H, NR, ix.p2bi, ix.bi2p)
        ctx2 = Fn.folded_attention(plan2, (q0_mm, kv_t, kv_r), mask_add, None)                          # [NP, H]
        y2 = Fn.layer_tail(ctx2, seq2, ix.roi_res_idx, ix.roi_res_inv, _tail_params(mm), engine=engine,
                           drop1=drop("mm_out1", mm.attention.output), drop2=drop("mm_out2", mm.output))
    else:
        plan2 = Fn.AttnPlan(NP, nh, dh, mask_div=NI, drop=drop("mm_attn", mm.attention.self))
        for role, col in (("q", 0), ("k", H), ("v", 2 * H)):
            plan2.add(role, 0, col, L, ix.p2ba, ix.ba2p).add(role, 1, col, NR, ix.p2bi, ix.bi2p)
        ctx2 = Fn.folded_attention(plan2, (qkv_t, qkv_r), mask_add, None)                               # [NP*S, H]
        text_roi = torch.cat((seq2, rel), 0)                                                            # residual rows
        y2 = Fn.layer_tail(ctx2, text_roi, ix.roi_res_idx, ix.roi_res_inv, _tail_params(mm), engine=engine,
                           drop1=drop("mm_out1", mm.attention.output), drop2=drop("mm_out2", mm.output), out_every=S)
    r_img = Fn.linear(y2, enc.text2roi_pooler.dense.weight, enc.text2roi_pooler.dense.bias,
                      act="tanh", engine=engine)                                                        # [NP, H]

    # ---- fusion layer: [CLS] + h_1..NI + r_1..NI (fcmf_pretraining.py:127-139), same mm_attention weights ----
    mark("fusion: 15-token fusion layer")
    fusion = torch.cat((sequence_output.to(dt)[:, 0:1, :], h_img.view(BA, NI, H), r_img.view(BA, NI, H)), 1)
    x = fusion.reshape(BA * F, H)
    qkv_f = Fn.linear(x, w_mm, b_mm, engine=engine)
    plan3 = Fn.AttnPlan(BA, nh, dh, mask_div=1, drop=drop("fus_attn", mm.attention.self)).add("q", 0, 0, F, None, None).add("k", 0, H, F, None, None) \
        .add("v", 0, 2 * H, F, None, None)
    ctx3 = Fn.folded_attention(plan3, (qkv_f,), mask_add, None)
    out = Fn.layer_tail(ctx3, x, None, None, _tail_params(mm), engine=engine,
                        drop1=drop("fus_out1", mm.attention.output), drop2=drop("fus_out2", mm.output))
    return out.view(BA, F, H)
