"""Deterministic synthetic weights and batches for the FCMF fusion path.

Everything is drawn from ``numpy.random.RandomState`` (a frozen bit stream) so
that the golden fixtures under ``tests/golden`` can be regenerated on any box
from a seed instead of committing 86 MB of weights.

Shapes follow SURVEY.md section 8(d): the reference dataset
(vimacsa_dataset.py:101-106,169-199) and training loop
(run_multimodal_fcmf.py:439-473) define them.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

import numpy as np
import torch

PATCHES = 49          # fcmf_pretraining.py:53 hard-codes the 49-slice of the mask
VIS_DIM = 2048        # fcmf_pretraining.py:25-26  nn.Linear(2048, HIDDEN_SIZE)
GEO_DIM = 64          # roi_modeling.py:66
BOX_HEADS = 8         # fcmf_pretraining.py:29  BoxMultiHeadedAttention(8, HIDDEN_SIZE)


@dataclass(frozen=True)
class FusionDims:
    """Static dimensions of one fusion problem."""
    hidden: int = 768
    heads: int = 12
    inter: int = 3072
    num_imgs: int = 7
    num_roi: int = 4
    num_labels: int = 4
    seq_len: int = 170
    aspects: int = 6
    batch: int = 4

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def fused_len(self) -> int:          # 1 + 2*NI, fcmf_pretraining.py:131
        return 1 + 2 * self.num_imgs

    @property
    def mask_len(self) -> int:
        return max(self.seq_len + PATCHES, self.seq_len + self.num_roi, self.fused_len)

    def to_dict(self):
        return asdict(self)


def _layer_keys(prefix: str, hidden: int, inter: int):
    p = prefix + ".layer.0."
    out = []
    for n in ("query", "key", "value"):
        out += [(p + f"attention.self.{n}.weight", (hidden, hidden)), (p + f"attention.self.{n}.bias", (hidden,))]
    out += [(p + "attention.output.dense.weight", (hidden, hidden)), (p + "attention.output.dense.bias", (hidden,)),
            (p + "attention.output.LayerNorm.weight", (hidden,)), (p + "attention.output.LayerNorm.bias", (hidden,)),
            (p + "intermediate.dense.weight", (inter, hidden)), (p + "intermediate.dense.bias", (inter,)),
            (p + "output.dense.weight", (hidden, inter)), (p + "output.dense.bias", (hidden,)),
            (p + "output.LayerNorm.weight", (hidden,)), (p + "output.LayerNorm.bias", (hidden,))]
    return out


def fusion_param_spec(dims: FusionDims, with_head: bool = True):
    """(key, shape) for every fusion-path parameter, in the reference's state_dict naming
    (probed from FCMF(...).state_dict())."""
    H, I = dims.hidden, dims.inter
    spec = [("encoder.vismap2text.weight", (H, VIS_DIM)), ("encoder.vismap2text.bias", (H,)),
            ("encoder.roimap2text.weight", (H, VIS_DIM)), ("encoder.roimap2text.bias", (H,))]
    for i in range(4):
        spec += [(f"encoder.box_head.linears.{i}.weight", (H, H)), (f"encoder.box_head.linears.{i}.bias", (H,))]
    for i in range(BOX_HEADS):
        spec += [(f"encoder.box_head.WGs.{i}.weight", (1, GEO_DIM)), (f"encoder.box_head.WGs.{i}.bias", (1,))]
    spec += _layer_keys("encoder.text2img_attention", H, I)
    spec += [("encoder.text2img_pooler.dense.weight", (H, H)), ("encoder.text2img_pooler.dense.bias", (H,)),
             ("encoder.text2roi_pooler.dense.weight", (H, H)), ("encoder.text2roi_pooler.dense.bias", (H,))]
    spec += _layer_keys("encoder.mm_attention", H, I)
    if with_head:
        spec += [("text_pooler.dense.weight", (H, H)), ("text_pooler.dense.bias", (H,)),
                 ("classifier.weight", (dims.num_labels, H)), ("classifier.bias", (dims.num_labels,))]
    return spec


def make_params(dims: FusionDims, seed: int = 42, with_head: bool = True, w_std: float = 0.05):
    """fp32 CPU tensors keyed like the reference state_dict. Scales are chosen so that the
    attention softmaxes are far from uniform and the ReLU/clamp in the geometry path is
    exercised on both sides (so a wrong mask or a wrong clamp shows up in parity)."""
    rs = np.random.RandomState(seed)
    out = {}
    for key, shape in fusion_param_spec(dims, with_head):
        n = int(np.prod(shape))
        z = rs.standard_normal(n).astype(np.float32).reshape(shape)
        if "LayerNorm.weight" in key:
            v = 1.0 + 0.1 * z
        elif "LayerNorm.bias" in key:
            v = 0.1 * z
        elif ".WGs." in key and key.endswith("weight"):
            v = 0.3 * z
        elif ".WGs." in key and key.endswith("bias"):
            v = 0.5 + 0.3 * z
        elif key.endswith("bias"):
            v = 0.02 * z
        else:
            v = w_std * z
        out[key] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return out


def make_batch(dims: FusionDims, seed: int = 1234, mask: str = "ones"):
    """One synthetic training batch of the fusion path (text encoder stubbed by a leaf
    ``sequence_output``), SURVEY.md section 8(d) table.

    mask = "ones"      -> all-ones added_attention_mask (vimacsa_dataset.py:106)
    mask = "bernoulli" -> Bernoulli(0.8) with index 0 forced to 1 (parity tests only)
    """
    rs = np.random.RandomState(seed)
    B, A, L, H = dims.batch, dims.aspects, dims.seq_len, dims.hidden
    NI, NR = dims.num_imgs, dims.num_roi
    seq = rs.standard_normal(B * A * L * H).astype(np.float32).reshape(B, A, L, H)
    vis = rs.random_sample(B * NI * PATCHES * VIS_DIM).astype(np.float32).reshape(B, NI, PATCHES, VIS_DIM)
    roi = rs.random_sample(B * NI * NR * VIS_DIM).astype(np.float32).reshape(B, NI, NR, VIS_DIM)
    xy = rs.random_sample(B * NI * NR * 4).reshape(B, NI, NR, 2, 2)
    xy.sort(axis=-1)                                   # (x_min,x_max),(y_min,y_max)  roi_modeling.py:95
    coors = xy.reshape(B, NI, NR, 4).astype(np.float64)
    pad = rs.random_sample((B, NI, NR)) < 0.25         # padded ROIs are all-zero boxes (vimacsa_dataset.py:169-172)
    coors[pad] = 0.0
    if mask == "ones":
        m = np.ones((B, A, dims.mask_len), dtype=np.int64)
    else:
        m = (rs.random_sample((B, A, dims.mask_len)) < 0.8).astype(np.int64)
        m[..., 0] = 1
    labels = rs.randint(0, dims.num_labels, size=(B, A)).astype(np.int64)
    return {
        "sequence_output": torch.from_numpy(seq),
        "visual_embeds_att": torch.from_numpy(vis),
        "roi_embeds_att": torch.from_numpy(roi),
        "roi_coors": torch.from_numpy(coors),
        "added_attention_mask": torch.from_numpy(m),
        "labels": torch.from_numpy(labels),
    }


# --------------------------------------------------------------------------------------------
# FLOP accounting: multiply-add = 2, backward = 2x forward.

def flops_forward_per_sample(dims: FusionDims, mode: str) -> float:
    """Forward FLOPs of the fusion path for ONE sample (= all aspects x all images).

    mode: "exec" what the reference executes, "full" all rows with aspect-/image-independent
    projections hoisted, "live" only rows that reach an output."""
    H, I, L, P = dims.hidden, dims.inter, dims.seq_len, PATCHES
    NI, NR, A, Dv = dims.num_imgs, dims.num_roi, dims.aspects, VIS_DIM
    S, F = L + NR, dims.fused_len
    vismap = 2 * P * Dv * H
    t2i_q, t2i_kv = 2 * L * H * H, 4 * P * H * H
    t2i_att = 4 * L * P * H
    t2i_out, t2i_ffn = 2 * L * H * H, 4 * L * H * I
    roimap, box = 2 * NR * Dv * H, 8 * NR * H * H
    mm_qkv, mm_att, mm_out, mm_ffn = 6 * S * H * H, 4 * S * S * H, 2 * S * H * H, 4 * S * H * I
    pool = 2 * H * H
    final = 6 * F * H * H + 4 * F * F * H + 2 * F * H * H + 4 * F * H * I
    head = pool + 2 * H * dims.num_labels
    if mode == "exec":
        per_img = vismap + t2i_q + t2i_kv + t2i_att + t2i_out + t2i_ffn + roimap + box + mm_qkv + mm_att + mm_out + mm_ffn + 2 * pool
        return A * (NI * per_img + final + head)
    if mode == "full":
        per_sample = NI * (vismap + t2i_kv + roimap + box)
        per_sa = t2i_q + NI * (t2i_att + t2i_out + t2i_ffn + mm_qkv + mm_att + mm_out + mm_ffn + 2 * pool) + final + head
        return per_sample + A * per_sa
    if mode == "live":
        row = 2 * H * H                      # one row through an HxH projection
        ffn_row = 4 * H * I
        per_sample = NI * (vismap + t2i_kv + roimap + box + 4 * NR * H * H)      # + ROI K/V under mm_attention
        per_sa = (row + 4 * L * H * H + row                                      # t2i Q row 0; text K/V; mm Q row 0
                  + NI * (4 * P * H + row + ffn_row + 4 * S * H + row + ffn_row + 2 * pool)
                  + 4 * F * H * H + row + 4 * F * H + row + ffn_row + head)
        return per_sample + A * per_sa
    raise ValueError(mode)
