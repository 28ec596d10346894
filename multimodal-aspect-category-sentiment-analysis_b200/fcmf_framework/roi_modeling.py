"""Geometric ROI attention with the reference's names (reference: fcmf_framework/roi_modeling.py).

``BoxMultiHeadedAttention`` keeps ``linears.{0-3}`` and ``WGs.{0-7}`` so checkpoints load; the pairwise box
embedding, the per-head geometry weights and the 4x4 (NR x NR) attention all run in kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as Fn


class BoxMultiHeadedAttention(nn.Module):
    def __init__(self, h, d_model, trignometric_embedding=True, legacy_extra_skip=False, dropout=0.1):
        super().__init__()
        if d_model % h != 0:
            raise AssertionError("d_model must be a multiple of the head count")
        if not trignometric_embedding:
            raise NotImplementedError("only the trigonometric embedding (the configuration FCMF uses) is built")
        self.trignometric_embedding = trignometric_embedding
        self.legacy_extra_skip = legacy_extra_skip
        self.h = h
        self.d_k = d_model // h
        self.dim_g = 64
        self.linears = nn.ModuleList([nn.Linear(d_model, d_model) for _ in range(4)])
        self.WGs = nn.ModuleList([nn.Linear(self.dim_g, 1, bias=True) for _ in range(8)])
        if len(self.WGs) != h:
            raise AssertionError("the reference builds exactly 8 geometry heads (roi_modeling.py:74)")
        self.attn = None
        self.box_attn = None
        self.dropout_p = dropout

    def forward(self, input_query, input_key, input_value, input_box, mask=None):
        """q/k/v [G, NR, d_model], input_box [G, NR, 4] as (x_min, x_max, y_min, y_max) -> [G, NR, d_model]."""
        if mask is not None:
            raise NotImplementedError("FCMF never masks the ROI box attention (fcmf_pretraining.py:106-111)")
        G, NR, H = input_query.shape
        flat = lambda t: t.reshape(G * NR, H)
        if input_key is input_query and input_value is input_query:
            w = torch.cat([l.weight for l in self.linears[:3]], 0)
            b = torch.cat([l.bias for l in self.linears[:3]], 0)
            qkv = Fn.linear(flat(input_query), w, b)
        else:
            qkv = torch.cat([Fn.linear(flat(x), l.weight, l.bias)
                             for l, x in zip(self.linears[:3], (input_query, input_key, input_value))], 1)
        wg_w = torch.cat([g.weight for g in self.WGs], 0)
        wg_b = torch.cat([g.bias for g in self.WGs], 0)
        geo = Fn.box_geometry(input_box.reshape(G, NR, 4), wg_w, wg_b)
        plan = Fn.AttnPlan(G, self.h, self.d_k, drop=Fn.fresh_drop(self.dropout_p, self.training)).add("q", 0, 0, NR, None, None).add("k", 0, H, NR, None, None) \
            .add("v", 0, 2 * H, NR, None, None)
        x = Fn.folded_attention(plan, (qkv,), None, geo)
        if self.legacy_extra_skip:
            x = flat(input_value) + x
        return Fn.linear(x, self.linears[3].weight, self.linears[3].bias).view(G, NR, H)
