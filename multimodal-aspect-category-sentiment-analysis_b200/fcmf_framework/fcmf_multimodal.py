"""FCMF classifier model with the reference's signature and state_dict keys
(reference: fcmf_framework/fcmf_multimodal.py:12-51), plus the folded all-aspects entry point that replaces the
Python loop of run_multimodal_fcmf.py:462-475 by one launch sequence."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import mm_modeling as M
from .mm_modeling import BertPooler
from .fcmf_pretraining import FCMFEncoder
from .. import functional as Fn


class FCMF(nn.Module):
    def __init__(self, pretrained_path, num_labels=4, num_imgs=7, num_roi=7, alpha=0.7):
        super().__init__()
        self.encoder = FCMFEncoder(pretrained_path, num_imgs, num_roi, alpha)
        self.text_pooler = BertPooler()
        self.dropout_p = M.HIDDEN_DROPOUT_PROB
        self.classifier = nn.Linear(M.HIDDEN_SIZE, num_labels)

    # ---- head: text_pooler -> (dropout) -> classifier [-> cross entropy] on the folded rows -------------------
    def head(self, fused: torch.Tensor, labels: Optional[torch.Tensor] = None, row_scale: float = 1.0,
             step_seed: Optional[int] = None):
        """fused [R, F, H] -> (logits [R, C] fp32, loss = row_scale * sum_r CE) ; labels None -> loss is 0.
        train() mode: dropout(p=0.1) on the pooled vector inside the classifier kernel (fcmf_multimodal.py:49)."""
        from ..fusion import DROP_SITES
        pooled = self.text_pooler(fused)
        drop = None
        if self.training:
            drop = Fn.site_drop(Fn.new_step_seed() if step_seed is None else step_seed, DROP_SITES["head"], self.dropout_p)
        return Fn.classifier_ce(pooled, self.classifier.weight, self.classifier.bias, labels, row_scale, drop)

    def forward(self, input_ids, visual_embeds_att, roi_embeds_att, roi_coors=None, token_type_ids=None,
                attention_mask=None, added_attention_mask=None):
        sequence_output, _pooled, _att = self.encoder.bert(input_ids, token_type_ids, attention_mask)
        fused = self.encoder.fuse(sequence_output, visual_embeds_att, roi_embeds_att, roi_coors, added_attention_mask)
        logits, _ = self.head(fused)
        return logits

    def forward_all_aspects(self, input_ids, visual_embeds_att, roi_embeds_att, roi_coors, token_type_ids,
                            attention_mask, added_attention_mask, labels=None, rows: Optional[str] = None):
        """id/mask tensors [B, A, L] / [B, A, Lm], labels [B, A] -> (logits [B, A, C], loss).
        loss = sum_a mean_b CE(logits[:, a], labels[:, a]) exactly as run_multimodal_fcmf.py:474-475 accumulates it."""
        B, A, _ = input_ids.shape
        flat = lambda t: None if t is None else t.reshape(B * A, -1)
        sequence_output, _pooled, _att = self.encoder.bert(flat(input_ids), flat(token_type_ids), flat(attention_mask))
        return self.fuse_all_aspects(sequence_output, visual_embeds_att, roi_embeds_att, roi_coors,
                                     added_attention_mask, labels, aspects=A, rows=rows)

    def fuse_all_aspects(self, sequence_output, visual_embeds_att, roi_embeds_att, roi_coors, added_attention_mask,
                         labels=None, aspects: int = 1, rows: Optional[str] = None, step_seed: Optional[int] = None):
        """The fusion hot path given text-encoder states [B*A, L, H] (or [B, A, L, H]). In train() mode every dropout
        of the reference path is applied inside the kernels; step_seed (default: drawn from torch's CPU generator)
        determines all masks of the step."""
        if sequence_output.dim() == 4:
            aspects = sequence_output.shape[1]
            sequence_output = sequence_output.reshape(-1, *sequence_output.shape[2:])
        BA = sequence_output.shape[0]
        B = BA // aspects
        if self.training and step_seed is None:
            step_seed = Fn.new_step_seed()
        fused = self.encoder.fuse(sequence_output, visual_embeds_att, roi_embeds_att, roi_coors,
                                  added_attention_mask.reshape(BA, -1), aspects=aspects, rows=rows, step_seed=step_seed)
        lab = None if labels is None else labels.reshape(BA)
        logits, loss = self.head(fused, lab, row_scale=1.0 / B, step_seed=step_seed)
        return logits.view(B, aspects, -1), loss
