"""FCMFEncoder / FCMFSeq2Seq with the reference's constructor, forward signature and state_dict keys
(reference: fcmf_framework/fcmf_pretraining.py:14-221). The per-image Python loop is replaced by the folded,
hoisted kernel path in ``..fusion``."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import mm_modeling as M
from .mm_modeling import BertCrossEncoder, BertPooler, FeatureExtractor, MultimodalEncoder
from .roi_modeling import BoxMultiHeadedAttention
from .. import fusion


def _compute_dtype(module, like: torch.Tensor) -> torch.dtype:
    """bf16 under autocast (the reference's --fp16 AMP path maps to bf16 here), else the module's setting."""
    if torch.is_autocast_enabled():
        return torch.bfloat16
    return module.compute_dtype or (like.dtype if like.dtype in (torch.float32, torch.bfloat16) else torch.float32)


class FCMFEncoder(nn.Module):
    def __init__(self, pretrained_hf_path, num_imgs=7, num_roi=4, alpha=0.7):
        super().__init__()
        self.num_imgs = num_imgs
        self.num_roi = num_roi
        self.alpha = alpha                      # stored, unused by the live model (as in the reference)
        self.bert = FeatureExtractor(pretrained_hf_path) if pretrained_hf_path is not None else None
        self.vismap2text = nn.Linear(2048, M.HIDDEN_SIZE)
        self.roimap2text = nn.Linear(2048, M.HIDDEN_SIZE)
        self.box_head = BoxMultiHeadedAttention(8, M.HIDDEN_SIZE)
        self.text2img_attention = BertCrossEncoder()
        self.text2img_pooler = BertPooler()
        self.text2roi_pooler = BertPooler()
        self.mm_attention = MultimodalEncoder()
        # kernel-path knobs (not parameters, not in state_dict)
        self.compute_dtype: Optional[torch.dtype] = None     # None: follow the input / autocast
        self.rows = "full"                                    # "full" = every row the reference computes; "live" = rows that reach an output
        self.engine = 0                                       # GEMM engine: 0 auto, 1 CUDA-core, 2 tcgen05

    # ---- the fusion path given text-encoder states (what the kernels cover) ---------------------------------
    def fuse(self, sequence_output, visual_embeds_att, roi_embeds_att, roi_coors, added_attention_mask,
             aspects: int = 1, rows: Optional[str] = None, step_seed: Optional[int] = None):
        """sequence_output [B*aspects, L, H] -> [B*aspects, 1+2*num_imgs, H]; visual tensors are per SAMPLE.
        train() mode applies the reference's dropouts in the kernels; step_seed fixes the masks (tests)."""
        if roi_coors is None:
            raise ValueError("roi_coors is required (the reference dereferences it at fcmf_pretraining.py:110)")
        if added_attention_mask is None:
            raise ValueError("added_attention_mask is required (fcmf_pretraining.py:53)")
        dt = _compute_dtype(self, sequence_output)
        return fusion.fused_forward(self, sequence_output, visual_embeds_att, roi_embeds_att, roi_coors,
                                    added_attention_mask, aspects=aspects, rows=rows or self.rows,
                                    engine=self.engine, compute_dtype=dt, step_seed=step_seed)

    def forward(self, input_ids, visual_embeds_att, roi_embeds_att, roi_coors=None, token_type_ids=None,
                attention_mask=None, added_attention_mask=None):
        sequence_output, _pooled, enc_attentions = self.bert(input_ids, token_type_ids, attention_mask)
        fused = self.fuse(sequence_output, visual_embeds_att, roi_embeds_att, roi_coors, added_attention_mask)
        return fused.to(sequence_output.dtype), enc_attentions

    def forward_all_aspects(self, input_ids, visual_embeds_att, roi_embeds_att, roi_coors, token_type_ids,
                            attention_mask, added_attention_mask, rows: Optional[str] = None):
        """All aspect prompts of a sample in one launch: id/mask tensors are [B, A, L] / [B, A, Lm]
        (the loop of run_multimodal_fcmf.py:464-473 folded into the batch). Returns ([B, A, F, H], attentions)."""
        B, A, L = input_ids.shape
        flat = lambda t: None if t is None else t.reshape(B * A, -1)
        sequence_output, _pooled, enc_attentions = self.bert(flat(input_ids), flat(token_type_ids), flat(attention_mask))
        fused = self.fuse(sequence_output, visual_embeds_att, roi_embeds_att, roi_coors,
                          added_attention_mask.reshape(B * A, -1), aspects=A, rows=rows)
        return fused.view(B, A, fused.shape[1], fused.shape[2]), enc_attentions


class FCMFSeq2Seq(nn.Module):
    """Encoder + IAOG decoder wrapper (reference fcmf_pretraining.py:143-221). The encoder runs on the kernel path;
    the decoder is the next row of the scope table and is built in ``..iaog``."""

    def __init__(self, vocab_size, max_len_decoder, pretrained_hf_path, num_imgs, num_roi, alpha):
        super().__init__()
        from ..iaog import IAOGDecoder
        self.encoder = FCMFEncoder(pretrained_hf_path, num_imgs=num_imgs, num_roi=num_roi, alpha=alpha)
        self.decoder = IAOGDecoder(vocab_size=vocab_size)
        self.num_imgs = num_imgs
        for mod in (self.decoder, self.encoder.vismap2text, self.encoder.roimap2text, self.encoder.box_head,
                    self.encoder.text2img_attention, self.encoder.mm_attention):
            mod.apply(self._init_weights)
        cell = self.encoder.bert.cell if self.encoder.bert is not None else None
        if cell is not None and hasattr(cell, "resize_token_embeddings"):
            cell.resize_token_embeddings(vocab_size)
        if cell is not None and hasattr(cell, "embeddings"):
            self.decoder.embedding.weight = cell.embeddings.word_embeddings.weight
        self.decoder.dense.weight = self.decoder.embedding.weight

    def forward(self, enc_X, dec_X, visual_embeds_att, roi_embeds_att, roi_coors=None, token_type_ids=None,
                attention_mask=None, added_attention_mask=None, source_valid_len=None, is_train=True):
        enc_output, enc_attentions = self.encoder(enc_X, visual_embeds_att, roi_embeds_att, roi_coors, token_type_ids,
                                                  attention_mask, added_attention_mask)
        n_vis = self.num_imgs * 2
        text_len = enc_output.size(1) - n_vis
        text_mask = attention_mask[:, :text_len]
        combined_mask = torch.cat((text_mask, torch.ones((text_mask.size(0), n_vis), device=text_mask.device,
                                                         dtype=text_mask.dtype)), dim=1)
        state = [enc_output, combined_mask, [None] * self.decoder.num_blks]
        logits = self.decoder(dec_X, state, is_train=is_train)
        return logits if is_train else (logits, enc_attentions)

    @staticmethod
    def loss(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
        """The pre-training loss of run_pretraining_fcmf.py:320-322 -- CrossEntropyLoss(ignore_index=-100) over the
        vocabulary axis of logits [B, T, V] -- as one forward and one backward kernel (no [B, V, T] permute copy)."""
        from .. import functional as Fn
        return Fn.vocab_cross_entropy(logits, labels, ignore_index)

    @staticmethod
    def _init_weights(module):
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()
