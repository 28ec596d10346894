"""The XLM-R text encoder on the fusion path's kernels
(reference: FeatureExtractor, fcmf_framework/mm_modeling.py:433-446, which calls the stock
Hugging Face ``XLMRobertaModel``).

Once the fusion runs live-rows the text encoder is > 95 % of a training step (179.7 vs 5.9 GFLOP/sample forward), and it is
the same BERT layer the fusion path already has kernels for: fused QKV GEMM -> folded attention (all B*A sequences and
heads in one launch) -> dense + dropout + residual + LayerNorm -> GELU FFN -> dense + dropout + residual + LayerNorm.
``KernelFeatureExtractor`` keeps the HF module as the parameter container (``cell``: state_dict keys, checkpoint loading
and ``resize_token_embeddings`` unchanged) and replaces only the execution of the 12 encoder layers and the pooler.
Embeddings (three gathers + LayerNorm + dropout over [B, L, H]) stay on the HF module.

Differences to state: attention probabilities are not materialised (``attentions`` is None; the HF path is used when a
caller asks for them), masked keys get -10000 (fusion-path convention) instead of finfo.min -- both vanish in the softmax
unless every key of a row is masked.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import functional as Fn
from . import ops


class KernelFeatureExtractor(nn.Module):
    def __init__(self, pretrained_path=None, cell: Optional[nn.Module] = None):
        super().__init__()
        if cell is None:
            from transformers import AutoModel
            cell = AutoModel.from_pretrained(pretrained_path, local_files_only=True, attn_implementation="eager")
        self.cell = cell
        self.compute_dtype: Optional[torch.dtype] = None     # None: fp32 parameters -> fp32 kernels; torch.bfloat16 for the tensor-core path
        self.engine = 0
        self.return_attentions = False                       # True: run the stock HF forward (it materialises the probabilities)

    def _hf(self, input_ids, token_type_ids, attention_mask):
        out = self.cell(input_ids=input_ids, token_type_ids=token_type_ids, attention_mask=attention_mask,
                        output_attentions=True)
        return out[0], out[1], out[2]

    def forward(self, input_ids, token_type_ids, attention_mask):
        if self.return_attentions or not input_ids.is_cuda:
            return self._hf(input_ids, token_type_ids, attention_mask)
        return encode(self.cell, input_ids, token_type_ids, attention_mask, self.compute_dtype, self.engine, self.training)


def encode(cell, input_ids, token_type_ids, attention_mask, compute_dtype=None, engine: int = 0, training: bool = False):
    """(sequence_output [B, L, H], pooled [B, H] or None, None) of the HF XLM-R module ``cell`` with its 12 encoder layers
    and pooler executed by this library's kernels (fused QKV GEMM, folded attention over all sequences and heads, layer tails)."""
    cfg = cell.config
    B, L = input_ids.shape
    H, nh = cfg.hidden_size, cfg.num_attention_heads
    dh = H // nh
    dt = compute_dtype or torch.float32
    if torch.is_autocast_enabled():
        dt = torch.bfloat16
    x = cell.embeddings(input_ids=input_ids, token_type_ids=token_type_ids)              # [B, L, H], LayerNorm + dropout inside
    x2 = x.to(dt).reshape(B * L, H)
    if attention_mask is None:
        attention_mask = torch.ones((B, L), dtype=torch.int64, device=input_ids.device)
    mask_add = ops.mask_additive(attention_mask.reshape(B, -1), L)                       # [B, L] fp32, -10000 on padded keys
    for layer in cell.encoder.layer:
        att, so, it, out = layer.attention.self, layer.attention.output, layer.intermediate, layer.output
        w_qkv = torch.cat((att.query.weight, att.key.weight, att.value.weight), 0)
        b_qkv = torch.cat((att.query.bias, att.key.bias, att.value.bias), 0)
        qkv = Fn.linear(x2, w_qkv, b_qkv, engine=engine)                                 # [B*L, 3H]
        plan = Fn.AttnPlan(B, nh, dh, mask_div=1, drop=Fn.fresh_drop(float(cfg.attention_probs_dropout_prob), training)) \
            .add("q", 0, 0, L, None, None).add("k", 0, H, L, None, None).add("v", 0, 2 * H, L, None, None)
        ctx = Fn.folded_attention(plan, (qkv,), mask_add, None)                          # [B*L, H]
        params = (so.dense.weight, so.dense.bias, so.LayerNorm.weight, so.LayerNorm.bias, it.dense.weight, it.dense.bias,
                  out.dense.weight, out.dense.bias, out.LayerNorm.weight, out.LayerNorm.bias)
        x2 = Fn.layer_tail(ctx, x2, None, None, params, engine=engine,
                           drop1=Fn.fresh_drop(float(cfg.hidden_dropout_prob), training),
                           drop2=Fn.fresh_drop(float(cfg.hidden_dropout_prob), training), eps=float(cfg.layer_norm_eps))
    seq = x2.view(B, L, H)
    pooled = None
    if getattr(cell, "pooler", None) is not None:
        pooled = Fn.linear(seq[:, 0, :], cell.pooler.dense.weight, cell.pooler.dense.bias, act="tanh", engine=engine)
    return seq, pooled, None
