"""Fusion primitives with the reference's names and parameter layout (reference: fcmf_framework/mm_modeling.py).

Each class keeps the reference's attribute names (so ``state_dict`` keys match and checkpoints load) and its
``forward`` signature; the bodies dispatch to the kernel Functions in ``..functional``. Model dimensions are module
level constants exactly as in the reference (mm_modeling.py:21-30): patch them before constructing a model to get
the XLM-R-large variant (HIDDEN_SIZE=1024, NUM_ATTENTION_HEADS=16, INTERMEDIATE_SIZE=4096).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import ops

HIDDEN_SIZE = 768
NUM_HIDDEN_LAYERS = 12
NUM_ATTENTION_HEADS = 12
INTERMEDIATE_SIZE = 3072
HIDDEN_ACT = "gelu"
HIDDEN_DROPOUT_PROB = 0.1
ATTENTION_PROBS_DROPOUT_PROB = 0.1
MAX_POSITION_EMBEDDINGS = 512
TYPE_VOCAB_SIZE = 2
INITIALIZER_RANGE = 0.02


def _dims():
    return HIDDEN_SIZE, NUM_ATTENTION_HEADS, INTERMEDIATE_SIZE


def gelu(x: torch.Tensor) -> torch.Tensor:
    """erf-GELU (reference mm_modeling.py:10-15); host-side helper, the kernels fuse it into the GEMM epilogue."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


class FCMFLayerNorm(nn.Module):
    """TF-style LayerNorm, epsilon inside the square root (reference mm_modeling.py:158-171)."""

    def __init__(self, hidden_size: int, eps: float = 1e-12):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.bias = nn.Parameter(torch.zeros(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        return _LayerNormOnly.apply(x.reshape(-1, shape[-1]).contiguous(), self.weight, self.bias,
                                    self.variance_epsilon).view(shape)


class _LayerNormOnly(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        y, mean, rstd = ops.ln_fwd(x, None, None, w, b, eps)
        ctx.save_for_backward(x, w, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        ds, dg, db = ops.ln_bwd(dy.contiguous(), None, x, None, None, w, mean, rstd)
        return ds, dg, db, None


class _QKV(nn.Module):
    """query/key/value projections of one attention (reference BertSelfAttention / BertCoAttention __init__)."""

    def __init__(self):
        super().__init__()
        hidden, heads, _ = _dims()
        self.num_attention_heads = heads
        self.attention_head_size = hidden // heads
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(hidden, self.all_head_size)
        self.key = nn.Linear(hidden, self.all_head_size)
        self.value = nn.Linear(hidden, self.all_head_size)
        self.dropout_p = ATTENTION_PROBS_DROPOUT_PROB

    def _attend(self, q_in: torch.Tensor, kv_in: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
        B, Lq, H = q_in.shape
        Lk = kv_in.shape[1]
        q = Fn.linear(q_in.reshape(B * Lq, H), self.query.weight, self.query.bias)
        w_kv = torch.cat((self.key.weight, self.value.weight), 0)
        b_kv = torch.cat((self.key.bias, self.value.bias), 0)
        kv = Fn.linear(kv_in.reshape(B * Lk, H), w_kv, b_kv)
        plan = Fn.AttnPlan(B, self.num_attention_heads, self.attention_head_size,
                           drop=Fn.fresh_drop(self.dropout_p, self.training)) \
            .add("q", 0, 0, Lq, None, None).add("k", 1, 0, Lk, None, None).add("v", 1, H, Lk, None, None)
        mask_add = None
        if mask is not None:                     # the reference passes the additive [B,1,1,Lk] mask
            mask_add = mask.to(torch.float32).expand(B, 1, 1, Lk).reshape(B, Lk).contiguous()
        return Fn.folded_attention(plan, (q, kv), mask_add, None).view(B, Lq, H)


class BertSelfAttention(_QKV):
    def forward(self, hidden_states, attention_mask):
        return self._attend(hidden_states, hidden_states, attention_mask)


class BertCoAttention(_QKV):
    def forward(self, s1_hidden_states, s2_hidden_states, s2_attention_mask):
        return self._attend(s1_hidden_states, s2_hidden_states, s2_attention_mask)


class BertSelfOutput(nn.Module):
    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(HIDDEN_SIZE, HIDDEN_SIZE)
        self.LayerNorm = FCMFLayerNorm(HIDDEN_SIZE, eps=1e-12)
        self.dropout_p = HIDDEN_DROPOUT_PROB

    def forward(self, hidden_states, input_tensor):
        shape = input_tensor.shape
        d = Fn.linear(hidden_states.reshape(-1, hidden_states.shape[-1]), self.dense.weight, self.dense.bias)
        return _ResidualLayerNorm.apply(d, input_tensor.reshape(-1, shape[-1]).contiguous(), self.LayerNorm.weight,
                                        self.LayerNorm.bias, self.LayerNorm.variance_epsilon,
                                        Fn.fresh_drop(self.dropout_p, self.training)).view(shape)


class _ResidualLayerNorm(torch.autograd.Function):
    """LN(dropout(d) + res) -- BertSelfOutput / BertOutput (reference mm_modeling.py:276-280, 324-328)."""

    @staticmethod
    def forward(ctx, d, res, w, b, eps, drop):
        y, mean, rstd = ops.ln_fwd(d, res, None, w, b, eps, drop=drop)
        ctx.drop = drop
        ctx.save_for_backward(d, res, w, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        d, res, w, mean, rstd = ctx.saved_tensors
        ds, dd, dg, db = ops.ln_bwd_drop(dy.contiguous(), None, d, res, None, w, mean, rstd, ctx.drop)
        return dd, ds, dg, db, None, None


class BertAttention(nn.Module):
    def __init__(self):
        super().__init__()
        self.self = BertSelfAttention()
        self.output = BertSelfOutput()

    def forward(self, input_tensor, attention_mask):
        return self.output(self.self(input_tensor, attention_mask), input_tensor)


class BertCrossAttention(nn.Module):
    def __init__(self):
        super().__init__()
        self.self = BertCoAttention()
        self.output = BertSelfOutput()

    def forward(self, s1_input_tensor, s2_input_tensor, s2_attention_mask):
        return self.output(self.self(s1_input_tensor, s2_input_tensor, s2_attention_mask), s1_input_tensor)


class BertIntermediate(nn.Module):
    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(HIDDEN_SIZE, INTERMEDIATE_SIZE)

    def forward(self, hidden_states):
        shape = hidden_states.shape
        y = _GeluLinear.apply(hidden_states.reshape(-1, shape[-1]), self.dense.weight, self.dense.bias)
        return y.view(*shape[:-1], -1)


class _GeluLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        g, pre = ops.gemm_tn(x, ops.cast_matrix(w, x.dtype), b, ops.EPI_GELU, want_aux=True)
        ctx.save_for_backward(x, w, pre)
        return g

    @staticmethod
    def backward(ctx, dg):
        x, w, pre = ctx.saved_tensors
        # dpre = dg * gelu'(pre): identity "GEMM-free" form via the epilogue of the input-gradient GEMM is only
        # available when the producer of dg is a GEMM (layer_tail does that); here dg arrives from autograd.
        dpre = (dg.float() * _gelu_grad(pre.float())).to(x.dtype).contiguous()
        dx = ops.gemm_tn(dpre, ops.cast_matrix(w, x.dtype, transpose=True))
        dw, db = ops.gemm_wgrad(dpre, x)
        return dx, dw, db


def _gelu_grad(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


class BertOutput(nn.Module):
    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(INTERMEDIATE_SIZE, HIDDEN_SIZE)
        self.LayerNorm = FCMFLayerNorm(HIDDEN_SIZE, eps=1e-12)
        self.dropout_p = HIDDEN_DROPOUT_PROB

    def forward(self, hidden_states, input_tensor):
        shape = input_tensor.shape
        d = Fn.linear(hidden_states.reshape(-1, hidden_states.shape[-1]), self.dense.weight, self.dense.bias)
        return _ResidualLayerNorm.apply(d, input_tensor.reshape(-1, shape[-1]).contiguous(), self.LayerNorm.weight,
                                        self.LayerNorm.bias, self.LayerNorm.variance_epsilon,
                                        Fn.fresh_drop(self.dropout_p, self.training)).view(shape)


def _tail(layer, ctx_rows, residual):
    from ..fusion import _tail_params
    shape = residual.shape
    so, out = layer.attention.output, layer.output
    y = Fn.layer_tail(ctx_rows.reshape(-1, shape[-1]), residual.reshape(-1, shape[-1]).contiguous(), None, None,
                      _tail_params(layer), drop1=Fn.fresh_drop(so.dropout_p, so.training),
                      drop2=Fn.fresh_drop(out.dropout_p, out.training))
    return y.view(shape)


class BertLayer(nn.Module):
    """Self-attention encoder layer (reference mm_modeling.py:331-342); attention core + one fused tail."""

    def __init__(self):
        super().__init__()
        self.attention = BertAttention()
        self.intermediate = BertIntermediate()
        self.output = BertOutput()

    def forward(self, hidden_states, attention_mask):
        ctx_rows = self.attention.self(hidden_states, attention_mask)
        return _tail(self, ctx_rows, hidden_states)


class BertCrossAttentionLayer(nn.Module):
    """Cross-attention encoder layer (reference mm_modeling.py:344-355)."""

    def __init__(self):
        super().__init__()
        self.attention = BertCrossAttention()
        self.intermediate = BertIntermediate()
        self.output = BertOutput()

    def forward(self, s1_hidden_states, s2_hidden_states, s2_attention_mask):
        ctx_rows = self.attention.self(s1_hidden_states, s2_hidden_states, s2_attention_mask)
        return _tail(self, ctx_rows, s1_hidden_states)


class MultimodalEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer()])

    def forward(self, hidden_states, attention_mask, output_all_encoded_layers=True) -> List[torch.Tensor]:
        outs = []
        for blk in self.layer:
            hidden_states = blk(hidden_states, attention_mask)
            if output_all_encoded_layers:
                outs.append(hidden_states)
        return outs if output_all_encoded_layers else [hidden_states]


class BertCrossEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer = nn.ModuleList([BertCrossAttentionLayer()])

    def forward(self, s1_hidden_states, s2_hidden_states, s2_attention_mask, output_all_encoded_layers=True):
        outs = []
        for blk in self.layer:
            s1_hidden_states = blk(s1_hidden_states, s2_hidden_states, s2_attention_mask)
            if output_all_encoded_layers:
                outs.append(s1_hidden_states)
        return outs if output_all_encoded_layers else [s1_hidden_states]


class BertPooler(nn.Module):
    """tanh(dense(first token)) (reference mm_modeling.py:419-431): a strided-row GEMM with a tanh epilogue."""

    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(HIDDEN_SIZE, HIDDEN_SIZE)

    def forward(self, hidden_states):
        return Fn.linear(hidden_states[:, 0], self.dense.weight, self.dense.bias, act="tanh")


class FeatureExtractor(nn.Module):
    """Upstream text encoder (reference mm_modeling.py:433-446): the HF module is the parameter container (``cell``:
    state_dict keys, checkpoint loading, resize_token_embeddings unchanged). ``use_kernels = True`` executes its encoder
    layers and pooler on this library's kernels (``..xlmr``); attention probabilities are then
    not materialised (third return value None). Default: the stock HF forward."""

    def __init__(self, pretrained_path=None, cell=None):
        super().__init__()
        if cell is None:
            from transformers import AutoModel
            cell = AutoModel.from_pretrained(pretrained_path, local_files_only=True, attn_implementation="eager")
        self.cell = cell
        self.use_kernels = False
        self.compute_dtype = None
        self.engine = 0

    def forward(self, input_ids, token_type_ids, attention_mask):
        if self.use_kernels and input_ids.is_cuda:
            from .. import xlmr
            return xlmr.encode(self.cell, input_ids, token_type_ids, attention_mask, self.compute_dtype, self.engine, self.training)
        out = self.cell(input_ids=input_ids, token_type_ids=token_type_ids, attention_mask=attention_mask,
                        output_attentions=True)
        return out[0], out[1], out[2]
