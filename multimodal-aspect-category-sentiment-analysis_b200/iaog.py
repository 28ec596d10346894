"""IAOG Seq2Seq decoder -- consumer of the fusion output (reference
mm_modeling.py:35-132 ``Attention``, :558-613 decoder block, :615-666 ``IAOGDecoder``).

Scope note: this is the first "next" row after the fusion hot path. Every contraction (per-head projections folded
into one GEMM, output projection, FFN, the vocabulary projection), every LayerNorm and the T x T / T x 15 attention
cores (``folded_attention`` with the causal masked_fill flag, keys used as values) run on the fusion path's kernels.
``attention_weights`` (a visualisation by-product in the reference) is therefore not materialised in training mode.

Contract kept from the reference (each looks odd but is what checkpoints were trained with):
  * per-head weight tensors ``w_kx`` / ``w_qx`` of shape [heads, H, dh]; the projected KEYS are also the values;
  * a 2-D mask argument means tril(q_len x k_len) on self- AND cross-attention; masked_fill(-1e4), not -inf;
  * no dropout inside ``Attention``; ``dense.weight`` is tied to ``embedding.weight``;
  * the per-block key/value cache in ``state[2]`` is written but never read by attention1.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import functional as Fn
from .fcmf_framework import mm_modeling as M


class Attention(nn.Module):
    def __init__(self, embed_dim, hidden_dim=None, n_head=1, score_function="scaled_dot_product", dropout=0.1):
        super().__init__()
        if score_function != "scaled_dot_product":
            raise NotImplementedError("the IAOG decoder only instantiates scaled_dot_product attention")
        hidden_dim = hidden_dim or embed_dim // n_head
        self.embed_dim, self.hidden_dim, self.n_head, self.score_function = embed_dim, hidden_dim, n_head, score_function
        self.w_kx = nn.Parameter(torch.empty(n_head, embed_dim, hidden_dim))
        self.w_qx = nn.Parameter(torch.empty(n_head, embed_dim, hidden_dim))
        self.proj = nn.Linear(n_head * hidden_dim, embed_dim)
        self.register_parameter("weight", None)
        nn.init.xavier_uniform_(self.w_kx)
        nn.init.xavier_uniform_(self.w_qx)
        self.attention_weights = None

    def _project(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        """x [B,T,E] times per-head [nh,E,dh] as ONE GEMM against the [nh*dh, E] stacking -> [B*T, slots*dh].

        The reference pairs input ``k.repeat(nh,1,1)[n]`` (batch n % B) with weight ``w.repeat(B,1,1)[n]`` (head n % nh)
        and then places entry n = c*B + b in output slot c of batch b (mm_modeling.py:79-85, 130): slot c of batch b uses
        head (c*B + b) % nh. The GEMM computes every head for every row; the slot->head map is a gather on the head axis."""
        B, T, E = x.shape
        nh, dh = self.n_head, self.hidden_dim
        w2 = w.permute(0, 2, 1).reshape(nh * dh, E)
        y = Fn.linear(x.reshape(B * T, E), w2, None).view(B, T, nh, dh)
        slot_head = (torch.arange(nh, device=x.device).view(1, nh) * B + torch.arange(B, device=x.device).view(B, 1)) % nh
        y = torch.gather(y, 2, slot_head.view(B, 1, nh, 1).expand(B, T, nh, dh))                 # [B, T, slot, dh]
        return y.reshape(B * T, nh * dh)

    def forward(self, k, q, memory_len=None):
        if k.dim() == 2:
            k = k.unsqueeze(1)
        if q.dim() == 2:
            q = q.unsqueeze(1)
        B, k_len, q_len = k.size(0), k.size(1), q.size(1)
        nh, dh = self.n_head, self.hidden_dim
        kx = self._project(k, self.w_kx)                              # [B*k_len, nh*dh]; keys are also the values (line 129)
        qx = self._project(q, self.w_qx)
        if isinstance(memory_len, (list, tuple)):
            memory_len = torch.tensor(memory_len, device=k.device)
        if memory_len is None or memory_len.dim() == 2:
            # training path: no mask, or the 2-D mask = tril(q_len x k_len) on self- AND cross-attention (lines 115-118):
            # one folded launch for all (batch, slot) problems, masked_fill(-1e4) semantics in the kernel
            plan = Fn.AttnPlan(B, nh, dh, causal=memory_len is not None) \
                .add("q", 0, 0, q_len, None, None).add("k", 1, 0, k_len, None, None).add("v", 1, 0, k_len, None, None)
            out = Fn.folded_attention(plan, (qx, kx), None, None)                    # [B*q_len, nh*dh], slot-major features
            self.attention_weights = None                                            # probabilities are not materialised
        else:
            # 1-D valid-length mask (inference helper of the reference): small, kept on PyTorch ops
            kx4 = kx.view(B, k_len, nh, dh).permute(0, 2, 1, 3).float()
            qx4 = qx.view(B, q_len, nh, dh).permute(0, 2, 1, 3).float()
            score = torch.matmul(qx4, kx4.transpose(-1, -2)) / math.sqrt(dh)
            keep = torch.arange(k_len, device=k.device).unsqueeze(0) < memory_len.unsqueeze(1)
            prob = torch.softmax(score.masked_fill(~keep.view(B, 1, 1, k_len), -1e4), dim=-1)
            self.attention_weights = prob.permute(1, 0, 2, 3).reshape(nh * B, q_len, k_len)     # reference layout [nh*B, q, k]
            out = torch.matmul(prob, kx4).permute(0, 2, 1, 3).reshape(B * q_len, nh * dh).to(k.dtype)
        out = Fn.linear(out, self.proj.weight, self.proj.bias)
        return out.view(B, q_len, self.embed_dim), self.attention_weights


class PositionWiseFFN(nn.Module):
    def __init__(self, ffn_num_hiddens, ffn_num_outputs):
        super().__init__()
        self.dense1 = nn.Linear(M.HIDDEN_SIZE, ffn_num_hiddens)
        self.dense2 = nn.Linear(ffn_num_hiddens, ffn_num_outputs)

    def forward(self, x):
        shape = x.shape
        h = M._GeluLinear.apply(x.reshape(-1, shape[-1]), self.dense1.weight, self.dense1.bias)
        return Fn.linear(h, self.dense2.weight, self.dense2.bias).view(*shape[:-1], -1)


class AddNorm(nn.Module):
    def __init__(self, norm_shape, dropout):
        super().__init__()
        self.dropout_p = dropout
        self.ln = M.FCMFLayerNorm(norm_shape)

    def forward(self, X, Y):
        shape = X.shape
        return M._ResidualLayerNorm.apply(Y.reshape(-1, shape[-1]).contiguous(), X.reshape(-1, shape[-1]).contiguous(),
                                          self.ln.weight, self.ln.bias, self.ln.variance_epsilon,
                                          Fn.fresh_drop(self.dropout_p, self.training)).view(shape)   # ln(dropout(Y) + X), :573


class TransformerDecoderBlock(nn.Module):
    def __init__(self, i):
        super().__init__()
        self.i = i
        H, nh = M.HIDDEN_SIZE, M.NUM_ATTENTION_HEADS
        self.attention1 = Attention(H, H // nh, nh, "scaled_dot_product", M.ATTENTION_PROBS_DROPOUT_PROB)
        self.addnorm1 = AddNorm(H, M.ATTENTION_PROBS_DROPOUT_PROB)
        self.attention2 = Attention(H, H // nh, nh, "scaled_dot_product", M.ATTENTION_PROBS_DROPOUT_PROB)
        self.addnorm2 = AddNorm(H, M.ATTENTION_PROBS_DROPOUT_PROB)
        self.ffn = PositionWiseFFN(H, H)
        self.add_norm3 = AddNorm(H, M.ATTENTION_PROBS_DROPOUT_PROB)

    def forward(self, X, state, enc_attention_mask=None, is_train=True):
        enc_outputs, enc_valid_lens = state[0], state[1]
        if state[2][self.i] is not None:                         # written, never read by attention1 (reference :588-601)
            state[2][self.i] = torch.cat((state[2][self.i], X), dim=1)
        dec_valid_lens = None
        if is_train:
            B, T, _ = X.shape
            dec_valid_lens = torch.arange(1, T + 1, device=X.device).repeat(B, 1)
        X2, _ = self.attention1(X, X, dec_valid_lens)
        Y = self.addnorm1(X, X2)
        cross_mask = enc_attention_mask if enc_attention_mask is not None else enc_valid_lens
        Y2, _ = self.attention2(enc_outputs.to(Y.dtype), Y, cross_mask)
        Z = self.addnorm2(Y, Y2)
        return self.add_norm3(Z, self.ffn(Z)), state


class PositionalEncoding(nn.Module):
    def __init__(self):
        super().__init__()
        H, n = M.HIDDEN_SIZE, M.MAX_POSITION_EMBEDDINGS
        P = torch.zeros((1, n, H))
        X = torch.arange(n, dtype=torch.float32).reshape(-1, 1) / torch.pow(
            10000, torch.arange(0, H, 2, dtype=torch.float32) / H)
        P[:, :, 0::2] = torch.sin(X)
        P[:, :, 1::2] = torch.cos(X)
        self.register_buffer("P", P)
        self.dropout_p = M.ATTENTION_PROBS_DROPOUT_PROB            # reference mm_modeling.py:619

    def forward(self, X):
        X = X + self.P[:, :X.size(1), :].to(device=X.device).type_as(X)
        return torch.nn.functional.dropout(X, self.dropout_p, self.training)    # reference :633 (embedding side, torch RNG)


class IAOGDecoder(nn.Module):
    def __init__(self, vocab_size):
        super().__init__()
        self.num_hiddens = M.HIDDEN_SIZE
        self.num_blks = M.NUM_HIDDEN_LAYERS
        self.embedding = nn.Embedding(vocab_size, self.num_hiddens)
        self.pos_encoding = PositionalEncoding()
        self.blks = nn.Sequential()
        for i in range(self.num_blks):
            self.blks.add_module("block" + str(i), TransformerDecoderBlock(i))
        self.dense = nn.Linear(self.num_hiddens, vocab_size)
        self.dense.weight = self.embedding.weight
        self.compute_dtype = None

    def init_state(self, enc_outputs, enc_valid_lens):
        return [enc_outputs, enc_valid_lens, [None] * self.num_blks]

    def forward(self, X, state, enc_attention_mask=None, is_train=True):
        X = self.pos_encoding(self.embedding(X) * math.sqrt(self.num_hiddens))
        if self.compute_dtype is not None:
            X = X.to(self.compute_dtype)
        self._attention_weights = [[None] * len(self.blks) for _ in range(2)]
        for i, blk in enumerate(self.blks):
            X, state = blk(X, state, enc_attention_mask=enc_attention_mask, is_train=is_train)
            self._attention_weights[0][i] = blk.attention1.attention_weights
            self._attention_weights[1][i] = blk.attention2.attention_weights
        B, T, H = X.shape
        V = self.dense.weight.shape[0]
        if X.is_cuda and X.dtype == torch.bfloat16:      # tensor-core path also for V % 8 != 0 (250 002): zero-padded weight rows
            return Fn.vocab_linear(X.reshape(B * T, H), self.dense.weight, self.dense.bias).unflatten(0, (B, T))
        return Fn.linear(X.reshape(B * T, H), self.dense.weight, self.dense.bias).view(B, T, V)

    @property
    def attention_weights(self):
        return self._attention_weights
