"""Plain-torch stand-ins with the CONTRACTS of the kernel Functions (functional.py / ops.py), for CPU tests of the
orchestration that sits on top of them (fold / hoist / index tables / residual gathers / pooled rows / dropout sites). They
take the same arguments as the Functions they replace and are differentiable through autograd. Dropout arguments (ops.Drop)
are honoured with the numpy restatement of the kernels' mask function and the row conventions of include/fcmf_b200.h:
attention row = (problem * heads + head) * Lq + query, LayerNorm / classifier row = row of the input matrix."""
import math

import numpy as np
import torch

from oracle import dropout_mask as DM
from oracle import fcmf_oracle as O

LN_EPS = 1e-12


def _keep(drop, rows, ncols):
    """float32 [len(rows), ncols] = keep / (1 - p) for an ops.Drop (seed_dev unsupported here); None -> None."""
    if drop is None or drop.p <= 0.0:
        return None
    assert drop.seed_dev is None
    return torch.from_numpy(DM.scaled_mask(drop.seed, np.asarray(rows, dtype=np.int64), ncols, drop.p))


def linear(x, weight, bias, act="none", engine=0):
    y = torch.nn.functional.linear(x, weight, bias)
    return torch.tanh(y) if act == "tanh" else y


def mask_additive(mask, n):
    if mask.shape[1] < n:
        raise RuntimeError(f"added_attention_mask of shape {tuple(mask.shape)} is shorter than the {n} positions needed")
    return (1.0 - mask[:, :n].float()) * -10000.0


def box_geometry(boxes, wg_w, wg_b):
    emb = O.box_relational_embedding(boxes).to(wg_w.dtype)                       # [G, NR, NR, 64]
    z = torch.relu(torch.einsum("gijc,hc->ghij", emb, wg_w) + wg_b.view(1, -1, 1, 1))
    return torch.log(torch.clamp(z, min=1e-6))


def _gather(plan, tensors, role):
    HD = plan.heads * plan.dh
    parts = []
    for (slot, col, rows, idx, _inv) in plan.roles[role]:
        t = tensors[slot]
        g = t.view(t.shape[0] // rows, rows, t.shape[1])
        sel = g if idx is None else g[idx.long()]
        assert sel.shape[0] == plan.NP
        parts.append(sel[:, :, col:col + HD])
    x = torch.cat(parts, 1)                                                       # [NP, L, HD]
    return x.view(plan.NP, x.shape[1], plan.heads, plan.dh).permute(0, 2, 1, 3)


def folded_attention(plan, tensors, mask_add=None, bias=None):
    q, k, v = _gather(plan, tensors, "q"), _gather(plan, tensors, "k"), _gather(plan, tensors, "v")
    Lq, Lk = q.shape[2], k.shape[2]
    s = q @ k.transpose(-1, -2) / math.sqrt(plan.dh)
    if mask_add is not None:
        rows = torch.arange(plan.NP) // plan.mask_div
        s = s + mask_add[rows][:, None, None, :Lk]
    if bias is not None:
        s = s + bias
    if plan.causal:                                                   # masked_fill(-1e4) of the IAOG decoder (mm_modeling.py:115-124)
        keep = torch.tril(torch.ones(Lq, Lk, dtype=torch.bool))
        s = s.masked_fill(~keep, -1e4)
    probs = torch.softmax(s, -1)
    m = _keep(plan.drop, np.arange(plan.NP * plan.heads * Lq), Lk)
    if m is not None:
        probs = probs * m.view(plan.NP, plan.heads, Lq, Lk)
    return (probs @ v).permute(0, 2, 1, 3).reshape(plan.NP * Lq, plan.heads * plan.dh)


def _ln(s, g, b):
    return O.tf_layer_norm(s, g, b, LN_EPS)


def layer_tail(a, res_src, res_idx, res_inv, params, engine=0, drop1=None, drop2=None, out_every=0):
    wo, bo, g1, b1, w1, bi1, w2, bi2, g2, b2 = params
    res = res_src if res_idx is None else res_src[res_idx.long()]
    M, H = a.shape[0], wo.shape[0]
    d = torch.nn.functional.linear(a, wo, bo)
    m1, m2 = _keep(drop1, np.arange(M), H), _keep(drop2, np.arange(M), H)
    x1 = _ln((d if m1 is None else d * m1) + res, g1, b1)
    o = torch.nn.functional.linear(O.erf_gelu(torch.nn.functional.linear(x1, w1, bi1)), w2, bi2)
    y = _ln((o if m2 is None else o * m2) + x1, g2, b2)
    return y.view(y.shape[0] // out_every, out_every, y.shape[1])[:, 0, :] if out_every > 1 else y


def classifier_ce(pooled, wc, bc, labels, row_scale=1.0, drop=None):
    m = _keep(drop, np.arange(pooled.shape[0]), pooled.shape[1])
    pooled = pooled.float() if m is None else pooled.float() * m
    logits = torch.nn.functional.linear(pooled, wc, bc)
    if labels is None:
        return logits, logits.new_zeros(())
    return logits, torch.nn.functional.cross_entropy(logits, labels, reduction="sum") * row_scale


class _ResidualLayerNorm:
    """mm_modeling._ResidualLayerNorm.apply(d, res, w, b, eps, drop) = LN(dropout(d) + res)."""
    @staticmethod
    def apply(d, res, w, b, eps, drop=None):
        m = _keep(drop, np.arange(d.shape[0]), d.shape[1])
        return O.tf_layer_norm((d if m is None else d * m) + res, w, b, eps)


class _LayerNormOnly:
    @staticmethod
    def apply(x, w, b, eps):
        return O.tf_layer_norm(x, w, b, eps)


class _GeluLinear:
    @staticmethod
    def apply(x, w, b):
        return O.erf_gelu(torch.nn.functional.linear(x, w, b))


def install(monkeypatch, pkg):
    Fn, ops = pkg("functional"), pkg("ops")
    M = pkg("fcmf_framework.mm_modeling")
    for name, cls in (("_ResidualLayerNorm", _ResidualLayerNorm), ("_LayerNormOnly", _LayerNormOnly), ("_GeluLinear", _GeluLinear)):
        monkeypatch.setattr(M, name, cls)
    for name, fn in (("linear", linear), ("box_geometry", box_geometry), ("folded_attention", folded_attention),
                     ("layer_tail", layer_tail), ("classifier_ce", classifier_ce)):
        monkeypatch.setattr(Fn, name, fn)
    monkeypatch.setattr(ops, "mask_additive", mask_additive)


class _Setter:
    """Same interface as pytest's monkeypatch.setattr, for spawned worker processes; undo() restores the originals."""

    def __init__(self):
        self.saved = []

    def setattr(self, obj, name, value):
        self.saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    def undo(self):
        for obj, name, old in reversed(self.saved):
            setattr(obj, name, old)
        self.saved = []


def install_plain(pkg) -> _Setter:
    st = _Setter()
    install(st, pkg)
    return st
