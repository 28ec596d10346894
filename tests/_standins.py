"""Plain-torch stand-ins with the CONTRACTS of the kernel Functions (functional.py / ops.py), for CPU tests of the
orchestration that sits on top of them (fold / hoist / index tables / residual gathers / pooled rows). They take the same
arguments as the Functions they replace and are differentiable through autograd; dropout arguments must be None (eval)."""
import math

import torch

from oracle import fcmf_oracle as O

LN_EPS = 1e-12


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
    assert plan.drop is None and not plan.causal
    q, k, v = _gather(plan, tensors, "q"), _gather(plan, tensors, "k"), _gather(plan, tensors, "v")
    s = q @ k.transpose(-1, -2) / math.sqrt(plan.dh)
    if mask_add is not None:
        rows = torch.arange(plan.NP) // plan.mask_div
        s = s + mask_add[rows][:, None, None, : s.shape[-1]]
    if bias is not None:
        s = s + bias
    ctx = torch.softmax(s, -1) @ v
    return ctx.permute(0, 2, 1, 3).reshape(plan.NP * q.shape[2], plan.heads * plan.dh)


def _ln(s, g, b):
    return O.tf_layer_norm(s, g, b, LN_EPS)


def layer_tail(a, res_src, res_idx, res_inv, params, engine=0, drop1=None, drop2=None, out_every=0):
    assert drop1 is None and drop2 is None
    wo, bo, g1, b1, w1, bi1, w2, bi2, g2, b2 = params
    res = res_src if res_idx is None else res_src[res_idx.long()]
    x1 = _ln(torch.nn.functional.linear(a, wo, bo) + res, g1, b1)
    y = _ln(torch.nn.functional.linear(O.erf_gelu(torch.nn.functional.linear(x1, w1, bi1)), w2, bi2) + x1, g2, b2)
    return y.view(y.shape[0] // out_every, out_every, y.shape[1])[:, 0, :] if out_every > 1 else y


def classifier_ce(pooled, wc, bc, labels, row_scale=1.0, drop=None):
    assert drop is None
    logits = torch.nn.functional.linear(pooled.float(), wc, bc)
    if labels is None:
        return logits, logits.new_zeros(())
    return logits, torch.nn.functional.cross_entropy(logits, labels, reduction="sum") * row_scale


def install(monkeypatch, pkg):
    Fn, ops = pkg("functional"), pkg("ops")
    for name, fn in (("linear", linear), ("box_geometry", box_geometry), ("folded_attention", folded_attention),
                     ("layer_tail", layer_tail), ("classifier_ce", classifier_ce)):
        monkeypatch.setattr(Fn, name, fn)
    monkeypatch.setattr(ops, "mask_additive", mask_additive)
