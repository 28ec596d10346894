"""The torch custom-op layer: autograd Functions whose forward AND backward are C-ABI kernel calls.

Granularity follows the reference's module boundaries so that each Function replaces one module 1:1
(SURVEY.md section 8(b2)):
    linear(...)          nn.Linear (+ tanh of BertPooler, mm_modeling.py:425-431)
    layer_tail(...)      BertSelfOutput -> BertIntermediate -> BertOutput (mm_modeling.py:276-328)
    folded_attention(...) BertSelfAttention / BertCoAttention core (mm_modeling.py:193-266) and box_attention
                         (roi_modeling.py:14-47), all problems and heads in one launch
    box_geometry(...)    BoxRelationalEmbedding + WGs + relu + log/clamp (roi_modeling.py:79-162, 40)
    classifier_ce(...)   classifier + CrossEntropyLoss (fcmf_multimodal.py:50, run_multimodal_fcmf.py:290)
"""
from __future__ import annotations

import math
import os
import threading
from typing import List, Optional, Sequence, Tuple

import torch
from torch.autograd import Function

from . import ops
from ._lib import ENGINE_AUTO, ENGINE_SIMT, EPI_DGELU, EPI_GELU, EPI_NONE, EPI_TANH

Tensor = torch.Tensor
_DEBUG = os.environ.get("FCMF_DEBUG", "0") not in ("", "0")     # host-side argument checks that cost a device synchronisation


# ------------------------------------------------------------------------------------------------- dropout seeds
# A dropout SITE is one nn.Dropout call of the reference. Its mask is a pure function of (site seed, row, column) that
# the forward and backward kernels regenerate (no mask tensor); the site seed is step_seed + site_id * an odd constant.
# step seeds come from torch's CPU generator, so torch.manual_seed() makes a training run reproducible.
_GOLDEN64 = 0x9E3779B97F4A7C15
_TLS = threading.local()                      # .seed_dev: device int64 [1] added to every seed by the kernels (CUDA-graph replay);
                                              # per host thread: the nn.DataParallel fallback runs one thread per GPU


def _seed_dev() -> Optional[Tensor]:
    return getattr(_TLS, "seed_dev", None)


def new_step_seed() -> int:
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def site_seed(step_seed: int, site: int) -> int:
    return (step_seed + site * _GOLDEN64) & 0xFFFFFFFFFFFFFFFF


def site_drop(step_seed: Optional[int], site: int, p: float) -> Optional[ops.Drop]:
    """Drop description of `site` for this step; None when dropout is off (eval mode: step_seed None, or p == 0)."""
    if step_seed is None or p <= 0.0:
        return None
    return ops.Drop(p, site_seed(step_seed, site), _seed_dev())


def fresh_drop(p: float, training: bool) -> Optional[ops.Drop]:
    """A site with its own fresh seed (module-level calls outside the folded path)."""
    return ops.Drop(p, new_step_seed(), _seed_dev()) if (training and p > 0.0) else None


def set_seed_device_tensor(t: Optional[Tensor]) -> None:
    """Device int64 [1] that every dropout kernel adds to its seed; graphed.py bumps it inside the captured graph so
    that each replay draws new masks. None = off."""
    _TLS.seed_dev = t


def _c2(t: Tensor) -> Tensor:
    """2-D gradient with unit column stride (autograd may hand us expanded/strided grads)."""
    return t if (t.dim() == 2 and (t.shape[1] == 1 or t.stride(1) == 1) and t.stride(0) >= t.shape[1]) else t.contiguous()


# ------------------------------------------------------------------------------------------------- linear
class _Linear(Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], act: str, engine: int):
        w = ops.cast_matrix(weight, x.dtype)
        y = ops.gemm_tn(x, w, bias, EPI_TANH if act == "tanh" else EPI_NONE, engine=engine)
        ctx.act, ctx.engine, ctx.has_bias = act, engine, bias is not None
        ctx.save_for_backward(x, weight, y if act == "tanh" else None)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, weight, y = ctx.saved_tensors
        dy = _c2(dy)
        if ctx.act == "tanh":
            dy = ops.dtanh(dy.contiguous(), y)
        dx = None
        if ctx.needs_input_grad[0]:
            wt = ops.cast_matrix(weight, x.dtype, transpose=True)            # [K, N]
            dx = ops.gemm_tn(dy, wt, None, EPI_NONE, engine=ctx.engine)
        dw, db = ops.gemm_wgrad(dy, x, want_bias=ctx.has_bias, engine=ctx.engine)
        return dx, dw, db, None, None


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], act: str = "none", engine: int = ENGINE_AUTO) -> Tensor:
    """y = act(x @ weight^T + bias); x is a 2-D row view (rows may be strided), weight/bias are fp32 parameters."""
    return _Linear.apply(x, weight, bias, act, engine)


# ------------------------------------------------------------------------------------------------- layer tail
class _LayerTail(Function):
    """ctx_rows -> LN1(dense(ctx) + residual) -> GELU FFN -> LN2(dense2 + .)   (one BERT layer after attention)."""

    @staticmethod
    def forward(ctx, a: Tensor, res_src: Tensor, res_idx: Optional[Tensor], res_inv: Optional[Tensor],
                wo, bo, g1, b1, w1, bi1, w2, bi2, g2, b2, engine: int, drop1: Optional[ops.Drop], drop2: Optional[ops.Drop],
                out_every: int, eps: float):
        dt = a.dtype
        d = ops.gemm_tn(a, ops.cast_matrix(wo, dt), bo, EPI_NONE, engine=engine)
        x1, mean1, rstd1 = ops.ln_fwd(d, res_src, res_idx, g1, b1, eps, drop=drop1)  # LN(dropout(dense) + residual)
        g, pre = ops.gemm_tn(x1, ops.cast_matrix(w1, dt), bi1, EPI_GELU, engine=engine, want_aux=True)
        o = ops.gemm_tn(g, ops.cast_matrix(w2, dt), bi2, EPI_NONE, engine=engine)
        y, mean2, rstd2 = ops.ln_fwd(o, x1, None, g2, b2, eps, drop=drop2)
        ctx.engine, ctx.drop1, ctx.drop2, ctx.out_every = engine, drop1, drop2, out_every
        ctx.save_for_backward(a, res_src, res_idx, res_inv, wo, g1, w1, w2, g2, d, x1, pre, g, o, mean1, rstd1, mean2, rstd2)
        if out_every > 1:            # the caller consumes rows 0, out_every, ... only (token 0 of every problem): hand out that view
            return y.view(y.shape[0] // out_every, out_every, y.shape[1])[:, 0, :]
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        (a, res_src, res_idx, res_inv, wo, g1, w1, w2, g2, d, x1, pre, g, o, mean1, rstd1, mean2, rstd2) = ctx.saved_tensors
        eng, dt = ctx.engine, a.dtype
        dy = dy.contiguous()
        # ds2 = grad of s2 = dropout(o) + x1 (goes to x1), do2 = grad of o (== ds2 without dropout); with out_every the
        # incoming gradient is compact (one row per problem) and the zero rows are never materialised
        ds2, do2, dg2, db2 = ops.ln_bwd_drop(dy, None, o, x1, None, g2, mean2, rstd2, ctx.drop2,
                                             dy_every=ctx.out_every if ctx.out_every > 1 else 0)
        dw2, dbi2 = ops.gemm_wgrad(do2, g, engine=eng)
        dpre = ops.gemm_tn(do2, ops.cast_matrix(w2, dt, transpose=True), None, EPI_DGELU, aux=pre, engine=eng)
        dw1, dbi1 = ops.gemm_wgrad(dpre, x1, engine=eng)
        dx1 = ops.gemm_tn(dpre, ops.cast_matrix(w1, dt, transpose=True), None, EPI_NONE, engine=eng)
        # (dx1 + ds2) through LN1: ds1 = grad of s1 = dropout(d) + residual (goes to the residual), dd = grad of d
        ds1, dd, dg1, db1 = ops.ln_bwd_drop(dx1, ds2, d, res_src, res_idx, g1, mean1, rstd1, ctx.drop1)
        dwo, dbo = ops.gemm_wgrad(dd, a, engine=eng)
        da = ops.gemm_tn(dd, ops.cast_matrix(wo, dt, transpose=True), None, EPI_NONE, engine=eng) if ctx.needs_input_grad[0] else None
        dres = None
        if ctx.needs_input_grad[1]:
            if res_idx is None:
                dres = ds1
            else:
                dres = ops.gather_sum_rows(ds1, res_inv, res_src.shape[0], res_inv.shape[1])
        return da, dres, None, None, dwo, dbo, dg1, db1, dw1, dbi1, dw2, dbi2, dg2, db2, None, None, None, None, None


def layer_tail(a: Tensor, res_src: Tensor, res_idx: Optional[Tensor], res_inv: Optional[Tensor], params: Sequence[Tensor],
               engine: int = ENGINE_AUTO, drop1: Optional[ops.Drop] = None, drop2: Optional[ops.Drop] = None,
               out_every: int = 0, eps: float = ops.LN_EPS) -> Tensor:
    """params = (Wo, bo, ln1.w, ln1.b, W1, b1, W2, b2, ln2.w, ln2.b). Row m of `a` takes residual row
    res_idx[m] of res_src (identity when res_idx is None); res_inv [rows(res_src), G] lists, for every residual
    row, the rows of `a` that used it (-1 padded) so that the backward reduction needs no atomics.
    drop1 / drop2: hidden dropout of BertSelfOutput / BertOutput (mm_modeling.py:278, 326), mask row = row of `a`.
    out_every > 1: every row is computed, but only rows 0, out_every, 2*out_every, ... are returned ([rows/out_every, H], a
    view) -- what BertPooler reads of a per-image branch; the backward pass then takes a compact gradient."""
    return _LayerTail.apply(a, res_src, res_idx, res_inv, *params, engine, drop1, drop2, out_every, eps)


# ------------------------------------------------------------------------------------------------- attention
class AttnPlan:
    """Static description of one folded attention launch: which tensor/columns/rows feed each segment."""

    def __init__(self, NP: int, heads: int, dh: int, mask_div: int = 1, causal: bool = False,
                 drop: Optional[ops.Drop] = None, engine: int = ENGINE_AUTO):
        self.NP, self.heads, self.dh, self.mask_div, self.causal = NP, heads, dh, mask_div, causal
        self.engine = engine                          # attention engine of this launch (0: the process default)
        self.drop = drop                              # dropout on the probabilities, mask row = (p*heads + h)*Lq + i
        self.roles = {"q": [], "k": [], "v": []}     # lists of (tensor_slot, col, rows, idx, inv)

    def add(self, role: str, slot: int, col: int, rows: int, idx: Optional[Tensor], inv: Optional[Tensor]):
        """idx [NP] int32: group of problem p (None = p). inv [groups, G] int32: problems of each group (None = identity)."""
        self.roles[role].append((slot, col, rows, idx, inv))
        return self

    @property
    def Lq(self):
        return sum(r[2] for r in self.roles["q"])

    @property
    def Lk(self):
        return sum(r[2] for r in self.roles["k"])


def _desc(plan: AttnPlan, tensors: Sequence[Tensor], mask_add, bias):
    segs = {role: [ops.SegSpec(tensors[s], col, rows, idx) for (s, col, rows, idx, _) in plan.roles[role]]
            for role in ("q", "k", "v")}
    return ops.make_attn_desc(segs["q"], segs["k"], segs["v"], plan.NP, plan.heads, plan.dh,
                              1.0 / math.sqrt(plan.dh), mask_add, plan.mask_div, bias, plan.causal, plan.drop, plan.engine)


class _FoldedAttention(Function):
    @staticmethod
    def forward(ctx, plan: AttnPlan, mask_add: Optional[Tensor], bias: Optional[Tensor], *tensors: Tensor):
        desc = _desc(plan, tensors, mask_add, bias)
        out, lse = ops.attn_fwd(desc, plan.Lq, tensors[0].dtype, tensors[0].device)
        ctx.plan = plan
        ctx.n = len(tensors)
        ctx.save_for_backward(mask_add, bias, out, lse, *tensors)
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        plan: AttnPlan = ctx.plan
        mask_add, bias, out, lse = ctx.saved_tensors[:4]
        tensors = ctx.saved_tensors[4:]
        desc = _desc(plan, tensors, mask_add, bias)
        Lq, Lk, HD = plan.Lq, plan.Lk, plan.heads * plan.dh
        want_dbias = bias is not None and ctx.needs_input_grad[2]
        dq, dk, dv, dbias = ops.attn_bwd(desc, Lq, Lk, out, _c2(dout), lse, want_dbias)
        grads: List[Optional[Tensor]] = [None] * len(tensors)
        covered = [set() for _ in tensors]
        for role, per_problem, L in (("q", dq, Lq), ("k", dk, Lk), ("v", dv, Lk)):
            off = 0
            for (slot, col, rows, idx, inv) in plan.roles[role]:
                t = tensors[slot]
                if grads[slot] is None:
                    grads[slot] = torch.empty_like(t, memory_format=torch.contiguous_format)
                again = col in covered[slot]             # e.g. keys used as values: the second write accumulates
                covered[slot].add(col)
                dst = grads[slot][:, col:col + HD]
                n_groups = t.shape[0] // rows
                if inv is None:          # one problem per group, in order
                    inv_rows = _identity_rows(plan.NP, L, off, rows, t.device)
                    G = 1
                else:
                    inv_rows = _group_rows(inv, L, off, rows)
                    G = inv.shape[1]
                ops.gather_sum_rows(per_problem, inv_rows, n_groups * rows, G, out=dst, accumulate=again)
                off += rows
        for slot, t in enumerate(tensors):
            if grads[slot] is not None and len(covered[slot]) * HD < t.shape[1]:
                raise RuntimeError("folded_attention: every column block of a packed tensor must be used by a segment")
        return (None, None, dbias) + tuple(grads)


_ROW_CACHE = {}


def _identity_rows(NP: int, L: int, off: int, rows: int, device) -> Tensor:
    key = ("id", NP, L, off, rows, str(device))
    if key not in _ROW_CACHE:
        p = torch.arange(NP, device=device, dtype=torch.int32).view(NP, 1)
        r = torch.arange(rows, device=device, dtype=torch.int32).view(1, rows)
        _ROW_CACHE[key] = (p * L + off + r).reshape(-1, 1).contiguous()
    return _ROW_CACHE[key]


def _group_rows(inv: Tensor, L: int, off: int, rows: int) -> Tensor:
    """inv [groups, G] (problem ids, -1 padded) -> row ids [groups*rows, G] into a per-problem [NP*L, HD] buffer."""
    key = ("grp", inv.data_ptr(), tuple(inv.shape), L, off, rows)
    if key not in _ROW_CACHE:
        r = torch.arange(rows, device=inv.device, dtype=torch.int32).view(1, rows, 1)
        p = inv.view(inv.shape[0], 1, inv.shape[1])
        rows_id = torch.where(p >= 0, p * L + off + r, torch.full_like(p, -1).expand(-1, rows, -1))
        _ROW_CACHE[key] = (rows_id.reshape(-1, inv.shape[1]).contiguous(), inv)      # keep inv alive: key uses its address
    return _ROW_CACHE[key][0]


def folded_attention(plan: AttnPlan, tensors: Sequence[Tensor], mask_add: Optional[Tensor] = None,
                     bias: Optional[Tensor] = None) -> Tensor:
    """softmax(q k^T / sqrt(dh) + mask + bias) v for plan.NP problems x plan.heads heads in one launch.
    Returns [NP * Lq, heads*dh]."""
    return _FoldedAttention.apply(plan, mask_add, bias, *tensors)


# ------------------------------------------------------------------------------------------------- geometry
class _BoxGeometry(Function):
    @staticmethod
    def forward(ctx, boxes: Tensor, wg_w: Tensor, wg_b: Tensor):
        emb, bias = ops.box_geometry_fwd(boxes, wg_w.contiguous(), wg_b.contiguous(), wg_w.shape[0])
        ctx.save_for_backward(emb, wg_w, wg_b)
        return bias

    @staticmethod
    def backward(ctx, dbias: Tensor):
        emb, wg_w, wg_b = ctx.saved_tensors
        dw, db = ops.box_geometry_bwd(emb, wg_w.contiguous(), wg_b.contiguous(), dbias)
        return None, dw, db


def box_geometry(boxes: Tensor, wg_w: Tensor, wg_b: Tensor) -> Tensor:
    """boxes f64 [G,NR,4], wg_w [heads,64], wg_b [heads] -> additive score bias [G, heads, NR, NR] (fp32)."""
    return _BoxGeometry.apply(boxes, wg_w, wg_b)


# ------------------------------------------------------------------------------------------------- head
class _ClassifierCE(Function):
    @staticmethod
    def forward(ctx, pooled: Tensor, wc: Tensor, bc: Tensor, labels: Optional[Tensor], row_scale: float,
                drop: Optional[ops.Drop]):
        if labels is not None and _DEBUG:       # nn.CrossEntropyLoss raises on labels outside [0, C); the kernel ignores such rows
            bad = (labels < 0) | (labels >= wc.shape[0])
            if bool(bad.any()):
                raise ValueError(f"classifier_ce: {int(bad.sum())} label(s) outside [0, {wc.shape[0]}) (checked because FCMF_DEBUG is set)")
        logits, probs, loss_rows = ops.cls_ce_fwd(pooled.contiguous(), wc.contiguous(), bc.contiguous(), labels, drop)
        ctx.row_scale, ctx.drop = row_scale, drop
        ctx.has_labels = labels is not None
        ctx.set_materialize_grads(False)          # an unused output arrives as None, not as a zero tensor
        ctx.save_for_backward(pooled, wc, probs, labels)
        loss = loss_rows.sum() * row_scale if labels is not None else logits.new_zeros(())
        return logits, loss

    @staticmethod
    def backward(ctx, dlogits: Optional[Tensor], dloss: Optional[Tensor]):
        pooled, wc, probs, labels = ctx.saved_tensors
        pooled = pooled.contiguous()
        out = None
        if ctx.has_labels and dloss is not None:
            dp, dw, db = ops.cls_ce_bwd(pooled, wc.contiguous(), probs, labels.to(torch.int64).contiguous(), None, ctx.row_scale, ctx.drop)
            out = [dp * dloss.to(dp.dtype), dw * dloss, db * dloss]
        if dlogits is not None:
            dp, dw, db = ops.cls_ce_bwd(pooled, wc.contiguous(), None, None, dlogits, 1.0, ctx.drop)
            out = [dp, dw, db] if out is None else [out[0] + dp, out[1] + dw, out[2] + db]
        if out is None:
            return None, None, None, None, None, None
        return out[0], out[1], out[2], None, None, None


def classifier_ce(pooled: Tensor, wc: Tensor, bc: Tensor, labels: Optional[Tensor], row_scale: float = 1.0,
                  drop: Optional[ops.Drop] = None):
    """(logits [R,C] fp32, loss scalar = row_scale * sum_r CE(logits[r], labels[r])); labels=None -> logits only.
    Labels outside [0, C) contribute 0 loss and 0 gradient (and the sum is still scaled by row_scale): the reference's labels
    are always in range (vimacsa_dataset.py builds them from a 4-entry polarity map), where torch would raise; set
    FCMF_DEBUG=1 to have them checked on the host (one synchronisation per call).
    drop: dropout on `pooled` before the classifier (fcmf_multimodal.py:49)."""
    return _ClassifierCE.apply(pooled, wc, bc, labels, row_scale, drop)


# ------------------------------------------------------------------------------------------------- vocabulary projection
def _pad8(n: int) -> int:
    """Padded width of a projection: a multiple of 8 (16-byte rows), of 256 when wide (whole CTA-pair tiles)."""
    return (n + 7) // 8 * 8 if n <= 1024 else (n + 255) // 256 * 256


class _VocabLinear(Function):
    """logits = x @ W^T + b for an output width V that need not be a multiple of 8 (IAOG: V = 250 002): the weight is staged
    with zero rows up to Vp = ceil8(V), so forward, input-gradient and weight-gradient GEMMs all take the tcgen05 engine
    (N = Vp resp. K = Vp); the caller sees the [M, V] view of the [M, Vp] buffer."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], engine: int):
        V, H = weight.shape
        Vp = _pad8(V)
        w = ops.cast_matrix_padded(weight, x.dtype, Vp)
        b = None
        if bias is not None:
            b = torch.zeros((Vp,), dtype=torch.float32, device=x.device)
            b[:V].copy_(bias)
        y = ops.gemm_tn(x, w, b, EPI_NONE, engine=engine)                       # [M, Vp]; pad columns are exactly 0
        ctx.engine, ctx.has_bias, ctx.V, ctx.Vp = engine, bias is not None, V, Vp
        ctx.save_for_backward(x, weight)
        return y[:, :V]

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, weight = ctx.saved_tensors
        V, Vp, M = ctx.V, ctx.Vp, dy.shape[0]
        if (Vp > V and dy.stride(1) == 1 and dy.stride(0) == Vp and
                dy.untyped_storage().nbytes() >= (dy.storage_offset() + M * Vp) * dy.element_size()):
            dyp = dy.as_strided((M, Vp), (Vp, 1), dy.storage_offset())          # our own vocab-CE gradient: zero-padded storage
            dyp[:, V:].zero_()
        elif Vp == V:
            dyp = _c2(dy)
        else:
            dyp = torch.zeros((M, Vp), dtype=dy.dtype, device=dy.device)
            dyp[:, :V].copy_(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            wt = ops.cast_matrix_padded(weight, x.dtype, Vp, transpose=True)    # [H, Vp]
            if Vp >= 64 * x.shape[1] and x.dtype == torch.bfloat16 and ctx.engine != ENGINE_SIMT:
                dx = ops.cast_matrix(ops.gemm_tn_f32(dyp, wt), x.dtype)        # small output, long reduction: split-K
            else:
                dx = ops.gemm_tn(dyp, wt, None, EPI_NONE, engine=ctx.engine)
        dw, db = ops.gemm_wgrad(dyp, x, want_bias=ctx.has_bias, engine=ctx.engine)
        return dx, dw[:V], (db[:V] if db is not None else None), None


def vocab_linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], engine: int = ENGINE_AUTO) -> Tensor:
    """x [M, H] @ weight[V, H]^T + bias -> [M, V] (a view with row stride ceil8(V)); IAOGDecoder.dense (mm_modeling.py:645, 662)."""
    return _VocabLinear.apply(x, weight, bias, engine)


# ------------------------------------------------------------------------------------------------- vocabulary loss
class _VocabCE(Function):
    @staticmethod
    def forward(ctx, logits: Tensor, labels: Tensor, ignore_index: int):
        labels = labels.to(torch.int64).contiguous()
        loss_rows, lse = ops.vocab_ce_fwd(logits, labels, ignore_index)
        counted = (labels != ignore_index).sum().clamp_min(1).to(torch.float32)     # stays on the device: no sync
        ctx.ignore_index = ignore_index
        ctx.save_for_backward(logits, labels, lse, counted)
        return loss_rows.sum() / counted

    @staticmethod
    def backward(ctx, dloss: Tensor):
        logits, labels, lse, counted = ctx.saved_tensors
        scale = (dloss.to(torch.float32) / counted).reshape(1).contiguous()
        return ops.vocab_ce_bwd(logits, labels, lse, scale, ctx.ignore_index), None, None


def vocab_cross_entropy(logits: Tensor, labels: Tensor, ignore_index: int = -100) -> Tensor:
    """mean_{counted rows} CE(logits[r], labels[r]) over a wide class axis -- nn.CrossEntropyLoss(ignore_index) of the
    IAOG pre-training loop (run_pretraining_fcmf.py:320-322). logits [..., V] (any leading dims), labels [...]."""
    V = logits.shape[-1]
    return _VocabCE.apply(logits.reshape(-1, V), labels.reshape(-1), ignore_index)
