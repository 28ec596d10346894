"""Tensor-level wrappers over the C ABI: they only turn torch tensors into (pointer, leading dimension) pairs,
pick the current CUDA stream and allocate outputs. No arithmetic happens here."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (BF16, F32, ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05, EPI_DGELU, EPI_GELU, EPI_NONE, EPI_TANH,
                   AttnDesc, Seg)

Tensor = torch.Tensor
LN_EPS = 1e-12                     # mm_modeling.py:159

# bench.py sets this to a list to time every GEMM launch with CUDA events on the launching stream:
# entries are (kind, M, N, K, start_event, end_event).
GEMM_PROFILE = None


def _timed_call(kind, M, N, K, name, *args):
    prof = GEMM_PROFILE
    if prof is None:
        _lib.call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call(name, *args)
    e1.record()
    prof.append((kind, M, N, K, e0, e1))


def dtype_code(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"fcmf_b200 kernels take float32 or bfloat16 activations, got {t.dtype}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("fcmf_b200: the fusion path runs on CUDA tensors only (no CPU fallback)")


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _rows2d(t: Tensor) -> Tuple[int, int, int]:
    """(rows, cols, ld) of a 2-D row-major view with unit column stride."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError(f"expected a 2-D tensor with contiguous columns, got shape {tuple(t.shape)} strides {t.stride()}")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    return t.shape[0], t.shape[1], ld


# ---------------------------------------------------------------------------------------------- GEMM
def gemm_tn(a: Tensor, b: Tensor, bias: Optional[Tensor] = None, epi: int = EPI_NONE, aux: Optional[Tensor] = None,
            out: Optional[Tensor] = None, engine: int = ENGINE_AUTO, want_aux: bool = False):
    """out[M,N] = epi(a[M,K] @ b[N,K]^T + bias). Returns out, or (out, aux) when epi == GELU and want_aux."""
    _need_cuda(a, b, bias, aux, out)
    M, K, lda = _rows2d(a)
    N, Kb, ldb = _rows2d(b)
    if K != Kb:
        raise RuntimeError(f"gemm_tn: inner dimensions differ: a is {tuple(a.shape)}, b is {tuple(b.shape)}")
    if a.dtype != b.dtype:
        raise TypeError(f"gemm_tn: operand dtypes differ ({a.dtype} vs {b.dtype})")
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    if epi == EPI_GELU and want_aux and aux is None:
        aux = torch.empty((M, N), dtype=a.dtype, device=a.device)
    _, _, ldd = _rows2d(out)
    ldaux = _rows2d(aux)[2] if aux is not None else 0
    if bias is not None and (bias.dtype != torch.float32 or not bias.is_contiguous()):
        raise TypeError("gemm_tn: bias must be contiguous float32")
    _timed_call("tn", M, N, K, "fcmf_gemm_tn", _p(a), lda, _p(b), ldb, _p(bias), _p(out), ldd, _p(aux), ldaux, M, N, K,
                epi, dtype_code(a), engine, _stream())
    return (out, aux) if (epi == EPI_GELU and want_aux) else out


def gemm_wgrad(dy: Tensor, x: Tensor, want_bias: bool = True, engine: int = ENGINE_AUTO,
               dw: Optional[Tensor] = None, db: Optional[Tensor] = None, accumulate: bool = False):
    """dw[N,K] (+)= dy[M,N]^T @ x[M,K], db[N] (+)= dy.sum(0); fp32 outputs."""
    _need_cuda(dy, x)
    M, N, lddy = _rows2d(dy)
    Mx, K, ldx = _rows2d(x)
    if M != Mx or dy.dtype != x.dtype:
        raise RuntimeError(f"gemm_wgrad: dy {tuple(dy.shape)} {dy.dtype} vs x {tuple(x.shape)} {x.dtype}")
    if dw is None:
        dw = torch.empty((N, K), dtype=torch.float32, device=dy.device)
        accumulate = False
    if want_bias and db is None:
        db = torch.empty((N,), dtype=torch.float32, device=dy.device)
    _timed_call("wgrad", M, N, K, "fcmf_gemm_wgrad", _p(dy), lddy, _p(x), ldx, _p(dw), _p(db) if want_bias else None,
                M, N, K, 1 if accumulate else 0, dtype_code(dy), engine, _stream())
    return dw, (db if want_bias else None)


# ---------------------------------------------------------------------------------------------- row-wise
def ln_fwd(x: Tensor, res: Optional[Tensor], res_idx: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float = LN_EPS):
    _need_cuda(x, res, res_idx, gamma, beta)
    M, H = x.shape
    assert x.is_contiguous() and (res is None or (res.is_contiguous() and res.shape[1] == H and res.dtype == x.dtype))
    assert res_idx is None or (res_idx.dtype == torch.int32 and res_idx.numel() == M)
    y = torch.empty_like(x)
    mean = torch.empty((M,), dtype=torch.float32, device=x.device)
    rstd = torch.empty((M,), dtype=torch.float32, device=x.device)
    _lib.call("fcmf_ln_fwd", _p(x), _p(res), _p(res_idx), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), M, H,
              float(eps), dtype_code(x), _stream())
    return y, mean, rstd


def ln_bwd(dy: Tensor, dy_add: Optional[Tensor], x: Tensor, res: Optional[Tensor], res_idx: Optional[Tensor],
           gamma: Tensor, mean: Tensor, rstd: Tensor):
    """Returns (ds, dgamma, dbeta); dgamma/dbeta are fresh fp32 tensors."""
    _need_cuda(dy, x)
    M, H = x.shape
    assert dy.is_contiguous() and dy.shape == x.shape and dy.dtype == x.dtype
    assert dy_add is None or (dy_add.is_contiguous() and dy_add.shape == x.shape and dy_add.dtype == x.dtype)
    ds = torch.empty_like(x)
    dgb = torch.zeros((2, H), dtype=torch.float32, device=x.device)
    _lib.call("fcmf_ln_bwd", _p(dy), _p(dy_add), _p(x), _p(res), _p(res_idx), _p(gamma), _p(mean), _p(rstd), _p(ds),
              dgb[0].data_ptr(), dgb[1].data_ptr(), M, H, dtype_code(x), _stream())
    return ds, dgb[0], dgb[1]


def mask_additive(mask: Tensor, n: int) -> Tensor:
    """(1 - mask[:, :n]) * -10000 as fp32 [rows, n]."""
    _need_cuda(mask)
    if mask.dtype != torch.int64:
        mask = mask.to(torch.int64)
    if mask.dim() != 2 or mask.shape[1] < n:
        raise RuntimeError(f"added_attention_mask of shape {tuple(mask.shape)} is shorter than the {n} positions needed")
    if mask.stride(1) != 1:
        mask = mask.contiguous()
    out = torch.empty((mask.shape[0], n), dtype=torch.float32, device=mask.device)
    _lib.call("fcmf_mask_additive", _p(mask), mask.stride(0), _p(out), mask.shape[0], n, _stream())
    return out


def gather_sum_rows(src: Tensor, idx: Tensor, n_out: int, G: int, out: Optional[Tensor] = None, accumulate: bool = False):
    """out[o] = sum_g src[idx[o, g]] (negative indices skipped). src/out are 2-D row views."""
    _need_cuda(src, idx, out)
    _, width, ldsrc = _rows2d(src)
    assert idx.dtype == torch.int32 and idx.is_contiguous() and idx.numel() == n_out * G
    if out is None:
        out = torch.empty((n_out, width), dtype=src.dtype, device=src.device)
        accumulate = False
    _, wo, ldout = _rows2d(out)
    assert wo == width and out.dtype == src.dtype
    _lib.call("fcmf_gather_sum_rows", _p(src), ldsrc, _p(idx), _p(out), ldout, n_out, G, width, 1 if accumulate else 0,
              dtype_code(src), _stream())
    return out


def dtanh(dy: Tensor, y: Tensor) -> Tensor:
    _need_cuda(dy, y)
    assert dy.is_contiguous() and y.is_contiguous() and dy.shape == y.shape and dy.dtype == y.dtype
    out = torch.empty_like(dy)
    _lib.call("fcmf_dtanh", _p(dy), _p(y), _p(out), dy.numel(), dtype_code(dy), _stream())
    return out


def cast_matrix(w: Tensor, dtype: torch.dtype, transpose: bool = False) -> Tensor:
    """fp32 [rows, cols] parameter -> compute dtype, optionally transposed (weight staging)."""
    _need_cuda(w)
    assert w.dtype == torch.float32 and w.dim() == 2
    w = w.contiguous()
    if dtype == torch.float32 and not transpose:
        return w
    rows, cols = w.shape
    out = torch.empty((cols, rows) if transpose else (rows, cols), dtype=dtype, device=w.device)
    _lib.call("fcmf_cast_matrix", _p(w), _p(out), rows, cols, 1 if transpose else 0, BF16 if dtype == torch.bfloat16 else F32,
              _stream())
    return out


def cast_to_f32(t: Tensor) -> Tensor:
    _need_cuda(t)
    if t.dtype == torch.float32:
        return t
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    _lib.call("fcmf_cast_to_f32", _p(t), _p(out), t.numel(), dtype_code(t), _stream())
    return out


# ---------------------------------------------------------------------------------------------- attention
def set_attn_engine(engine: int) -> None:
    """ENGINE_AUTO (tcgen05 where supported), ENGINE_SIMT or ENGINE_TCGEN05 for subsequent attention launches."""
    _lib.call("fcmf_set_attn_engine", engine)


class SegSpec:
    """One query/key/value segment: `rows` rows per group taken from columns [col, col+heads*dh) of a 2-D tensor."""
    __slots__ = ("t", "col", "rows", "idx")

    def __init__(self, t: Tensor, col: int, rows: int, idx: Optional[Tensor]):
        self.t, self.col, self.rows, self.idx = t, col, rows, idx


def _fill_seg(dst: Seg, s: Optional[SegSpec], esize: int) -> None:
    if s is None:
        dst.ptr, dst.ld, dst.rows, dst.idx = None, 0, 0, None
        return
    _, _, ld = _rows2d(s.t)
    dst.ptr = s.t.data_ptr() + s.col * esize
    dst.ld = ld
    dst.rows = s.rows
    dst.idx = None if s.idx is None else s.idx.data_ptr()


def make_attn_desc(q: Sequence[Optional[SegSpec]], k: Sequence[Optional[SegSpec]], v: Sequence[Optional[SegSpec]],
                   NP: int, heads: int, dh: int, scale: float, mask_add: Optional[Tensor], mask_div: int,
                   bias: Optional[Tensor], causal: bool = False) -> AttnDesc:
    d = AttnDesc()
    esize = q[0].t.element_size()
    for i in range(2):
        _fill_seg(d.q[i], q[i] if i < len(q) else None, esize)
        _fill_seg(d.k[i], k[i] if i < len(k) else None, esize)
        _fill_seg(d.v[i], v[i] if i < len(v) else None, esize)
    d.mask_add = _p(mask_add)
    d.ld_mask = mask_add.stride(0) if mask_add is not None else 0
    d.mask_div = mask_div
    d.bias = _p(bias)
    d.NP, d.heads, d.dh, d.scale = NP, heads, dh, float(scale)
    d.causal = 1 if causal else 0
    return d


def attn_fwd(desc: AttnDesc, Lq: int, dtype: torch.dtype, device) -> Tuple[Tensor, Tensor]:
    HD = desc.heads * desc.dh
    ctx = torch.empty((desc.NP * Lq, HD), dtype=dtype, device=device)
    lse = torch.empty((desc.NP, desc.heads, Lq), dtype=torch.float32, device=device)
    _lib.call("fcmf_attn_fwd", C.byref(desc), _p(ctx), HD, _p(lse), BF16 if dtype == torch.bfloat16 else F32, _stream())
    return ctx, lse


def attn_bwd(desc: AttnDesc, Lq: int, Lk: int, ctx: Tensor, dctx: Tensor, lse: Tensor, want_dbias: bool):
    HD = desc.heads * desc.dh
    dev, dt = ctx.device, ctx.dtype
    assert dctx.shape == ctx.shape and dctx.dtype == dt and dctx.stride(1) == 1
    dq = torch.empty((desc.NP * Lq, HD), dtype=dt, device=dev)
    dk = torch.empty((desc.NP * Lk, HD), dtype=dt, device=dev)
    dv = torch.empty((desc.NP * Lk, HD), dtype=dt, device=dev)
    delta = torch.empty((desc.NP, desc.heads, Lq), dtype=torch.float32, device=dev)
    dbias = torch.empty((desc.NP, desc.heads, Lq, Lk), dtype=torch.float32, device=dev) if want_dbias else None
    _lib.call("fcmf_attn_bwd", C.byref(desc), _p(ctx), HD, _p(dctx), dctx.stride(0), _p(lse), _p(delta), _p(dq), _p(dk),
              _p(dv), _p(dbias), BF16 if dt == torch.bfloat16 else F32, _stream())
    return dq, dk, dv, dbias


# ---------------------------------------------------------------------------------------------- geometry / head
_FREQ8 = None


def _freq8():
    """1 / 1000^(k/8), k = 0..7, built in float32 on the host with the same torch ops as roi_modeling.py:123-125."""
    global _FREQ8
    if _FREQ8 is None:
        f = torch.arange(64 / 8)
        f = 1.0 / torch.pow(1000, f / (64 / 8))
        _FREQ8 = (C.c_float * 8)(*[float(v) for v in f.tolist()])
    return _FREQ8


def box_geometry_fwd(boxes: Tensor, wg_w: Tensor, wg_b: Tensor, heads: int):
    """boxes f64 [G, NR, 4] -> (emb f32 [G,NR,NR,64], bias f32 [G,heads,NR,NR])."""
    _need_cuda(boxes, wg_w, wg_b)
    if boxes.dtype != torch.float64:
        boxes = boxes.to(torch.float64)
    boxes = boxes.contiguous()
    G, NR, _ = boxes.shape
    emb = torch.empty((G, NR, NR, 64), dtype=torch.float32, device=boxes.device)
    bias = torch.empty((G, heads, NR, NR), dtype=torch.float32, device=boxes.device)
    _lib.call("fcmf_box_geometry_fwd", _p(boxes), _p(wg_w), _p(wg_b), _freq8(), _p(emb), _p(bias), G, NR, heads, _stream())
    return emb, bias


def box_geometry_bwd(emb: Tensor, wg_w: Tensor, wg_b: Tensor, dbias: Tensor):
    G, NR = emb.shape[0], emb.shape[1]
    heads = wg_w.shape[0]
    dz = torch.empty_like(dbias)
    dw = torch.zeros_like(wg_w)
    db = torch.zeros_like(wg_b)
    _lib.call("fcmf_box_geometry_bwd", _p(emb), _p(wg_w), _p(wg_b), _p(dbias.contiguous()), _p(dz), _p(dw), _p(db), G, NR,
              heads, _stream())
    return dw, db


def cls_ce_fwd(pooled: Tensor, Wc: Tensor, bc: Tensor, labels: Optional[Tensor]):
    _need_cuda(pooled, Wc, bc, labels)
    R, H = pooled.shape
    Cn = Wc.shape[0]
    assert pooled.is_contiguous() and Wc.is_contiguous() and Wc.dtype == torch.float32
    logits = torch.empty((R, Cn), dtype=torch.float32, device=pooled.device)
    probs = torch.empty((R, Cn), dtype=torch.float32, device=pooled.device)
    loss_rows = torch.empty((R,), dtype=torch.float32, device=pooled.device) if labels is not None else None
    if labels is not None:
        labels = labels.to(torch.int64).contiguous()
    _lib.call("fcmf_cls_ce_fwd", _p(pooled), _p(Wc), _p(bc), _p(labels), _p(logits), _p(probs), _p(loss_rows), R, H, Cn,
              dtype_code(pooled), _stream())
    return logits, probs, loss_rows


def cls_ce_bwd(pooled: Tensor, Wc: Tensor, probs: Optional[Tensor], labels: Optional[Tensor],
               dlogits_in: Optional[Tensor], row_scale: float):
    R, H = pooled.shape
    Cn = Wc.shape[0]
    dev = pooled.device
    dlogits = torch.empty((R, Cn), dtype=torch.float32, device=dev)
    dpooled = torch.empty_like(pooled)
    dWc = torch.zeros_like(Wc)
    dbc = torch.zeros((Cn,), dtype=torch.float32, device=dev)
    if dlogits_in is not None:
        dlogits_in = dlogits_in.to(torch.float32).contiguous()
    _lib.call("fcmf_cls_ce_bwd", _p(pooled), _p(Wc), _p(probs), _p(labels), _p(dlogits_in), float(row_scale), _p(dlogits),
              _p(dpooled), _p(dWc), _p(dbc), R, H, Cn, dtype_code(pooled), _stream())
    return dpooled, dWc, dbc
