"""``torch.library`` registration of the C-ABI kernels: ``torch.ops.fcmf_b200.*`` (SURVEY.md section 7.3 / north_star: "a thin
C-ABI torch custom-op layer").

The drop-in modules call the kernels through plain ``torch.autograd.Function``s over ctypes (functional.py): that path has the
lowest per-call host overhead, which matters for the launch-bound shapes (live rows, the IAOG decoder). This module exposes
the same kernels as registered custom ops -- schema, fake (meta) implementations and autograd formulas -- so that
``torch.compile`` / ``torch.export`` / ``make_fx`` see opaque ops with known output shapes instead of ctypes calls, and so
that other code can call ``torch.ops.fcmf_b200.linear(x, w, b, "tanh")`` directly. Both surfaces execute the identical
library entry points; there is no second implementation.

    import fcmf_b200.torch_ops            # registers the ops (idempotent)
    y = torch.ops.fcmf_b200.linear(x, weight, bias, "none")
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops
from ._lib import EPI_DGELU, EPI_GELU, EPI_NONE, EPI_TANH

Tensor = torch.Tensor
NS = "fcmf_b200"
_EPI = {"none": EPI_NONE, "tanh": EPI_TANH}


@torch.library.custom_op(f"{NS}::gemm_tn", mutates_args=())
def gemm_tn(a: Tensor, b: Tensor, bias: Optional[Tensor], act: str) -> Tensor:
    """act(a[M,K] @ b[N,K]^T + bias)  -- fcmf_gemm_tn"""
    return ops.gemm_tn(a, b, bias, _EPI[act])


@gemm_tn.register_fake
def _(a, b, bias, act):
    return a.new_empty((a.shape[0], b.shape[0]))


@torch.library.custom_op(f"{NS}::gemm_wgrad", mutates_args=())
def gemm_wgrad(dy: Tensor, x: Tensor) -> Tuple[Tensor, Tensor]:
    """(dy^T @ x [N,K] fp32, column sums of dy [N] fp32)  -- fcmf_gemm_wgrad"""
    dw, db = ops.gemm_wgrad(dy, x, want_bias=True)
    return dw, db


@gemm_wgrad.register_fake
def _(dy, x):
    return (dy.new_empty((dy.shape[1], x.shape[1]), dtype=torch.float32), dy.new_empty((dy.shape[1],), dtype=torch.float32))


@torch.library.custom_op(f"{NS}::dtanh", mutates_args=())
def dtanh(dy: Tensor, y: Tensor) -> Tensor:
    return ops.dtanh(dy.contiguous(), y.contiguous())


@dtanh.register_fake
def _(dy, y):
    return torch.empty_like(dy)


@torch.library.custom_op(f"{NS}::cast_matrix", mutates_args=())
def cast_matrix(w: Tensor, to_bf16: bool, transpose: bool) -> Tensor:
    return ops.cast_matrix(w, torch.bfloat16 if to_bf16 else torch.float32, transpose=transpose)


@cast_matrix.register_fake
def _(w, to_bf16, transpose):
    shape = (w.shape[1], w.shape[0]) if transpose else tuple(w.shape)
    return w.new_empty(shape, dtype=torch.bfloat16 if to_bf16 else torch.float32)


@torch.library.custom_op(f"{NS}::linear", mutates_args=())
def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], act: str) -> Tensor:
    """act(x @ weight^T + bias): nn.Linear (+ BertPooler's tanh) with an fp32 master weight staged to x.dtype."""
    return ops.gemm_tn(x, ops.cast_matrix(weight, x.dtype), bias, _EPI[act])


@linear.register_fake
def _(x, weight, bias, act):
    return x.new_empty((x.shape[0], weight.shape[0]))


def _linear_setup(ctx, inputs, output):
    x, weight, bias, act = inputs
    ctx.act, ctx.has_bias = act, bias is not None
    ctx.save_for_backward(x, weight, output if act == "tanh" else None)


def _linear_backward(ctx, dy):
    x, weight, y = ctx.saved_tensors
    dy = dy.contiguous()
    if ctx.act == "tanh":
        dy = torch.ops.fcmf_b200.dtanh(dy, y)
    dx = torch.ops.fcmf_b200.gemm_tn(dy, torch.ops.fcmf_b200.cast_matrix(weight, x.dtype == torch.bfloat16, True), None, "none")
    dw, db = torch.ops.fcmf_b200.gemm_wgrad(dy, x)
    return dx, dw, (db if ctx.has_bias else None), None


linear.register_autograd(_linear_backward, setup_context=_linear_setup)


@torch.library.custom_op(f"{NS}::layer_norm_residual", mutates_args=())
def layer_norm_residual(x: Tensor, res: Tensor, gamma: Tensor, beta: Tensor, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    """TF-style LayerNorm(x + res) (mm_modeling.py:166-171, 276-280) -> (y, mean, rstd)  -- fcmf_ln_fwd"""
    return ops.ln_fwd(x.contiguous(), res.contiguous(), None, gamma, beta, eps)


@layer_norm_residual.register_fake
def _(x, res, gamma, beta, eps):
    return torch.empty_like(x), x.new_empty((x.shape[0],), dtype=torch.float32), x.new_empty((x.shape[0],), dtype=torch.float32)


@torch.library.custom_op(f"{NS}::layer_norm_residual_bwd", mutates_args=())
def layer_norm_residual_bwd(dy: Tensor, x: Tensor, res: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    ds, dg, db = ops.ln_bwd(dy.contiguous(), None, x.contiguous(), res.contiguous(), None, gamma, mean, rstd)
    return ds, dg.clone(), db.clone()            # dg / db are rows of one buffer: custom-op outputs must not alias each other


@layer_norm_residual_bwd.register_fake
def _(dy, x, res, gamma, mean, rstd):
    return torch.empty_like(x), torch.empty_like(gamma), torch.empty_like(gamma)


def _ln_setup(ctx, inputs, output):
    x, res, gamma, beta, eps = inputs
    ctx.save_for_backward(x, res, gamma, output[1], output[2])


def _ln_backward(ctx, dy, dmean, drstd):
    x, res, gamma, mean, rstd = ctx.saved_tensors
    ds, dg, db = torch.ops.fcmf_b200.layer_norm_residual_bwd(dy, x, res, gamma, mean, rstd)
    return ds, ds, dg, db, None


layer_norm_residual.register_autograd(_ln_backward, setup_context=_ln_setup)

REGISTERED = ("gemm_tn", "gemm_wgrad", "dtanh", "cast_matrix", "linear", "layer_norm_residual", "layer_norm_residual_bwd")
