"""Shared helpers for the test-suite (test infrastructure; may import oracle/)."""
import ast
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
PKG_NAME = "multimodal-aspect-category-sentiment-analysis_b200"


def pkg(sub: str = ""):
    return importlib.import_module(PKG_NAME + (("." + sub) if sub else ""))


synth = pkg("synth")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    dims = synth.FusionDims(**ast.literal_eval(str(z["dims"])))
    return z, dims


def golden_inputs(z, dims):
    params = synth.make_params(dims, seed=int(z["param_seed"]))
    batch = synth.make_batch(dims, seed=int(z["batch_seed"]), mask=str(z["mask_kind"]))
    return params, batch


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  -- the 'relative' of north_star's 1e-4 / 2e-2 bars: error relative to
    the tensor's own scale (a per-element ratio is undefined at the zeros gradients contain)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def golden_sample(t: torch.Tensor, stride: int) -> torch.Tensor:
    f = t.detach().reshape(-1).cpu()
    return f if f.numel() <= 4096 else f[::stride]


def grad_scale(z) -> float:
    """Largest |gradient| over all fusion parameters of a golden case: the floor for relative errors of gradients
    that are identically zero in exact arithmetic (every attention KEY bias: adding a constant to all keys shifts a
    softmax row uniformly), where both sides hold only rounding noise."""
    return max(float(abs(z[k]).max()) for k in z.files if k.startswith("gsample/"))


def rel_err_floor(a, b, floor: float) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), floor))
