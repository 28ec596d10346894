"""The optimizer tail of the reference's training step as three
kernel launches: ``clip_grad_norm_(model.parameters(), 1.0)`` + ``torch.optim.AdamW.step()``
over the 4 parameter groups (run_multimodal_fcmf.py:249-289, 483-489), with the gradient norm and the clip coefficient kept
on the device (no ``.item()`` synchronisation per step).

    opt = FusedAdamW(optimizer_grouped_parameters, lr=..., max_grad_norm=1.0)     # same group dicts as torch.optim.AdamW
    loss.backward(); opt.step(); opt.zero_grad()

``param_groups`` / ``state_dict()`` follow torch.optim.AdamW (exp_avg, exp_avg_sq, step), so LR schedulers and the
reference's checkpoint code (run_multimodal_fcmf.py:327-333) keep working. fp32 CUDA parameters only."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import numpy as np
import torch

from . import _lib
from ._lib import OptTensor

CHUNK = 8192          # csrc/optim.cu: OPT_CHUNK


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: float = 0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = float(max_grad_norm)
        self._plist: List[torch.nn.Parameter] = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
        if not self._plist:
            raise ValueError("FusedAdamW: no parameters")
        dev = self._plist[0].device
        for p in self._plist:
            if p.dtype != torch.float32 or not p.is_cuda or p.device != dev or not p.is_contiguous():
                raise TypeError("FusedAdamW takes contiguous fp32 CUDA parameters on one device")
            st = self.state[p]
            st["step"] = torch.zeros((), dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p)
            st["exp_avg_sq"] = torch.zeros_like(p)
        blk_t, blk_c = [], []
        for i, p in enumerate(self._plist):
            for c in range((p.numel() + CHUNK - 1) // CHUNK):
                blk_t.append(i)
                blk_c.append(c)
        self._n_blocks = len(blk_t)
        self._blk_t = torch.tensor(blk_t, dtype=torch.int32, device=dev)
        self._blk_c = torch.tensor(blk_c, dtype=torch.int32, device=dev)
        self._host = (OptTensor * len(self._plist))()
        self._host_t = torch.from_numpy(np.frombuffer(self._host, dtype=np.uint8)).pin_memory()   # pinned staging copy
        self._table = torch.empty(C.sizeof(self._host), dtype=torch.uint8, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._coef = torch.ones(1, dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)       # last step's total gradient norm (device)
        self._steps = 0

    def _fill_table(self) -> int:
        """(Re)write the pointer table: gradients may be new tensors each step, lr may have been changed by a scheduler."""
        n = 0
        for g in self.param_groups:
            for p in g["params"]:
                if not p.requires_grad:
                    continue
                e = self._host[n]
                grad = p.grad
                if grad is None:                       # parameter without a gradient this step: skip it (n = 0 elements)
                    e.p, e.g, e.m, e.v, e.n = p.data_ptr(), p.data_ptr(), p.data_ptr(), p.data_ptr(), 0
                else:
                    if grad.dtype != torch.float32 or not grad.is_contiguous():
                        raise TypeError("FusedAdamW: gradients have to be contiguous fp32")
                    st = self.state[p]
                    e.p, e.g, e.m, e.v, e.n = p.data_ptr(), grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()
                e.lr, e.wd = float(g["lr"]), float(g["weight_decay"])
                n += 1
        self._host_t.copy_(torch.from_numpy(np.frombuffer(self._host, dtype=np.uint8)))
        self._table.copy_(self._host_t, non_blocking=True)
        return n

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._fill_table()
        stream = torch.cuda.current_stream().cuda_stream
        b1, b2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        self._steps += 1
        coef = None
        if self.max_grad_norm > 0.0:
            _lib.call("fcmf_opt_sumsq", self._table.data_ptr(), self._blk_t.data_ptr(), self._blk_c.data_ptr(), self._n_blocks,
                      self._sumsq.data_ptr(), stream)
            _lib.call("fcmf_opt_clip_coef", self._sumsq.data_ptr(), self.max_grad_norm, self._coef.data_ptr(),
                      self.grad_norm.data_ptr(), stream)
            coef = self._coef.data_ptr()
        _lib.call("fcmf_opt_adamw", self._table.data_ptr(), self._blk_t.data_ptr(), self._blk_c.data_ptr(), self._n_blocks,
                  coef, float(b1), float(b2), float(eps), self._steps, 1 if coef else 0, stream)
        for p in self._plist:
            self.state[p]["step"] += 1
        return loss
