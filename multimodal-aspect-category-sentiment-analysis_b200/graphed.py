"""CUDA-graph capture of one folded fusion step (forward + backward [+ gradient all-reduce]).

The fusion path is a fixed sequence of ~170 kernel launches per step with static shapes; in live-rows mode the GPU work
(a few ms) is shorter than the time Python needs to issue it. Capturing the whole step once and replaying it removes the
launch path from the critical path -- the B200-first alternative to a tracing compiler (no torch.compile involved: the
captured launches are this library's own kernels on the capturing stream).

Usage:
    step = GraphedFusionStep(model, example_inputs, aspects=6, rows="live", reducer=None)
    logits, loss = step(inputs)        # copies inputs into the static buffers, replays, returns static outputs
Gradients land in ``param.grad`` (static tensors, overwritten every replay) and in ``step.seq_grad``.

train() mode: the dropout seeds captured in the graph are immediates, so the capture also holds a device counter that
every dropout kernel adds to its seed (fcmf_dropout.seed_dev) and that the graph itself increments first thing in
each replay -- every replay draws fresh masks without re-capturing.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import functional as Fn

Tensor = torch.Tensor


class GraphedFusionStep:
    def __init__(self, model, inputs: Dict[str, Tensor], aspects: int, rows: str, reducer=None, warmup: int = 3):
        self.model, self.aspects, self.rows, self.reducer = model, aspects, rows, reducer
        self.static = {k: v.clone() for k, v in inputs.items()}
        self.static["seq"].requires_grad_(True)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=self.static["seq"].device) if model.training else None
        Fn.set_seed_device_tensor(self.seed_dev)
        try:
            self._build(warmup)
        finally:
            Fn.set_seed_device_tensor(None)

    def _build(self, warmup: int):
        model = self.model
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up on a side stream (allocator + lazy init settle)
            for _ in range(warmup):
                self._zero()
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.reducer is None:                            # static gradient buffers: capture accumulates in place
            for p in self.params:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
        self.static["seq"].grad = torch.zeros_like(self.static["seq"])
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            if self.seed_dev is not None:
                self.seed_dev.add_(1)                       # a new mask family per replay
            self._zero()
            self.logits, self.loss = self._run()
        self.seq_grad = self.static["seq"].grad

    def _zero(self):
        if self.reducer is not None:
            self.reducer.zero_grad()
        else:
            for p in self.params:
                if p.grad is not None:
                    p.grad.zero_()
        if self.static["seq"].grad is not None:
            self.static["seq"].grad.zero_()

    def _run(self):
        s = self.static
        logits, loss = self.model.fuse_all_aspects(s["seq"], s["vis"], s["roi"], s["coors"], s["mask"], s["labels"],
                                                   aspects=self.aspects, rows=self.rows)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        return logits, loss

    def __call__(self, inputs: Optional[Dict[str, Tensor]] = None):
        if inputs is not None:
            for k, v in inputs.items():
                if v is not self.static[k]:
                    self.static[k].detach().copy_(v, non_blocking=True)
        self.graph.replay()
        return self.logits, self.loss


class GraphedStep:
    """CUDA-graph capture of an arbitrary fixed-shape training step built on this library's Functions (e.g. the IAOG
    pre-training step: ~900 launches of mostly tiny decoder kernels, CPU-launch-bound when issued eagerly).

        step = GraphedStep(run, params, static_inputs, training=model.training, reducer=None)
        out = step(new_inputs)          # copies into the static buffers, replays, returns run()'s static outputs

    ``run(static_inputs)`` performs forward + backward (NOT the gradient zeroing, NOT reducer.finish()) and returns a tensor or
    tuple of tensors; gradients accumulate into static ``param.grad`` buffers that are zeroed in place inside the graph."""

    def __init__(self, run, params, static_inputs: Dict[str, Tensor], training: bool, reducer=None, warmup: int = 3):
        self.run, self.reducer = run, reducer
        self.params = [p for p in params if p.requires_grad]
        self.static = static_inputs
        dev = next(iter(static_inputs.values())).device
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev) if training else None
        Fn.set_seed_device_tensor(self.seed_dev)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._zero()
                    self._step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if reducer is None:
                for p in self.params:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                if self.seed_dev is not None:
                    self.seed_dev.add_(1)
                self._zero()
                self.out = self._step()
        finally:
            Fn.set_seed_device_tensor(None)

    def _zero(self):
        if self.reducer is not None:
            self.reducer.zero_grad()
        else:
            for p in self.params:
                if p.grad is not None:
                    p.grad.zero_()

    def _step(self):
        out = self.run(self.static)
        if self.reducer is not None:
            self.reducer.finish()
        return out

    def __call__(self, inputs: Optional[Dict[str, Tensor]] = None):
        if inputs is not None:
            for k, v in inputs.items():
                if v is not self.static[k]:
                    self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.out
