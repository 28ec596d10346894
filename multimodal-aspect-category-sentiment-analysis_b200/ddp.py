"""Data-parallel gradient reduction for the fusion parameters: bucketed NCCL all-reduce on a side stream, launched
from autograd hooks in gradient-ready order so it overlaps the rest of backward.

Samples are independent in the fusion path (no cross-sample op), so the batch is sharded by sample across ranks with
no data-path collective; the only exchange is this all-reduce (DDP semantics: mean over ranks). The reference gets
the same thing implicitly from torch DDP (run_multimodal_fcmf.py:238-240); the text encoder's 1.1 GB of gradients
stays on stock DDP and is not touched here.

Gradients live directly inside the flat bucket buffers (param.grad is a view), so no pack/unpack copies are made.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist

# Grad-ready order of the fusion path's backward.
# Order in which backward finishes the gradients (fusion.fused_forward issues the hoisted projections right before their consumers):
# classifier -> fusion layer + text+ROI branch + its hoisted Q/K/V (all mm_attention) -> ROI side (box head, roimap2text) ->
# text->image FFN (~3 ms before the end) -> text->image attention projections and vismap2text (the very end).
BUCKET_ORDER: Sequence[Tuple[str, ...]] = (
    ("classifier.", "text_pooler."),
    ("encoder.text2roi_pooler.", "encoder.mm_attention."),
    ("encoder.box_head.", "encoder.roimap2text."),
    ("encoder.text2img_pooler.", "encoder.text2img_attention.layer.0.output.", "encoder.text2img_attention.layer.0.intermediate."),
    ("encoder.text2img_attention.", "encoder.vismap2text."),
)


# IAOG pre-training (FCMFSeq2Seq, run_pretraining_fcmf.py:301-337): the decoder's gradients are final first (it runs last in
# forward), upper blocks before lower ones; the tied embedding / vocabulary projection (768 MB in fp32) becomes final at the
# very end of the decoder's backward and overlaps the whole fusion backward.
SEQ2SEQ_BUCKET_ORDER: Sequence[Tuple[str, ...]] = (
    tuple(f"decoder.blks.block{i}." for i in range(11, 5, -1)),
    tuple(f"decoder.blks.block{i}." for i in range(5, -1, -1)),
    ("decoder.",),
) + tuple(BUCKET_ORDER[1:])


def fusion_named_parameters(model: torch.nn.Module):
    return [(n, p) for n, p in model.named_parameters() if not n.startswith("encoder.bert.") and p.requires_grad]


class BucketedGradReducer:
    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], process_group=None,
                 bucket_order: Sequence[Tuple[str, ...]] = BUCKET_ORDER, average: bool = True):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        named = list(named_params)
        buckets: List[List[Tuple[str, torch.nn.Parameter]]] = [[] for _ in bucket_order]
        rest: List[Tuple[str, torch.nn.Parameter]] = []
        for n, p in named:
            for b, prefixes in enumerate(bucket_order):
                if n.startswith(prefixes):
                    buckets[b].append((n, p))
                    break
            else:
                rest.append((n, p))
        if rest:
            buckets.append(rest)
        self.buckets = [b for b in buckets if b]
        self.flat: List[torch.Tensor] = []
        self._bucket_of: Dict[int, int] = {}
        self._pending: List[int] = []
        self._handles = []
        dev = named[0][1].device
        self.side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        for bi, bucket in enumerate(self.buckets):
            n = sum(p.numel() for _, p in bucket)
            flat = torch.zeros(n, dtype=torch.float32, device=dev)
            off = 0
            for _, p in bucket:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self._bucket_of[id(p)] = bi
                p.register_post_accumulate_grad_hook(self._on_grad_ready)
            self.flat.append(flat)
        self._sync = True
        self._reset_counts()

    # ---- per step ---------------------------------------------------------------------------------------------
    def _reset_counts(self):
        self._pending = [len(b) for b in self.buckets]
        self._handles = []

    def zero_grad(self):
        """Zero the flat buffers in place (param.grad views stay attached)."""
        for f in self.flat:
            f.zero_()
        self._reset_counts()

    @contextmanager
    def no_sync(self):
        """Gradient accumulation (the reference's --gradient_accumulation_steps, run_multimodal_fcmf.py:478-483): backward
        passes inside this context only accumulate into the flat buckets; the all-reduce runs with the first backward
        outside it (the last micro-batch), exactly as torch DDP's no_sync()."""
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev

    def _on_grad_ready(self, p: torch.nn.Parameter):
        bi = self._bucket_of[id(p)]
        flat, g = self.flat[bi], p.grad
        lo = flat.data_ptr()
        if g is None or not (lo <= g.data_ptr() < lo + flat.numel() * flat.element_size()):
            raise RuntimeError(
                "BucketedGradReducer: a parameter's .grad no longer lives in its flat bucket (optimizer.zero_grad() / "
                "model.zero_grad() default to set_to_none=True, which detaches it): zero gradients with reducer.zero_grad() "
                "or zero_grad(set_to_none=False)")
        if not self._sync:
            return                                                    # accumulation micro-batch: no collective
        self._pending[bi] -= 1
        if self._pending[bi] < 0:
            raise RuntimeError(
                "BucketedGradReducer: a second backward() reached a bucket that was already reduced this step; call "
                "reducer.zero_grad() once per optimizer step and run accumulation micro-batches under reducer.no_sync()")
        if self._pending[bi] == 0 and self.world > 1:
            self._launch(bi)

    def _launch(self, bi: int):
        flat = self.flat[bi]
        op = dist.ReduceOp.AVG if (self.average and flat.is_cuda) else dist.ReduceOp.SUM
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())        # gradients of this bucket are final
            with torch.cuda.stream(self.side):
                self._handles.append(dist.all_reduce(flat, op=op, group=self.group, async_op=True))
        else:                                                         # gloo / CPU (tests)
            dist.all_reduce(flat, op=op, group=self.group)
            if self.average:
                flat.div_(self.world)

    def finish(self):
        """Call after backward(): the compute stream waits for the outstanding all-reduces."""
        if self.world > 1:
            for bi, left in enumerate(self._pending):
                if left > 0:          # a parameter received no gradient this step (unused): reduce what we have
                    self._launch(bi)
                    self._pending[bi] = 0
            for h in self._handles:
                h.wait()
            if self.side is not None:
                torch.cuda.current_stream().wait_stream(self.side)
        self._handles = []

    def message_bytes(self) -> int:
        return sum(f.numel() * f.element_size() for f in self.flat)
