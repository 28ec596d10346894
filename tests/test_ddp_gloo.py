"""World-size-2 gloo test of the bucketed gradient reducer (host logic; CPU tensors, no kernels)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import pkg


def _free_port():
    """A rendezvous FILE (no TCP port race between the probe and the workers' bind)."""
    import tempfile
    fd, path = tempfile.mkstemp(prefix="fcmf_gloo_")
    os.close(fd)
    os.unlink(path)
    return path


def _init(rank, world, port):
    dist.init_process_group("gloo", init_method=f"file://{port}", rank=rank, world_size=world)


def _worker(rank, world, port, q):
    _init(rank, world, port)
    ddp = pkg("ddp")
    torch.manual_seed(0)
    model = torch.nn.ModuleDict({"classifier": torch.nn.Linear(4, 3), "text_pooler": torch.nn.Linear(4, 4),
                                 "other": torch.nn.Linear(4, 2)})
    named = [(n, p) for n, p in model.named_parameters()]
    red = ddp.BucketedGradReducer(named)
    assert len(red.buckets) == 2                       # ("classifier."+"text_pooler.") and the rest
    x = torch.full((2, 4), float(rank + 1))
    for step in range(2):
        red.zero_grad()
        loss = model["classifier"](model["text_pooler"](x)).sum() + model["other"](x).sum() * (step + 1)
        loss.backward()
        red.finish()
    q.put((rank, {n: p.grad.detach().numpy().copy() for n, p in named}))       # by value: the worker may exit before the parent reads
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_reducer_averages_over_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {r: {n: torch.from_numpy(v) for n, v in g.items()} for r, g in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process expectation: mean over the two ranks' inputs
    torch.manual_seed(0)
    model = torch.nn.ModuleDict({"classifier": torch.nn.Linear(4, 3), "text_pooler": torch.nn.Linear(4, 4),
                                 "other": torch.nn.Linear(4, 2)})
    grads = []
    for rank in range(world):
        model.zero_grad()
        x = torch.full((2, 4), float(rank + 1))
        (model["classifier"](model["text_pooler"](x)).sum() + model["other"](x).sum() * 2).backward()
        grads.append({n: p.grad.clone() for n, p in model.named_parameters()})
    for n in grads[0]:
        want = (grads[0][n] + grads[1][n]) / 2
        assert torch.allclose(got[0][n], want, atol=1e-6) and torch.allclose(got[1][n], want, atol=1e-6), n


# ------------------------------------------------------------------------------------------------------------------
# The folded FCMF step sharded by sample over 2 ranks (kernel Functions replaced by the CPU stand-ins): the reducer's
# averaged gradients equal the single-process gradients on the concatenated batch (tier 5).
def _fcmf_setup():
    from _util import synth
    dims = synth.FusionDims(batch=4, aspects=2, seq_len=8, num_imgs=2, num_roi=3)
    model = pkg().FCMF(None, num_labels=dims.num_labels, num_imgs=dims.num_imgs, num_roi=dims.num_roi)
    model.load_state_dict(synth.make_params(dims, seed=11), strict=True)
    model.eval()
    model.encoder.compute_dtype = torch.float32
    return dims, model, synth.make_batch(dims, seed=12, mask="bernoulli")


def _fcmf_step(model, batch, sl, per, A):
    _, loss = model.fuse_all_aspects(batch["sequence_output"][sl], batch["visual_embeds_att"][sl], batch["roi_embeds_att"][sl],
                                     batch["roi_coors"][sl], batch["added_attention_mask"][sl].reshape(per * A, -1),
                                     batch["labels"][sl], rows="full")
    loss.backward()


def _fcmf_worker(rank, world, port, want_path, q):
    _init(rank, world, port)
    import _standins
    _standins.install_plain(pkg)
    ddp = pkg("ddp")
    dims, model, batch = _fcmf_setup()
    red = ddp.BucketedGradReducer(ddp.fusion_named_parameters(model))
    per = dims.batch // world
    sl = slice(rank * per, (rank + 1) * per)                 # a sample's aspects stay on one rank
    for _ in range(2):                                       # two steps: the flat buckets are re-zeroed in place
        red.zero_grad()
        _fcmf_step(model, batch, sl, per, dims.aspects)
        red.finish()
    want = torch.load(want_path)
    gmax = max(float(g.abs().max()) for g in want.values())
    worst = max(float((p.grad - want[n]).abs().max()) for n, p in model.named_parameters())
    q.put((rank, worst / gmax))                              # gradients compared in the worker: only a float travels back
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_fcmf_step_gradients_equal_single_process(tmp_path):
    import _standins
    st = _standins.install_plain(pkg)
    try:
        dims, model, batch = _fcmf_setup()
        _fcmf_step(model, batch, slice(0, dims.batch), dims.batch, dims.aspects)      # single process, concatenated batch
        want_path = str(tmp_path / "want.pt")
        torch.save({n: p.grad.clone() for n, p in model.named_parameters()}, want_path)
    finally:
        st.undo()
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fcmf_worker, args=(r, world, port, want_path, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(got) == [0, 1] and max(got.values()) < 1e-5, got


def test_gradient_buckets_become_final_branch_by_branch():
    """BUCKET_ORDER is the order in which backward FINISHES the gradients, which the forward's issue order decides (autograd
    runs nodes in reverse creation order): every parameter of a bucket must be final before any parameter of a later bucket's
    LAST gradient -- in particular mm_attention and the ROI side before the text->image branch's backward, not at the end
    (measured at N = 8: 0.85 ms of all-reduce after the last kernel when the hoisted projections were all issued up front)."""
    import _standins
    st = _standins.install_plain(pkg)
    try:
        ddp = pkg("ddp")
        dims, model, batch = _fcmf_setup()
        named = ddp.fusion_named_parameters(model)
        order = []
        handles = [p.register_post_accumulate_grad_hook(lambda _p, n=n: order.append(n)) for n, p in named]
        _fcmf_step(model, batch, slice(0, dims.batch), dims.batch, dims.aspects)
        for h in handles:
            h.remove()
    finally:
        st.undo()
    assert sorted(order) == sorted(n for n, _ in named)            # every fusion parameter got exactly one gradient

    def bucket(n):
        return next(b for b, pre in enumerate(ddp.BUCKET_ORDER) if n.startswith(pre))

    final_at = {}
    for t, n in enumerate(order):
        final_at[bucket(n)] = t                                     # position of the bucket's LAST gradient
    assert [b for b, _ in sorted(final_at.items(), key=lambda kv: kv[1])] == list(range(len(ddp.BUCKET_ORDER))), final_at
    first_t2i = min(t for t, n in enumerate(order) if n.startswith("encoder.text2img_attention."))
    late = [n for t, n in enumerate(order) if t > first_t2i and n.startswith(("encoder.mm_attention.", "encoder.box_head.", "encoder.roimap2text."))]
    assert not late, late                                           # nothing of the text+ROI side is still open during the text->image backward


# ------------------------------------------------------------------------------------------------------------------
# Misuse is loud, accumulation is supported: set_to_none zeroing detaches the buckets; a second backward
# without no_sync() would reduce stale data.
def _toy():
    torch.manual_seed(0)
    model = torch.nn.ModuleDict({"classifier": torch.nn.Linear(4, 3), "other": torch.nn.Linear(4, 2)})
    return model, [(n, p) for n, p in model.named_parameters()]


def test_reducer_raises_when_grads_are_detached_or_backward_runs_twice():
    import pytest
    ddp = pkg("ddp")
    model, named = _toy()
    red = ddp.BucketedGradReducer(named)
    x = torch.ones(2, 4)
    red.zero_grad()
    (model["classifier"](x).sum() + model["other"](x).sum()).backward()
    red.finish()
    with pytest.raises(RuntimeError, match="already reduced"):
        (model["classifier"](x).sum() + model["other"](x).sum()).backward()         # no zero_grad(), no no_sync()
    model.zero_grad()                                                               # set_to_none=True: detaches .grad
    red.zero_grad()
    with pytest.raises(RuntimeError, match="no longer lives in its flat bucket"):
        (model["classifier"](x).sum() + model["other"](x).sum()).backward()


def _accum_worker(rank, world, port, q):
    _init(rank, world, port)
    ddp = pkg("ddp")
    model, named = _toy()
    red = ddp.BucketedGradReducer(named)
    xs = [torch.full((2, 4), float(rank + 1 + 10 * k)) for k in range(3)]
    red.zero_grad()
    with red.no_sync():
        for x in xs[:-1]:                                                           # accumulation micro-batches
            (model["classifier"](x).sum() + model["other"](x).sum()).backward()
    (model["classifier"](xs[-1]).sum() + model["other"](xs[-1]).sum()).backward()   # last micro-batch: reduces the sum
    red.finish()
    q.put((rank, {n: p.grad.detach().numpy().copy() for n, p in named}))       # by value: the worker may exit before the parent reads
    dist.barrier()
    dist.destroy_process_group()


def test_reducer_gradient_accumulation_under_no_sync():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_accum_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {r: {n: torch.from_numpy(v) for n, v in g.items()} for r, g in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    model, named = _toy()
    want = {n: torch.zeros_like(p) for n, p in named}
    for rank in range(world):
        model.zero_grad()
        for k in range(3):
            x = torch.full((2, 4), float(rank + 1 + 10 * k))
            (model["classifier"](x).sum() + model["other"](x).sum()).backward()
        for n, p in named:
            want[n] += p.grad / world
    for n in want:
        assert torch.allclose(got[0][n], want[n], atol=1e-5) and torch.allclose(got[1][n], want[n], atol=1e-5), n
