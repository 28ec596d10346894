"""World-size-2 gloo test of the bucketed gradient reducer (host logic; CPU tensors, no kernels)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
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
    q.put((rank, {n: p.grad.clone() for n, p in named}))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_reducer_averages_over_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
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
