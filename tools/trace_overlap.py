"""Where do the gradient all-reduces sit relative to backward at N > 1? (there is no nsys in this image)
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/trace_overlap.py [--steps 2]
Rank 0 records two train() steps of the bench's fusion workload (BASELINE config 2, B = 64 per GPU, bucketed NCCL all-reduce launched
from autograd hooks) with torch.profiler (CUPTI kernel records) and prints, per step: the step's GPU span, every NCCL kernel with its start
/ end relative to the step and how much of it ran while one of OUR kernels was running on the compute stream (= overlapped), and the
exposed tail (NCCL time after the last compute kernel). Output: a text summary; --trace PATH also keeps the chrome trace (gzip)."""
import argparse
import gzip
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import fcmf_b200 as pkg

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--trace", default="")
args = ap.parse_args()

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
synth = pkg.synth
ddp = importlib.import_module(pkg.__name__ + ".ddp")
dims = synth.FusionDims(batch=args.batch)
B, A = dims.batch, dims.aspects
model = pkg.FCMF(None, num_labels=dims.num_labels, num_imgs=dims.num_imgs, num_roi=dims.num_roi)
model.load_state_dict(synth.make_params(dims, seed=42), strict=True)
model = model.to(dev).train()
model.encoder.compute_dtype = torch.bfloat16
reducer = ddp.BucketedGradReducer(ddp.fusion_named_parameters(model)) if world > 1 else None
host = synth.make_batch(dims, seed=1234 + rank)
dt = torch.bfloat16
inp = {"seq": host["sequence_output"].reshape(B * A, dims.seq_len, dims.hidden).to(dt).to(dev), "vis": host["visual_embeds_att"].to(dt).to(dev),
       "roi": host["roi_embeds_att"].to(dt).to(dev), "coors": host["roi_coors"].to(dev),
       "mask": host["added_attention_mask"].reshape(B * A, -1).to(dev), "labels": host["labels"].to(dev)}


def step():
    if reducer is not None:
        reducer.zero_grad()
    else:
        for p in model.parameters():
            p.grad = None
    seq = inp["seq"].detach().requires_grad_(True)
    _, loss = model.fuse_all_aspects(seq, inp["vis"], inp["roi"], inp["coors"], inp["mask"], inp["labels"], aspects=A, rows="full")
    loss.backward()
    if reducer is not None:
        reducer.finish()


# which bucket becomes final when (GPU time of the compute stream at the moment its last gradient hook fires)
launch_log = []
if reducer is not None:
    _orig_launch = reducer._launch

    def _logged_launch(bi):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        launch_log.append((bi, ev))
        _orig_launch(bi)

    reducer._launch = _logged_launch

for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
from torch.profiler import profile, ProfilerActivity
marks = []
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launch_log.clear()
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        step()
        torch.cuda.synchronize()
        last_log = [(bi, ev0.elapsed_time(ev)) for bi, ev in launch_log]
if rank == 0:
    if reducer is not None:
        print("# buckets in the order they became final (last profiled step): index, first parameter, MB, compute-stream time when final")
        for bi, ms in last_log:
            print(f"   bucket {bi}: {reducer.buckets[bi][0][0]:60s} {reducer.flat[bi].numel() * 4 / 1e6:6.1f} MB  final at {ms:6.2f} ms")
    path = "/tmp/fcmf_trace_rank0.json"
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    kern = [e for e in ev if e.get("cat") == "kernel" and "ts" in e]
    kern.sort(key=lambda e: e["ts"])
    is_nccl = lambda e: "nccl" in e["name"].lower()
    ours = [e for e in kern if not is_nccl(e)]
    # every step launches the same kernels: split the compute kernels into args.steps equal runs
    per = len(ours) // args.steps
    steps = [ours[i * per:(i + 1) * per] for i in range(args.steps)]
    print(f"# N = {world}, rank 0, B = {args.batch} per GPU, train(), rows=full; times in ms relative to the first kernel of the step")
    for si, s in enumerate(steps):
        t0, t1 = s[0]["ts"], max(e["ts"] + e["dur"] for e in s)
        t_next = steps[si + 1][0]["ts"] if si + 1 < len(steps) else t1 + 5000
        nc = [e for e in kern if is_nccl(e) and t0 <= e["ts"] < t_next and not (e["ts"] > t1 and si + 1 < len(steps) and e["dur"] < 20 and e["ts"] + e["dur"] > t_next - 300)]
        print(f"step {si}: {len(s)} compute kernels, span {(t1 - t0) / 1e3:.2f} ms, {len(nc)} NCCL kernels")
        busy = [(e["ts"], e["ts"] + e["dur"]) for e in s]
        total_exposed = 0.0
        for e in nc:
            a, b = e["ts"], e["ts"] + e["dur"]
            ov = sum(max(0.0, min(b, y) - max(a, x)) for x, y in busy)
            exposed = max(0.0, b - max(a, t1))
            total_exposed += exposed
            print(f"   {e['name'][:48]:48s} start {(a - t0) / 1e3:7.2f}  end {(b - t0) / 1e3:7.2f}  dur {e['dur'] / 1e3:6.3f}  under compute {100 * ov / max(e['dur'], 1e-9):5.1f} %  after the last compute kernel {exposed / 1e3:.3f}")
        last_nc = max([e["ts"] + e["dur"] for e in nc], default=t1)
        print(f"   step end incl. NCCL: {(max(last_nc, t1) - t0) / 1e3:.2f} ms; NCCL time after the last compute kernel: {total_exposed / 1e3:.3f} ms")
    if args.trace:
        with open(path, "rb") as f, gzip.open(args.trace, "wb") as g:
            g.write(f.read())
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
