"""BASELINE.json configs[4] sweep: XLM-R-large dims (H=1024, 16 heads, I=4096), L=256, per-GPU batch 16..256, on the GPUs this
process group has. Runs bench.py once per batch size (sub-process; torchrun for N > 1) and writes one JSON list.

    python tools/sweep_config5.py --gpus 1 --out gpurun_out/config5_n1.json [--batches 16,32,64,128,256]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--batches", default="16,32,64,128,256")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config5.json"))
ap.add_argument("--port", type=int, default=29511)
a = ap.parse_args()
rows = []
for b in [int(x) for x in a.batches.split(",")]:
    args = ["bench.py", "--gpus", str(a.gpus), "--steps", str(a.steps), "--warmup", "3", "--hidden", "1024", "--heads", "16",
            "--inter", "4096", "--seq-len", "256", "--batch", str(b), "--no-cpu-baseline", "--no-second-mode", "--max-seconds", "400"]
    cmd = ([sys.executable] if a.gpus == 1 else
           [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
            "--master-port", str(a.port)]) + args
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode == 0 and line:
        d = json.loads(line[-1])
        rows.append({"per_gpu_batch": b, "n_gpus": a.gpus, "samples_per_s": d["value"], "ms_per_step": d["ms_per_step"],
                     "e2e_samples_per_s": d["e2e"]["value"], "gemm_tflops": d["roofline"]["achieved"], "gemm_frac": d["roofline"]["frac"],
                     "mode": d["config"]["mode"], "clocks": d["clocks"]})
    else:
        tail = (r.stderr or r.stdout)[-400:]
        rows.append({"per_gpu_batch": b, "n_gpus": a.gpus, "failed": "out of memory" if "out of memory" in tail.lower() else tail})
    print(rows[-1], flush=True)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump({"config": "BASELINE.json configs[4]: H=1024, heads=16, I=4096, L=256, 7 images x 49 patches, 4 ROIs, A=6 folded, bf16, train()",
           "rows": rows}, open(a.out, "w"), indent=1)
