"""bench.py contract checks that need no GPU: the reference arm's JSON line (rank 0 only under torchrun, other ranks exit 0
silently), and the kernel arm's loud refusal to run without CUDA (there is no CPU fallback for the fusion path)."""
import json
import os
import subprocess
import sys

import torch

from _util import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--seq-len", "24"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fcmf_fusion_fwd_bwd_samples_per_sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    # "reference" when the oracle/_ref snapshot of the reference's own modules is present (oracle/build_ref.py), else the port
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fcmf_framework", "fcmf_multimodal.py"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["config"]["sample_batch"] == 4                      # fixed: comparable across runs and GPU counts
    assert "workload" in d["config"] and d["gpu_launches"] == 0
    # a non-zero rank of a torchrun launch does no work and prints nothing
    r1 = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--seq-len", "24"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_kernel_arm_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
