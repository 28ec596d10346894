"""Generate the golden fixtures under tests/golden by running the UNMODIFIED reference
(/root/reference/fcmf_framework) in this container. ORACLE-side tooling: test infrastructure.

    python oracle/make_golden.py            # all cases
    python oracle/make_golden.py --case base_small

The reference cannot travel to the GPU box, so its outputs are committed as small .npz files;
weights and inputs are NOT stored -- they are regenerated from seeds by the package's
``synth.py`` (numpy RandomState: a frozen stream). Per case we store: logits [B,A,C], the
summed loss, d(loss)/d(sequence_output) in full (or strided), and for every fusion parameter
its gradient norm plus a strided sample (full for tensors <= 4096 elements).

Model dimensions are module-level constants in the reference (mm_modeling.py:21-30), copied by
``import *`` into fcmf_pretraining/fcmf_multimodal, so the H=1024 case runs in a subprocess that
patches the constants before first import."""
from __future__ import annotations

import argparse
import importlib
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("FCMF_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")

synth = importlib.import_module("multimodal-aspect-category-sentiment-analysis_b200.synth")

CASES = {
    # name: (dims kwargs, mask kind, param seed, batch seed)
    "base_small": (dict(batch=2, aspects=2, seq_len=24, num_imgs=2, num_roi=3), "bernoulli", 42, 1234),
    "base_cfg1_b1": (dict(batch=1, aspects=2, seq_len=170, num_imgs=7, num_roi=4), "ones", 43, 1235),
    "base_roi7": (dict(batch=2, aspects=1, seq_len=40, num_imgs=3, num_roi=7), "bernoulli", 44, 1236),
    "large_small": (dict(batch=2, aspects=2, seq_len=32, num_imgs=2, num_roi=4, hidden=1024, heads=16, inter=4096),
                    "bernoulli", 45, 1237),
}
SAMPLE_STRIDE = 997


def sample(t: torch.Tensor) -> np.ndarray:
    f = t.detach().reshape(-1)
    return (f if f.numel() <= 4096 else f[::SAMPLE_STRIDE]).numpy().copy()


def run_case(name: str) -> None:
    kw, mask_kind, pseed, bseed = CASES[name]
    dims = synth.FusionDims(**kw)
    sys.path.insert(0, REF)
    import fcmf_framework.mm_modeling as mm
    if dims.hidden != mm.HIDDEN_SIZE:
        mm.HIDDEN_SIZE, mm.NUM_ATTENTION_HEADS, mm.INTERMEDIATE_SIZE = dims.hidden, dims.heads, dims.inter
    from fcmf_framework.fcmf_multimodal import FCMF          # noqa: E402  (after the patch)
    from transformers import XLMRobertaConfig, XLMRobertaModel

    with tempfile.TemporaryDirectory() as d:                  # tiny local text encoder, replaced below
        XLMRobertaModel(XLMRobertaConfig(vocab_size=64, hidden_size=32, num_hidden_layers=1, num_attention_heads=2,
                                         intermediate_size=64, max_position_embeddings=40, type_vocab_size=1,
                                         pad_token_id=1)).save_pretrained(d)
        model = FCMF(d, num_labels=dims.num_labels, num_imgs=dims.num_imgs, num_roi=dims.num_roi).eval()

    class StubText(torch.nn.Module):                          # stands for FeatureExtractor (mm_modeling.py:433-446)
        def forward(self, input_ids, token_type_ids, attention_mask):
            return input_ids, None, None
    model.encoder.bert = StubText()

    params = synth.make_params(dims, seed=pseed)
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("encoder.bert") for k in missing), missing
    batch = synth.make_batch(dims, seed=bseed, mask=mask_kind)
    seq = batch["sequence_output"].clone().requires_grad_(True)
    crit = torch.nn.CrossEntropyLoss()                        # run_multimodal_fcmf.py:290

    total, logits_all = 0, []
    for a in range(dims.aspects):                             # run_multimodal_fcmf.py:464-475
        logits = model(input_ids=seq[:, a], token_type_ids=None, attention_mask=None,
                       added_attention_mask=batch["added_attention_mask"][:, a],
                       visual_embeds_att=batch["visual_embeds_att"], roi_embeds_att=batch["roi_embeds_att"],
                       roi_coors=batch["roi_coors"])
        total = total + crit(logits, batch["labels"][:, a])
        logits_all.append(logits)
    total.backward()

    out = {"logits": torch.stack(logits_all, 1).detach().numpy(), "loss": np.float64(total.item()),
           "d_sequence_output": seq.grad.numpy() if seq.grad.numel() <= 400_000 else sample(seq.grad),
           "dims": np.array(repr(dims.to_dict())), "mask_kind": np.array(mask_kind),
           "param_seed": np.int64(pseed), "batch_seed": np.int64(bseed), "sample_stride": np.int64(SAMPLE_STRIDE),
           "torch_version": np.array(torch.__version__)}
    for k, v in model.named_parameters():
        if k.startswith("encoder.bert"):
            continue
        assert v.grad is not None, k
        out["gnorm/" + k] = np.float64(v.grad.double().norm().item())
        out["gsample/" + k] = sample(v.grad)
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: loss={total.item():.6f} -> {path} ({os.path.getsize(path)/1024:.0f} KiB)")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    a = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    if a.case:
        run_case(a.case)
        return
    for name in CASES:                                        # one process per case: constants are import-time
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--case", name])


if __name__ == "__main__":
    main()
