"""End-to-end ``FCMFSeq2Seq.forward``: encoder -> combined_mask / dec_state hand-off -> IAOG
decoder -> pre-training loss, against tests/golden/seq2seq_t32.npz = outputs of the UNMODIFIED reference wrapper
(fcmf_pretraining.py:168-207, run_pretraining_fcmf.py:320-324; generator: oracle/make_golden_seq2seq.py).
T = 32 (BASELINE config 4), V = 1003 (V % 8 != 0 like the real 250 002)."""
import os
import sys

import numpy as np
import pytest
import torch

from _util import GOLD, ROOT, golden_sample, pkg, rel_err, rel_err_floor
from oracle import fcmf_oracle as O

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import make_golden_seq2seq as G                                                # noqa: E402
import _standins                                                               # noqa: E402


def _golden():
    return np.load(os.path.join(GOLD, "seq2seq_t32.npz"))


def _loss(logits, labels):
    return torch.nn.functional.cross_entropy(logits.permute(0, 2, 1), labels, ignore_index=-100)


def test_oracle_seq2seq_matches_reference_golden():
    """Pins the oracle's restatement of the hand-off: fusion_encoder -> iaog_decoder, both attentions tril-masked."""
    z = _golden()
    params, dims = G.seq2seq_params()
    params = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    params["decoder.dense.weight"] = params["decoder.embedding.weight"]
    batch, dec_x, labels, _attn = G.inputs(dims)
    seq = batch["sequence_output"][:, 0].clone().requires_grad_(True)
    fused = O.fusion_encoder(seq, batch["visual_embeds_att"], batch["roi_embeds_att"], batch["roi_coors"],
                             batch["added_attention_mask"][:, 0], params, dims.heads, dims.num_imgs, dims.num_roi)
    logits = O.iaog_decoder(dec_x, fused, params, num_blocks=12)
    loss = _loss(logits, labels)
    loss.backward()
    assert rel_err(logits, torch.from_numpy(z["logits"])) < 5e-5
    assert abs(loss.item() - float(z["loss"])) < 5e-5 * abs(float(z["loss"]))
    assert rel_err(golden_sample(seq.grad, int(z["sample_stride"])), torch.from_numpy(z["d_sequence_output"])) < 2e-4


class _StubText(torch.nn.Module):
    def forward(self, input_ids, token_type_ids, attention_mask):
        return input_ids, None, None


def _build(device, dtype):
    params, dims = G.seq2seq_params()
    model = pkg().FCMFSeq2Seq(G.VOCAB, G.T, None, dims.num_imgs, dims.num_roi, 0.7)
    model.encoder.bert = _StubText()
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not unexpected and all("pos_encoding" in k for k in missing), (missing, unexpected)
    model = model.to(device).eval()
    model.encoder.compute_dtype = dtype
    model.decoder.compute_dtype = dtype
    return model, dims


def _run(model, dims, device):
    batch, dec_x, labels, attn = G.inputs(dims)
    seq = batch["sequence_output"][:, 0].to(device).clone().requires_grad_(True)
    logits = model(seq, dec_x.to(device), batch["visual_embeds_att"].to(device), batch["roi_embeds_att"].to(device),
                   batch["roi_coors"].to(device), None, attn.to(device), batch["added_attention_mask"][:, 0].to(device), None, True)
    return seq, logits, labels.to(device)


def test_shipped_seq2seq_wrapper_on_stand_ins(monkeypatch):
    """The SHIPPED FCMFSeq2Seq (combined_mask construction, dec_state, tied projection) on CPU stand-ins of the kernel
    Functions equals the reference's golden."""
    z = _golden()
    _standins.install(monkeypatch, pkg)
    model, dims = _build("cpu", None)
    seq, logits, labels = _run(model, dims, "cpu")
    loss = _loss(logits, labels)
    loss.backward()
    assert logits.shape == (dims.batch, G.T, G.VOCAB)
    assert rel_err(logits, torch.from_numpy(z["logits"])) < 1e-4
    assert abs(loss.item() - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    assert rel_err(golden_sample(seq.grad, int(z["sample_stride"])), torch.from_numpy(z["d_sequence_output"])) < 5e-4


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float32, 1e-4, 5e-4), (torch.bfloat16, 3e-2, 0.15)])
def test_kernel_seq2seq_matches_reference_golden(dtype, tol, gtol):
    z = _golden()
    model, dims = _build("cuda", dtype)
    if dtype == torch.float32:
        model.encoder.engine = pkg("_lib").ENGINE_SIMT
    seq, logits, labels = _run(model, dims, "cuda")
    loss = pkg().FCMFSeq2Seq.loss(logits, labels)
    loss.backward()
    assert rel_err(logits, torch.from_numpy(z["logits"])) < tol
    assert abs(loss.item() - float(z["loss"])) < (2e-4 if dtype == torch.float32 else 3e-2) * abs(float(z["loss"]))
    stride = int(z["sample_stride"])
    assert rel_err(golden_sample(seq.grad.float(), stride), torch.from_numpy(z["d_sequence_output"])) < gtol
    scale = max(float(abs(z[k]).max()) for k in z.files if k.startswith("gsample/"))
    checked = 0
    for k, v in model.named_parameters():
        key = "gsample/" + k
        if key not in z.files or v.grad is None or ".WGs." in k:
            continue
        zero_grad = k.endswith("key.bias") or k.endswith("box_head.linears.1.bias")
        floor = (1e-1 if zero_grad else 1e-3) * scale * (1.0 if dtype == torch.float32 else 10.0)
        assert rel_err_floor(golden_sample(v.grad, stride), torch.from_numpy(z[key]), floor) < gtol, k
        checked += 1
    assert checked > 100
