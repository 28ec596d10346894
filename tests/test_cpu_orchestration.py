"""The folded / hoisted orchestration (fusion.fused_forward + FCMF.head: aspect and image loops folded into the launch
dimension, per-sample and per-(sample, aspect) hoisting, two-segment attention plans, residual gathers, pooled rows, live
rows) checked ON THE CPU: the kernel Functions are replaced by plain-torch stand-ins with the same contracts
(tests/_standins.py), everything above them is the shipped code. Must equal the reference's per-aspect / per-image loop
(oracle, pinned against the reference's goldens) in logits, loss and every gradient -- SURVEY.md section 4 tiers (3) and (4)."""
import pytest
import torch

import _standins
from _util import pkg, rel_err, rel_err_floor, synth
from oracle import fcmf_oracle as O


def _oracle(params, batch, dims):
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    seq = batch["sequence_output"].clone().requires_grad_(True)
    logits, loss = O.aspect_loop(seq, batch["visual_embeds_att"], batch["roi_embeds_att"], batch["roi_coors"],
                                 batch["added_attention_mask"], batch["labels"], p, dims.heads, dims.num_imgs, dims.num_roi)
    loss.backward()
    return logits.detach(), loss.detach(), seq.grad, {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize("rows", ["full", "live"])
@pytest.mark.parametrize("num_roi", [3, 7])
def test_folded_orchestration_equals_reference_loops(monkeypatch, rows, num_roi):
    dims = synth.FusionDims(batch=2, aspects=3, seq_len=12, num_imgs=2, num_roi=num_roi)
    params = synth.make_params(dims, seed=31)
    batch = synth.make_batch(dims, seed=32, mask="bernoulli")
    want_logits, want_loss, want_dseq, want_grads = _oracle(params, batch, dims)

    _standins.install(monkeypatch, pkg)
    model = pkg().FCMF(None, num_labels=dims.num_labels, num_imgs=dims.num_imgs, num_roi=dims.num_roi)
    model.load_state_dict(params, strict=True)
    model.eval()
    model.encoder.compute_dtype = torch.float32
    seq = batch["sequence_output"].clone().requires_grad_(True)
    B, A = dims.batch, dims.aspects
    logits, loss = model.fuse_all_aspects(seq, batch["visual_embeds_att"], batch["roi_embeds_att"], batch["roi_coors"],
                                          batch["added_attention_mask"].reshape(B * A, -1), batch["labels"], rows=rows)
    loss.backward()
    assert rel_err(logits, want_logits) < 1e-5 and abs(loss.item() - want_loss.item()) < 1e-5 * max(1.0, abs(want_loss.item()))
    assert rel_err(seq.grad, want_dseq) < 1e-4
    gmax = max(float(g.abs().max()) for g in want_grads.values() if g is not None)
    for k, v in model.named_parameters():
        zero_grad = k.endswith("key.bias") or k.endswith("box_head.linears.1.bias")     # exactly-zero true gradients
        tol = 3e-3 if ".WGs." in k else 1e-4
        assert rel_err_floor(v.grad, want_grads[k], (1e-1 if zero_grad else 1e-3) * gmax) < tol, k


def test_per_aspect_forward_signature_on_stand_ins(monkeypatch):
    """FCMF.forward / FCMFEncoder.fuse with A = 1 (the reference's call, run_multimodal_fcmf.py:464-473) == column a of the fold."""
    dims = synth.FusionDims(batch=2, aspects=2, seq_len=10, num_imgs=2, num_roi=3)
    params = synth.make_params(dims, seed=5)
    batch = synth.make_batch(dims, seed=6, mask="bernoulli")
    _standins.install(monkeypatch, pkg)
    model = pkg().FCMF(None, num_labels=dims.num_labels, num_imgs=dims.num_imgs, num_roi=dims.num_roi)
    model.load_state_dict(params, strict=True)
    model.eval()
    model.encoder.compute_dtype = torch.float32
    B, A = dims.batch, dims.aspects
    with torch.no_grad():
        folded, _ = model.fuse_all_aspects(batch["sequence_output"], batch["visual_embeds_att"], batch["roi_embeds_att"],
                                           batch["roi_coors"], batch["added_attention_mask"].reshape(B * A, -1), None)
        for a in range(A):
            fused = model.encoder.fuse(batch["sequence_output"][:, a], batch["visual_embeds_att"], batch["roi_embeds_att"],
                                       batch["roi_coors"], batch["added_attention_mask"][:, a])
            one, _ = model.head(fused)
            assert rel_err(one, folded[:, a]) < 1e-5
