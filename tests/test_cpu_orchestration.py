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


@pytest.mark.parametrize("rows", ["full", "live"])
def test_train_mode_dropout_sites_on_stand_ins(monkeypatch, rows):
    """train(): fusion.py hands every nn.Dropout site of the reference its own seed and the kernels' row convention -- with
    the stand-ins applying the numpy masks exactly as the kernels regenerate them, the folded step equals the reference loops
    with the same masks (oracle DropPlan)."""
    dims = synth.FusionDims(batch=2, aspects=2, seq_len=10, num_imgs=2, num_roi=3)
    params = synth.make_params(dims, seed=41)
    batch = synth.make_batch(dims, seed=42, mask="bernoulli")
    seed = 0xABCDEF
    plan = O.DropPlan(seed, pkg("fusion").DROP_SITES, dims.batch, dims.aspects, dims.num_imgs, 0.1, live=(rows == "live"))
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    seq_ref = batch["sequence_output"].clone().requires_grad_(True)
    want_logits, want_loss = O.aspect_loop(seq_ref, batch["visual_embeds_att"], batch["roi_embeds_att"], batch["roi_coors"],
                                           batch["added_attention_mask"], batch["labels"], p, dims.heads, dims.num_imgs,
                                           dims.num_roi, drop=plan)
    want_loss.backward()
    _standins.install(monkeypatch, pkg)
    model = pkg().FCMF(None, num_labels=dims.num_labels, num_imgs=dims.num_imgs, num_roi=dims.num_roi)
    model.load_state_dict(params, strict=True)
    model.train()
    model.encoder.compute_dtype = torch.float32
    seq = batch["sequence_output"].clone().requires_grad_(True)
    B, A = dims.batch, dims.aspects
    logits, loss = model.fuse_all_aspects(seq, batch["visual_embeds_att"], batch["roi_embeds_att"], batch["roi_coors"],
                                          batch["added_attention_mask"].reshape(B * A, -1), batch["labels"], rows=rows,
                                          step_seed=seed)
    loss.backward()
    assert rel_err(logits, want_logits) < 1e-5 and abs(loss.item() - want_loss.item()) < 1e-5 * max(1.0, abs(want_loss.item()))
    assert rel_err(seq.grad, seq_ref.grad) < 1e-4
    model.eval()
    with torch.no_grad():
        ev, _ = model.fuse_all_aspects(batch["sequence_output"], batch["visual_embeds_att"], batch["roi_embeds_att"],
                                       batch["roi_coors"], batch["added_attention_mask"].reshape(B * A, -1), None, rows=rows)
    assert rel_err(ev, want_logits) > 1e-3                             # the masks really acted


def test_submodule_drop_ins_on_stand_ins(monkeypatch):
    """SURVEY.md 8(b2) on the CPU: BertCrossEncoder / MultimodalEncoder / BoxMultiHeadedAttention / BertPooler keep the
    reference's call signatures and return values (list of layer outputs, [G, NR, H], [B, H])."""
    dims = synth.FusionDims(batch=3, aspects=1, seq_len=20, num_imgs=1, num_roi=5)
    params = synth.make_params(dims, seed=3)
    _standins.install(monkeypatch, pkg)
    model = pkg().FCMF(None, num_labels=dims.num_labels, num_imgs=dims.num_imgs, num_roi=dims.num_roi)
    model.load_state_dict(params, strict=True)
    enc = model.eval().encoder
    g = torch.Generator().manual_seed(0)
    s1, s2 = torch.randn(3, 20, 768, generator=g), torch.randn(3, 49, 768, generator=g)
    m = (torch.rand(3, 49, generator=g) < 0.8).long()
    m[:, 0] = 1
    with torch.no_grad():
        ext = O.extended_mask(m, 49)
        outs = enc.text2img_attention(s1, s2, ext)
        assert isinstance(outs, list) and len(outs) == 1
        assert rel_err(outs[-1], O.encoder_layer(s1, s2, ext, params, "encoder.text2img_attention.layer.0", dims.heads)) < 1e-5
        ext2 = O.extended_mask(m, 20)
        assert rel_err(enc.mm_attention(s1, ext2)[-1],
                       O.encoder_layer(s1, s1, ext2, params, "encoder.mm_attention.layer.0", dims.heads)) < 1e-5
        boxes = synth.make_batch(dims, seed=5)["roi_coors"][:, 0]
        x = torch.randn(3, 5, 768, generator=g)
        assert rel_err(enc.box_head(x, x, x, boxes), O.box_multihead_attention(x, boxes, params, "encoder.box_head")) < 1e-5
        assert rel_err(enc.text2img_pooler(s1), O.first_token_pooler(s1, params, "encoder.text2img_pooler")) < 1e-5
        # the un-fused sub-module path (BertAttention -> BertIntermediate -> BertOutput) agrees with the fused layer
        layer = enc.mm_attention.layer[0]
        a = layer.attention(s1, ext2)
        assert rel_err(layer.output(layer.intermediate(a), a), enc.mm_attention(s1, ext2)[-1]) < 1e-5


def test_iaog_decoder_wiring_on_stand_ins(monkeypatch):
    """IAOGDecoder (per-head weights folded into one GEMM, the slot->head gather, keys as values, tril masked_fill on self- and
    cross-attention, tied output projection) on the CPU against the golden of the reference decoder."""
    import os
    import sys
    import numpy as np
    from _util import GOLD, ROOT
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden_iaog as G
    z = np.load(os.path.join(GOLD, "iaog_decoder.npz"))
    _standins.install(monkeypatch, pkg)
    dec = pkg("iaog").IAOGDecoder(vocab_size=G.VOCAB)
    missing, unexpected = dec.load_state_dict(G.decoder_params(), strict=False)
    assert not unexpected and all(k.startswith("pos_encoding") for k in missing)
    dec.eval()
    enc, dec_x, labels = G.inputs()
    enc = enc.requires_grad_(True)
    mask = torch.ones(enc.shape[0], enc.shape[1], dtype=torch.int64)
    logits = dec(dec_x, [enc, mask, [None] * dec.num_blks], is_train=True)
    loss = torch.nn.functional.cross_entropy(logits.permute(0, 2, 1), labels, ignore_index=-100)
    loss.backward()
    assert rel_err(logits, torch.from_numpy(z["logits"])) < 5e-5
    assert abs(loss.item() - float(z["loss"])) < 5e-5 * abs(float(z["loss"]))
    assert rel_err(enc.grad, torch.from_numpy(z["d_enc"])) < 1e-4
    assert rel_err(dec.embedding.weight.grad[:8], torch.from_numpy(z["g_embedding"])) < 1e-4
