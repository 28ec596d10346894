"""train()-mode parity: every nn.Dropout site of the reference path, with the kernels' regenerated masks substituted
for torch's random stream on the reference side (oracle/dropout_mask.py restates the mask function in numpy and is
pinned against the library on the CPU, tests/test_cpu_host.py).

Kernel level: LayerNorm(dropout(x) + res) fwd/bwd, attention-probability dropout in all three attention engines
(CUDA-core, single-query, tcgen05) fwd/bwd, classifier dropout. Model level: the folded train() step against the oracle's
per-aspect / per-image loop with a DropPlan, fp32 at 1e-4 and bf16 at the bf16 bar; plus seeding behaviour."""
import math

import numpy as np
import pytest
import torch

from _util import pkg, rel_err, rel_err_floor, synth
from oracle import dropout_mask as DM
from test_gpu_ops import TOL, dev, ln_ref, rnd
from test_gpu_parity import build_model

pytestmark = pytest.mark.gpu

ops = pkg("ops")
Fn = pkg("functional")
L = pkg("_lib")
fusion = pkg("fusion")


def mask_t(seed, rows, ncols, p):
    return torch.from_numpy(DM.scaled_mask(seed, np.asarray(rows), ncols, p)).to(dev())


# ------------------------------------------------------------------------------------------- LayerNorm site
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("H,with_counter,M", [(768, False, 257), (1024, True, 257), (64, False, 257), (768, True, 7001), (896, False, 2100)])
def test_layernorm_hidden_dropout(dtype, H, with_counter, M):
    R, p, seed = 31, 0.1, 0x1234567887654321
    x, res = rnd(M, H, dtype=dtype), rnd(R, H, dtype=dtype, seed=2)
    idx = (torch.arange(M, device=dev()) % R).to(torch.int32)
    w, b = (1 + 0.1 * rnd(H, seed=4)), 0.1 * rnd(H, seed=5)
    counter = torch.tensor([41], dtype=torch.int64, device=dev()) if with_counter else None
    drop = ops.Drop(p, seed, counter)
    y, mean, rstd = ops.ln_fwd(x, res, idx, w, b, drop=drop)
    keep = mask_t(seed + (41 if with_counter else 0), np.arange(M), H, p)
    xs = x.float().clone().requires_grad_(True)
    rs = res.float().clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = ln_ref(xs * keep + rs[idx.long()], wr, br)
    assert rel_err(y, ref) < TOL[dtype]
    y0, _, _ = ops.ln_fwd(x, res, idx, w, b)
    assert rel_err(y0, ref) > 10 * TOL[dtype]                           # the mask really did something
    dy, dy2 = rnd(M, H, dtype=dtype, seed=7), rnd(M, H, dtype=dtype, seed=8)
    ref.backward(dy.float() + dy2.float())
    ds, dx, dg, db = ops.ln_bwd_drop(dy, dy2, x, res, idx, w, mean, rstd, drop)
    assert rel_err(dx, xs.grad) < TOL[dtype]
    assert rel_err(dg, wr.grad) < TOL[dtype] and rel_err(db, br.grad) < TOL[dtype]
    inv = torch.full((R, (M + R - 1) // R), -1, dtype=torch.int32)
    ar = torch.arange(M)
    inv[ar % R, ar // R] = ar.to(torch.int32)
    inv = inv.to(dev())
    assert rel_err(ops.gather_sum_rows(ds, inv, R, inv.shape[1]), rs.grad) < TOL[dtype]
    # exactly the dropped elements have a zero x-gradient
    assert torch.equal(dx.float() == 0, (keep == 0) | (ds.float() == 0))


# ------------------------------------------------------------------------------------------- attention-probability site
def attn_drop_ref(q, k, v, mask_add, keep, dh):
    s = q @ k.transpose(-1, -2) / math.sqrt(dh)
    if mask_add is not None:
        s = s + mask_add[:, None, None, : s.shape[-1]]
    return (torch.softmax(s, -1) * keep) @ v


@pytest.mark.parametrize("dtype,engine", [(torch.float32, L.ENGINE_SIMT), (torch.bfloat16, L.ENGINE_SIMT),
                                          (torch.bfloat16, L.ENGINE_TCGEN05), (torch.float32, L.ENGINE_AUTO)])
@pytest.mark.parametrize("heads,Lq,Lk,p", [(4, 9, 13, 0.1), (2, 170, 49, 0.1), (3, 174, 174, 0.1), (2, 130, 200, 0.5),
                                           (12, 1, 174, 0.1), (2, 64, 320, 0.25)])
def test_attention_probability_dropout(dtype, engine, heads, Lq, Lk, p):
    dh, NP, seed = 64, 5, 0x0F1E2D3C4B5A6978
    if engine == L.ENGINE_TCGEN05 and Lq < 16:
        pytest.skip("tcgen05 attention needs Lq >= 16")
    if engine == L.ENGINE_AUTO and Lq != 1:
        pytest.skip("AUTO + fp32 is the single-query engine's case")
    HD = heads * dh
    ops.set_attn_engine(engine)
    try:
        tq = rnd(NP * Lq, HD, dtype=dtype).requires_grad_(True)
        tkv = rnd(NP * Lk, 2 * HD, dtype=dtype, seed=3).requires_grad_(True)
        plan = Fn.AttnPlan(NP, heads, dh, drop=ops.Drop(p, seed)).add("q", 0, 0, Lq, None, None) \
            .add("k", 1, 0, Lk, None, None).add("v", 1, HD, Lk, None, None)
        mask = (torch.rand(NP, Lk, device=dev()) < 0.8).long()
        mask[:, 0] = 1
        mask_add = ops.mask_additive(mask, Lk)
        out = Fn.folded_attention(plan, (tq, tkv), mask_add, None)
        dout = rnd(NP * Lq, HD, dtype=dtype, seed=13)
        out.backward(dout)
    finally:
        ops.set_attn_engine(L.ENGINE_AUTO)
    rows = ((np.arange(NP).reshape(-1, 1, 1) * heads + np.arange(heads).reshape(1, -1, 1)) * Lq + np.arange(Lq).reshape(1, 1, -1))
    keep = mask_t(seed, rows.reshape(-1), Lk, p).view(NP, heads, Lq, Lk)
    rq = tq.detach().float().clone().requires_grad_(True)
    rkv = tkv.detach().float().clone().requires_grad_(True)
    q = rq.view(NP, Lq, heads, dh).permute(0, 2, 1, 3)
    k = rkv[:, :HD].reshape(NP, Lk, heads, dh).permute(0, 2, 1, 3)
    v = rkv[:, HD:].reshape(NP, Lk, heads, dh).permute(0, 2, 1, 3)
    ref = attn_drop_ref(q, k, v, mask_add, keep, dh).permute(0, 2, 1, 3).reshape(NP * Lq, HD)
    assert rel_err(out, ref) < TOL[dtype]
    ref.backward(dout.float())
    assert rel_err(tq.grad, rq.grad) < TOL[dtype]
    assert rel_err(tkv.grad, rkv.grad) < TOL[dtype]


def test_box_attention_dropout_with_geometry_bias():
    """roi_modeling.py:42-43: dropout on w_mn, the CUDA-core engine with the per-pair bias (d_k = 96)."""
    heads, dh, NR, G, p, seed = 8, 96, 4, 6, 0.1, 99
    HD = heads * dh
    t = rnd(G * NR, 3 * HD).requires_grad_(True)
    bias = rnd(G, heads, NR, NR, seed=11).requires_grad_(True)
    plan = Fn.AttnPlan(G, heads, dh, drop=ops.Drop(p, seed)).add("q", 0, 0, NR, None, None).add("k", 0, HD, NR, None, None) \
        .add("v", 0, 2 * HD, NR, None, None)
    out = Fn.folded_attention(plan, (t,), None, bias)
    dout = rnd(G * NR, HD, seed=13)
    out.backward(dout)
    rows = ((np.arange(G).reshape(-1, 1, 1) * heads + np.arange(heads).reshape(1, -1, 1)) * NR + np.arange(NR).reshape(1, 1, -1))
    keep = mask_t(seed, rows.reshape(-1), NR, p).view(G, heads, NR, NR)
    r = t.detach().clone().requires_grad_(True)
    rb = bias.detach().clone().requires_grad_(True)
    q, k, v = [r[:, i * HD:(i + 1) * HD].reshape(G, NR, heads, dh).permute(0, 2, 1, 3) for i in range(3)]
    s = q @ k.transpose(-1, -2) / math.sqrt(dh) + rb
    ref = ((torch.softmax(s, -1) * keep) @ v).permute(0, 2, 1, 3).reshape(G * NR, HD)
    assert rel_err(out, ref) < 1e-4
    ref.backward(dout)
    assert rel_err(t.grad, r.grad) < 1e-4 and rel_err(bias.grad, rb.grad) < 1e-4


# ------------------------------------------------------------------------------------------- classifier site
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_classifier_dropout(dtype):
    R, H, Cn, B, p, seed = 12, 768, 4, 3, 0.1, 31337
    pooled = rnd(R, H, dtype=dtype).requires_grad_(True)
    wc, bc = rnd(Cn, H, scale=0.05).requires_grad_(True), rnd(Cn, scale=0.1).requires_grad_(True)
    labels = torch.randint(0, Cn, (R,), device=dev())
    logits, loss = Fn.classifier_ce(pooled, wc, bc, labels, 1.0 / B, ops.Drop(p, seed))
    keep = mask_t(seed, np.arange(R), H, p)
    pr = pooled.detach().float().clone().requires_grad_(True)
    wr, br = wc.detach().clone().requires_grad_(True), bc.detach().clone().requires_grad_(True)
    ref_logits = (pr * keep) @ wr.t() + br
    ref_loss = torch.nn.functional.cross_entropy(ref_logits, labels, reduction="sum") / B
    assert rel_err(logits, ref_logits) < TOL[dtype]
    loss.backward()
    ref_loss.backward()
    assert rel_err(pooled.grad, pr.grad) < TOL[dtype] and rel_err(wc.grad, wr.grad) < TOL[dtype] and rel_err(bc.grad, br.grad) < TOL[dtype]


# ------------------------------------------------------------------------------------------- the folded train() step
def _train_step(model, batch, dims, rows, step_seed):
    seq = batch["sequence_output"].cuda().clone().requires_grad_(True)
    B, A = dims.batch, dims.aspects
    model.zero_grad()
    logits, loss = model.fuse_all_aspects(seq, batch["visual_embeds_att"].cuda(), batch["roi_embeds_att"].cuda(),
                                          batch["roi_coors"].cuda(), batch["added_attention_mask"].cuda().reshape(B * A, -1),
                                          batch["labels"].cuda(), rows=rows, step_seed=step_seed)
    loss.backward()
    return logits.detach(), loss.detach(), seq.grad


def _oracle_train_step(params, batch, dims, step_seed, live):
    from oracle import fcmf_oracle as O
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    seq = batch["sequence_output"].clone().requires_grad_(True)
    plan = O.DropPlan(step_seed, fusion.DROP_SITES, dims.batch, dims.aspects, dims.num_imgs, 0.1, live=live)
    logits, loss = O.aspect_loop(seq, batch["visual_embeds_att"], batch["roi_embeds_att"], batch["roi_coors"],
                                 batch["added_attention_mask"], batch["labels"], p, dims.heads, dims.num_imgs,
                                 dims.num_roi, drop=plan)
    loss.backward()
    return logits.detach(), loss.detach(), seq.grad, {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize("rows", ["full", "live"])
def test_train_mode_step_matches_oracle_with_the_same_masks_fp32(rows):
    dims = synth.FusionDims(batch=2, aspects=3, seq_len=24, num_imgs=2, num_roi=3)
    params = synth.make_params(dims, seed=21)
    batch = synth.make_batch(dims, seed=22, mask="bernoulli")
    model = build_model(dims, params, torch.float32, rows, L.ENGINE_SIMT).train()
    seed = 0x5EED5EED5EED
    logits, loss, dseq = _train_step(model, batch, dims, rows, seed)
    r_logits, r_loss, r_dseq, r_grads = _oracle_train_step(params, batch, dims, seed, rows == "live")
    assert rel_err(logits, r_logits) < 1e-4 and abs(loss.item() - r_loss.item()) < 1e-4 * max(1.0, abs(r_loss.item()))
    assert rel_err(dseq, r_dseq) < 1e-4
    gmax = max(float(g.abs().max()) for g in r_grads.values() if g is not None)
    for k, v in model.named_parameters():
        tol = 3e-3 if ".WGs." in k else 1e-4
        zero_grad = k.endswith("key.bias") or k.endswith("box_head.linears.1.bias")
        assert rel_err_floor(v.grad, r_grads[k], (1e-1 if zero_grad else 1e-3) * gmax) < tol, k
    # eval() is the dropout-free path and differs; the same seed reproduces; another seed gives other masks
    again, _, _ = _train_step(model, batch, dims, rows, seed)
    other, _, _ = _train_step(model, batch, dims, rows, seed + 1)
    assert torch.equal(again, logits) and not torch.equal(other, logits)
    ev, _, _ = _train_step(model.eval(), batch, dims, rows, None)
    assert rel_err(ev, logits) > 1e-3


@pytest.mark.parametrize("rows", ["full", "live"])
def test_train_mode_step_bf16_tensor_core_engines(rows):
    """bf16, tcgen05 GEMMs and attention with in-kernel dropout vs the fp32 oracle with the same masks."""
    dims = synth.FusionDims(batch=4, aspects=3, seq_len=40, num_imgs=2, num_roi=4)
    params = synth.make_params(dims, seed=7)
    batch = synth.make_batch(dims, seed=11, mask="bernoulli")
    model = build_model(dims, params, torch.bfloat16, rows, L.ENGINE_AUTO).train()
    seed = 777
    logits, loss, dseq = _train_step(model, batch, dims, rows, seed)
    r_logits, r_loss, r_dseq, r_grads = _oracle_train_step(params, batch, dims, seed, rows == "live")
    assert rel_err(logits, r_logits) < 2e-2
    assert rel_err(dseq.float(), r_dseq) < 6e-2
    gmax = max(float(g.abs().max()) for g in r_grads.values() if g is not None)
    for k, v in model.named_parameters():
        assert torch.isfinite(v.grad).all(), k
        if ".WGs." in k or k.endswith("key.bias") or k.endswith("box_head.linears.1.bias"):
            continue
        assert rel_err_floor(v.grad, r_grads[k], 1e-2 * gmax) < 8e-2, k


@pytest.mark.parametrize("rows", ["full", "live"])
def test_train_mode_step_bf16_at_config_shape(rows):
    """The headline mode of bench.py -- train(), bf16, tcgen05 engines -- at the measured SHAPE (L = 170, 7 images x 49
    patches, 4 ROIs: two query tiles, three key blocks in the text+ROI attention) against the fp32 oracle fed the same masks."""
    dims = synth.FusionDims(batch=2, aspects=2, seq_len=170, num_imgs=7, num_roi=4)
    params = synth.make_params(dims, seed=31)
    batch = synth.make_batch(dims, seed=32, mask="bernoulli")
    model = build_model(dims, params, torch.bfloat16, rows, L.ENGINE_AUTO).train()
    seed = 0xC0FFEE
    logits, loss, dseq = _train_step(model, batch, dims, rows, seed)
    r_logits, r_loss, r_dseq, r_grads = _oracle_train_step(params, batch, dims, seed, rows == "live")
    assert rel_err(logits, r_logits) < 2e-2
    assert rel_err(dseq.float(), r_dseq) < 6e-2
    gmax = max(float(g.abs().max()) for g in r_grads.values() if g is not None)
    for k, v in model.named_parameters():
        assert torch.isfinite(v.grad).all(), k
        if ".WGs." in k or k.endswith("key.bias") or k.endswith("box_head.linears.1.bias"):
            continue
        assert rel_err_floor(v.grad, r_grads[k], 1e-2 * gmax) < 8e-2, k


def test_default_seeding_follows_torch_manual_seed_and_graph_replays_redraw():
    dims = synth.FusionDims(batch=2, aspects=2, seq_len=24, num_imgs=2, num_roi=3)
    params = synth.make_params(dims, seed=5)
    batch = synth.make_batch(dims, seed=9)
    model = build_model(dims, params, torch.bfloat16, "live", L.ENGINE_AUTO).train()
    torch.manual_seed(123)
    a, _, _ = _train_step(model, batch, dims, "live", None)
    b, _, _ = _train_step(model, batch, dims, "live", None)
    torch.manual_seed(123)
    c, _, _ = _train_step(model, batch, dims, "live", None)
    assert torch.equal(a, c) and not torch.equal(a, b)
    BA = dims.batch * dims.aspects
    inp = {"seq": batch["sequence_output"].reshape(BA, dims.seq_len, dims.hidden).cuda().bfloat16(),
           "vis": batch["visual_embeds_att"].cuda().bfloat16(), "roi": batch["roi_embeds_att"].cuda().bfloat16(),
           "coors": batch["roi_coors"].cuda(), "mask": batch["added_attention_mask"].reshape(BA, -1).cuda(),
           "labels": batch["labels"].cuda()}
    model = build_model(dims, params, torch.bfloat16, "live", L.ENGINE_AUTO).train()
    step = pkg("graphed").GraphedFusionStep(model, inp, aspects=dims.aspects, rows="live")
    l1 = step(inp)[0].clone()
    l2 = step(inp)[0].clone()
    torch.cuda.synchronize()
    assert torch.isfinite(l1).all() and not torch.equal(l1, l2)        # the device seed counter advanced between replays
