"""Pins the oracle (oracle/fcmf_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, made by oracle/make_golden.py from /root/reference). CPU only."""
import numpy as np
import pytest
import torch

from _util import golden_inputs, golden_sample, grad_scale, load_golden, rel_err, rel_err_floor
from oracle import fcmf_oracle as O

CASES = ["base_small", "base_roi7", "large_small", "base_cfg1_b1"]
TOL = 2e-5      # fp32 CPU vs fp32 CPU, different op order only


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_outputs(name):
    z, dims = load_golden(name)
    params, batch = golden_inputs(z, dims)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    seq = batch["sequence_output"].clone().requires_grad_(True)
    logits, loss = O.aspect_loop(seq, batch["visual_embeds_att"], batch["roi_embeds_att"], batch["roi_coors"],
                                 batch["added_attention_mask"], batch["labels"], p,
                                 dims.heads, dims.num_imgs, dims.num_roi)
    loss.backward()
    assert rel_err(logits, torch.from_numpy(z["logits"])) < TOL
    assert abs(loss.item() - float(z["loss"])) < TOL * max(1.0, abs(float(z["loss"])))
    stride = int(z["sample_stride"])
    gold_dseq = torch.from_numpy(z["d_sequence_output"])
    got = seq.grad if gold_dseq.numel() == seq.grad.numel() else golden_sample(seq.grad, stride)
    assert rel_err(got.reshape(-1), gold_dseq.reshape(-1)) < TOL
    for k, v in p.items():
        g = torch.from_numpy(z["gsample/" + k])
        assert rel_err_floor(golden_sample(v.grad, stride), g, 1e-3 * grad_scale(z)) < 5 * TOL, k
        assert abs(v.grad.double().norm().item() - float(z["gnorm/" + k])) <= 5 * TOL * float(z["gnorm/" + k]) + 1e-12, k


def test_geometry_embedding_is_float64_with_float32_frequencies():
    """roi_modeling.py:123-127: the frequency table is float32, the positions float64."""
    boxes = torch.tensor([[[0.1, 0.4, 0.2, 0.9], [0.0, 0.0, 0.0, 0.0], [0.3, 0.35, 0.5, 0.55]]], dtype=torch.float64)
    emb = O.box_relational_embedding(boxes)
    assert emb.dtype == torch.float64 and emb.shape == (1, 3, 3, 64)
    # diagonal: dx=dy=log(1e-3), dw=dh=0 -> sin(0)=0 / cos(0)=1 in the w,h blocks
    assert torch.allclose(emb[0, 0, 0, 16:32], torch.zeros(16, dtype=torch.float64))
    assert torch.allclose(emb[0, 0, 0, 48:64], torch.ones(16, dtype=torch.float64))
    f1 = float(1.0 / torch.pow(torch.tensor(1000.0), torch.tensor(1.0 / 8)))     # float32 rounding kept
    assert emb[0, 0, 0, 1].item() == pytest.approx(np.sin(100.0 * np.log(1e-3) * f1), abs=1e-12)


def test_loss_is_sum_of_per_aspect_means():
    """run_multimodal_fcmf.py:474-478."""
    logits = torch.randn(3, 2, 4)
    labels = torch.randint(0, 4, (3, 2))
    want = sum(torch.nn.functional.cross_entropy(logits[:, a], labels[:, a]) for a in range(2))
    folded = torch.nn.functional.cross_entropy(logits.reshape(-1, 4), labels.reshape(-1), reduction="sum") / 3
    assert torch.allclose(want, folded, atol=1e-6)


def test_oracle_dropout_plan_is_identity_at_p0_and_deterministic():
    """train()-mode oracle (DropPlan): p = 0 reproduces the eval path bit for bit, a seed reproduces its masks, masks of
    different sites / seeds differ, and every site's keep rate is the requested one."""
    import numpy as np
    from oracle import dropout_mask as DM
    from _util import pkg, synth
    sites = pkg("fusion").DROP_SITES
    dims = synth.FusionDims(batch=2, aspects=2, seq_len=12, num_imgs=2, num_roi=3)
    params = synth.make_params(dims, seed=3)
    batch = synth.make_batch(dims, seed=4, mask="bernoulli")

    def run(plan):
        with torch.no_grad():
            return O.aspect_loop(batch["sequence_output"], batch["visual_embeds_att"], batch["roi_embeds_att"],
                                 batch["roi_coors"], batch["added_attention_mask"], batch["labels"], params, dims.heads,
                                 dims.num_imgs, dims.num_roi, drop=plan)[0]
    base = run(None)
    assert torch.equal(run(O.DropPlan(1, sites, 2, 2, 2, 0.0)), base)
    a, b, c = run(O.DropPlan(7, sites, 2, 2, 2, 0.1)), run(O.DropPlan(7, sites, 2, 2, 2, 0.1)), run(O.DropPlan(8, sites, 2, 2, 2, 0.1))
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, base)
    only_head = run(O.DropPlan(7, sites, 2, 2, 2, {"head": 0.5}))
    assert not torch.equal(only_head, base)
    assert len(set(sites.values())) == len(sites)                    # one seed per nn.Dropout site
    for site in sites.values():
        m = DM.keep_mask(DM.site_seed(7, site), np.arange(2048), 256, 0.1)
        assert abs(m.mean() - 0.9) < 3e-3
    m1 = DM.keep_mask(DM.site_seed(7, 1), np.arange(256), 256, 0.1)
    m2 = DM.keep_mask(DM.site_seed(7, 2), np.arange(256), 256, 0.1)
    assert (m1 != m2).mean() > 0.1
