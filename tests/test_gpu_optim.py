"""clip_grad_norm_(1.0) + torch.optim.AdamW over the reference's 4 parameter groups against
the fused three-launch tail, several steps, fp32."""
import pytest
import torch

from _util import pkg, rel_err

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_torch_clip_and_adamw():
    torch.manual_seed(0)
    shapes = [(768, 768), (768,), (3072, 768), (3072,), (4, 768), (4,), (9000,), (1,)]
    ref = [torch.nn.Parameter(torch.randn(s, device="cuda") * 0.1) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]

    def groups(ps):
        return [{"params": [ps[0], ps[2]], "weight_decay": 0.01, "lr": 3e-3}, {"params": [ps[1], ps[3]], "weight_decay": 0.0, "lr": 3e-3},
                {"params": [ps[4], ps[6]], "weight_decay": 0.01, "lr": 1e-2}, {"params": [ps[5], ps[7]], "weight_decay": 0.0, "lr": 1e-2}]
    o_ref = torch.optim.AdamW(groups(ref), lr=1e-2)
    o_mine = pkg("optim").FusedAdamW(groups(mine), lr=1e-2, max_grad_norm=1.0)
    for step in range(4):
        for a, b in zip(ref, mine):
            g = torch.randn_like(a) * (5.0 if step % 2 == 0 else 0.01)            # clipped and unclipped steps
            a.grad, b.grad = g.clone(), g.clone()
        total = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o_ref.step()
        o_mine.step()
        assert abs(o_mine.grad_norm.item() - total.item()) < 1e-4 * total.item()
        for a, b in zip(ref, mine):
            assert rel_err(b, a) < 1e-5
            assert rel_err(b.grad, a.grad) < 1e-5                                  # the clipped gradient is left in place
        if step == 1:                                                              # a scheduler changes the group lr
            for o in (o_ref, o_mine):
                for g in o.param_groups:
                    g["lr"] *= 0.5
    sd = o_mine.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and len(sd["param_groups"]) == 4
