"""torch.library surface of the kernels (torch.ops.fcmf_b200.*): registration + fake (meta) shapes on the CPU, values and
gradients equal to the autograd-Function path on the GPU."""
import pytest
import torch

from _util import pkg, rel_err

T = pkg("torch_ops")


def test_ops_are_registered_with_fake_implementations():
    for name in T.REGISTERED:
        assert hasattr(torch.ops.fcmf_b200, name), name
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x = torch.empty(10, 16, dtype=torch.bfloat16, device="cuda") if torch.cuda.is_available() else torch.empty(10, 16, dtype=torch.bfloat16)
        w = torch.empty(24, 16, dtype=torch.float32, device=x.device)
        b = torch.empty(24, dtype=torch.float32, device=x.device)
        y = torch.ops.fcmf_b200.linear(x, w, b, "tanh")
        assert y.shape == (10, 24) and y.dtype == torch.bfloat16
        yy, mean, rstd = torch.ops.fcmf_b200.layer_norm_residual(x, x, torch.empty(16), torch.empty(16), 1e-12)
        assert yy.shape == x.shape and mean.shape == (10,) and rstd.dtype == torch.float32
        dw, db = torch.ops.fcmf_b200.gemm_wgrad(y, x)
        assert dw.shape == (24, 16) and db.shape == (24,)


@pytest.mark.gpu
@pytest.mark.parametrize("act", ["none", "tanh"])
def test_registered_linear_and_layernorm_equal_the_function_path(act):
    Fn = pkg("functional")
    torch.manual_seed(0)
    x = torch.randn(300, 768, device="cuda").bfloat16()
    w = (torch.randn(768, 768, device="cuda") * 0.05)
    b = torch.randn(768, device="cuda") * 0.1
    xa, wa, ba = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    xb, wb, bb = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ya = torch.ops.fcmf_b200.linear(xa, wa, ba, act)
    yb = Fn.linear(xb, wb, bb, act=act)
    assert torch.equal(ya, yb)
    g = torch.randn_like(ya)
    ya.backward(g); yb.backward(g)
    assert torch.equal(xa.grad, xb.grad) and rel_err(wa.grad, wb.grad) < 1e-5 and rel_err(ba.grad, bb.grad) < 1e-5
    gamma, beta = torch.ones(768, device="cuda") + 0.1 * torch.randn(768, device="cuda"), 0.1 * torch.randn(768, device="cuda")
    xr, rr = x.clone().requires_grad_(True), x.flip(0).clone().requires_grad_(True)
    y, _, _ = torch.ops.fcmf_b200.layer_norm_residual(xr, rr, gamma, beta, 1e-12)
    s = (xr.detach().float() + rr.detach().float()).requires_grad_(True)
    u = s.mean(-1, keepdim=True)
    ref = gamma * ((s - u) / torch.sqrt(((s - u) ** 2).mean(-1, keepdim=True) + 1e-12)) + beta
    assert rel_err(y, ref) < 2e-2
    y.backward(g)
    ref.backward(g.float())
    assert rel_err(xr.grad, s.grad) < 3e-2 and torch.equal(xr.grad, rr.grad)
