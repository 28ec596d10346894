"""tcgen05 engine (bf16, TMA + TMEM) against fp32 PyTorch on identical bf16 inputs, and against the CUDA-core
engine. bf16 bar: 2e-2 relative (north_star); the accumulators are fp32 so the observed error is ~1e-3."""
import math

import pytest
import torch

from _util import pkg, rel_err

pytestmark = pytest.mark.gpu
ops = pkg("ops")
L = pkg("_lib")
BF = torch.bfloat16


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).cuda().to(BF)


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def gelu_grad(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)


SHAPES = [(128, 128, 64), (128, 256, 128), (300, 768, 768), (1000, 1536, 2048), (77, 3072, 768), (2688, 768, 768),
          (40000, 768, 768), (20000, 3072, 768), (4097, 768, 3072), (513, 136, 72)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_tn_plain_and_bias(M, N, K):
    a, b = rnd(M, K), rnd(N, K, scale=0.05, seed=1)
    bias = torch.randn(N, device="cuda") * 0.1
    ref = a.float() @ b.float().t()
    out = ops.gemm_tn(a, b, None, L.EPI_NONE, engine=L.ENGINE_TCGEN05)
    assert rel_err(out, ref) < 1e-2, "descriptor/layout error (not rounding) if this is large"
    out = ops.gemm_tn(a, b, bias, L.EPI_NONE, engine=L.ENGINE_TCGEN05)
    assert rel_err(out, ref + bias) < 1e-2


@pytest.mark.parametrize("M,N,K", [(300, 768, 768), (5000, 3072, 768), (700, 256, 3072)])
def test_tc_gemm_tn_epilogues(M, N, K):
    a, b = rnd(M, K), rnd(N, K, scale=0.05, seed=1)
    bias = torch.randn(N, device="cuda") * 0.1
    ref = a.float() @ b.float().t() + bias
    g, pre = ops.gemm_tn(a, b, bias, L.EPI_GELU, engine=L.ENGINE_TCGEN05, want_aux=True)
    assert rel_err(pre, ref) < 1e-2 and rel_err(g, gelu(ref)) < 1e-2
    t = ops.gemm_tn(a, b, bias, L.EPI_TANH, engine=L.ENGINE_TCGEN05)
    assert rel_err(t, torch.tanh(ref)) < 1e-2
    aux = rnd(M, N, seed=5)
    d = ops.gemm_tn(a, b, None, L.EPI_DGELU, aux=aux, engine=L.ENGINE_TCGEN05)
    assert rel_err(d, (a.float() @ b.float().t()) * gelu_grad(aux.float())) < 1e-2


def test_tc_gemm_tn_strided_rows_and_output_column_block():
    x = rnd(300, 5, 768)
    w = rnd(768, 768, scale=0.05, seed=2)
    out = ops.gemm_tn(x[:, 0, :], w, None, L.EPI_TANH, engine=L.ENGINE_TCGEN05)
    assert rel_err(out, torch.tanh(x[:, 0, :].float() @ w.float().t())) < 1e-2
    wide = torch.zeros(300, 1536, dtype=BF, device="cuda")
    ops.gemm_tn(x[:, 1, :], w, None, L.EPI_NONE, out=wide[:, 768:], engine=L.ENGINE_TCGEN05)
    assert rel_err(wide[:, 768:], x[:, 1, :].float() @ w.float().t()) < 1e-2 and float(wide[:, :768].abs().max()) == 0


@pytest.mark.parametrize("M,N,K", [(64, 128, 128), (1000, 768, 768), (333, 3072, 768), (50000, 768, 768),
                                   (30000, 768, 3072), (4096, 768, 2048), (777, 136, 72)])
def test_tc_gemm_wgrad(M, N, K):
    dy, x = rnd(M, N), rnd(M, K, seed=3)
    ref_w, ref_b = dy.float().t() @ x.float(), dy.float().sum(0)
    dw, db = ops.gemm_wgrad(dy, x, engine=L.ENGINE_TCGEN05)
    assert rel_err(dw, ref_w) < 1e-2, "MN-major descriptor error if this is large"
    assert rel_err(db, ref_b) < 1e-3
    dw2, _ = ops.gemm_wgrad(dy, x, engine=L.ENGINE_TCGEN05, dw=dw.clone(), db=db.clone(), accumulate=True)
    assert rel_err(dw2, 2 * ref_w) < 1e-2


def test_tc_matches_cuda_core_engine_bitwise_close():
    a, b = rnd(999, 768), rnd(768, 768, scale=0.05, seed=1)
    o1 = ops.gemm_tn(a, b, None, L.EPI_NONE, engine=L.ENGINE_TCGEN05)
    o2 = ops.gemm_tn(a, b, None, L.EPI_NONE, engine=L.ENGINE_SIMT)
    assert rel_err(o1, o2.float()) < 5e-3


@pytest.mark.parametrize("M,N,K", [(2048, 768, 25088), (300, 256, 20032), (130, 136, 8192), (64, 768, 640)])
def test_tc_gemm_tn_f32_split_reduction(M, N, K):
    """fp32 output with the reduction split across CTAs (the vocabulary projection's input gradient)."""
    a, b = rnd(M, K, scale=0.1), rnd(N, K, scale=0.1, seed=3)
    ref = a.float() @ b.float().t()
    out = ops.gemm_tn_f32(a, b)
    assert out.dtype == torch.float32 and rel_err(out, ref) < 2e-3


def test_vocab_linear_padded_projection_fwd_bwd():
    """V % 8 != 0 and wide (padded to a multiple of 256): forward, split-K input gradient, weight / bias gradients."""
    Fn = pkg("functional")
    M, H, V = 96, 64, 5003
    x = rnd(M, H).requires_grad_(True)
    w = (torch.randn(V, H, device="cuda") * 0.05).requires_grad_(True)
    b = (torch.randn(V, device="cuda") * 0.1).requires_grad_(True)
    y = Fn.vocab_linear(x, w, b)
    assert y.shape == (M, V) and y.stride(0) % 256 == 0
    xr, wr, br = x.detach().float().requires_grad_(True), w.detach().to(BF).float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = xr @ wr.t() + br
    assert rel_err(y, yr) < 1e-2
    g = rnd(M, V, scale=0.1, seed=5)
    y.backward(g)
    yr.backward(g.float())
    assert rel_err(x.grad, xr.grad) < 1e-2 and rel_err(w.grad, wr.grad) < 1e-2 and rel_err(b.grad, br.grad) < 1e-2
    # through the vocabulary softmax-CE kernels (the gradient keeps the padded storage: no copy in backward)
    labels = torch.randint(0, V, (M,), device="cuda")
    x.grad = None; w.grad = None
    loss = Fn.vocab_cross_entropy(Fn.vocab_linear(x, w, b), labels)
    loss.backward()
    xr.grad = None; wr.grad = None
    lr = torch.nn.functional.cross_entropy(xr @ wr.t() + br, labels)
    lr.backward()
    assert abs(loss.item() - lr.item()) < 1e-2 * abs(lr.item())
    assert rel_err(x.grad, xr.grad) < 2e-2 and rel_err(w.grad, wr.grad) < 2e-2
