"""Per-kernel parity (C ABI through ops.py) against plain PyTorch fp32 on the same GPU inputs.
fp32 storage: 1e-4 relative (north_star's fp32 bar); bf16 storage: 2e-2 relative."""
import math

import pytest
import torch

from _util import pkg, rel_err

pytestmark = pytest.mark.gpu

ops = pkg("ops")
Fn = pkg("functional")
L = pkg("_lib")

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}
DTYPES = [torch.float32, torch.bfloat16]


def dev():
    return torch.device("cuda:0")


def rnd(*shape, dtype=torch.float32, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dev()).to(dtype)


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def gelu_grad(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)


# ------------------------------------------------------------------------------------------- GEMM (CUDA-core engine)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(300, 768, 768), (65, 136, 72), (1000, 1536, 2048), (7, 3072, 768)])
def test_gemm_tn_simt(dtype, M, N, K):
    a, b, bias = rnd(M, K, dtype=dtype), rnd(N, K, dtype=dtype, scale=0.05), rnd(N, scale=0.1)
    ref = a.float() @ b.float().t() + bias
    out = ops.gemm_tn(a, b, bias, L.EPI_NONE, engine=L.ENGINE_SIMT)
    assert rel_err(out, ref) < TOL[dtype]
    g, pre = ops.gemm_tn(a, b, bias, L.EPI_GELU, engine=L.ENGINE_SIMT, want_aux=True)
    assert rel_err(pre, ref) < TOL[dtype] and rel_err(g, gelu(ref)) < TOL[dtype]
    t = ops.gemm_tn(a, b, bias, L.EPI_TANH, engine=L.ENGINE_SIMT)
    assert rel_err(t, torch.tanh(ref)) < TOL[dtype]
    aux = rnd(M, N, dtype=dtype, seed=5)
    d = ops.gemm_tn(a, b, None, L.EPI_DGELU, aux=aux, engine=L.ENGINE_SIMT)
    assert rel_err(d, (a.float() @ b.float().t()) * gelu_grad(aux.float())) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_tn_strided_rows(dtype):
    """The pooler reads token 0 of every [Lq, H] block: rows with a large stride (mm_modeling.py:428)."""
    x = rnd(40, 9, 256, dtype=dtype)
    w, bias = rnd(128, 256, dtype=dtype, scale=0.05), rnd(128, scale=0.1)
    out = ops.gemm_tn(x[:, 0, :], w, bias, L.EPI_TANH, engine=L.ENGINE_SIMT)
    assert rel_err(out, torch.tanh(x[:, 0, :].float() @ w.float().t() + bias)) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(1000, 768, 768), (333, 3072, 136), (64, 8, 2048)])
def test_gemm_wgrad_simt(dtype, M, N, K):
    dy, x = rnd(M, N, dtype=dtype), rnd(M, K, dtype=dtype, seed=3)
    dw, db = ops.gemm_wgrad(dy, x, engine=L.ENGINE_SIMT)
    ref_w, ref_b = dy.float().t() @ x.float(), dy.float().sum(0)
    assert rel_err(dw, ref_w) < 1e-4 and rel_err(db, ref_b) < 1e-4
    dw2, db2 = ops.gemm_wgrad(dy, x, engine=L.ENGINE_SIMT, dw=dw.clone(), db=db.clone(), accumulate=True)
    assert rel_err(dw2, 2 * ref_w) < 1e-4 and rel_err(db2, 2 * ref_b) < 1e-4


def test_gemm_argument_errors_are_loud():
    a, b = rnd(8, 16), rnd(4, 24)
    with pytest.raises(RuntimeError):
        ops.gemm_tn(a, b)
    with pytest.raises(TypeError):
        ops.gemm_tn(a.double(), a.double())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm_tn(a.cpu(), a.cpu())
    with pytest.raises(RuntimeError, match="tcgen05"):
        ops.gemm_tn(a, rnd(4, 16), engine=L.ENGINE_TCGEN05)      # fp32 operands cannot use the bf16 tensor path


# ------------------------------------------------------------------------------------------- LayerNorm
def ln_ref(s, w, b, eps=1e-12):
    u = s.mean(-1, keepdim=True)
    v = (s - u).pow(2).mean(-1, keepdim=True)
    return w * ((s - u) / torch.sqrt(v + eps)) + b


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("H,M", [(768, 257), (1024, 257), (64, 257), (320, 300), (896, 129), (768, 9001), (64, 40013)])
def test_layernorm_fwd_bwd(dtype, H, M):
    # H picks the warps-per-row instantiation (one 16-byte vector per lane; 896 leaves the last warp half empty); the large M
    # make every group walk several rows (grid-stride loop, next-row prefetch, double-buffered exchange slots)
    R = 31
    x, res = rnd(M, H, dtype=dtype), rnd(R, H, dtype=dtype, seed=2)
    idx = (torch.arange(M, device=dev()) % R).to(torch.int32)
    w, b = (1 + 0.1 * rnd(H, seed=4)), 0.1 * rnd(H, seed=5)
    y, mean, rstd = ops.ln_fwd(x, res, idx, w, b)
    xs = x.float().clone().requires_grad_(True)
    rs = res.float().clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = ln_ref(xs + rs[idx.long()], wr, br)
    assert rel_err(y, ref) < TOL[dtype]
    dy, dy2 = rnd(M, H, dtype=dtype, seed=7), rnd(M, H, dtype=dtype, seed=8)
    ref.backward(dy.float() + dy2.float())
    ds, dg, db = ops.ln_bwd(dy, dy2, x, res, idx, w, mean, rstd)
    assert rel_err(ds, xs.grad) < TOL[dtype]
    assert rel_err(dg, wr.grad) < TOL[dtype] and rel_err(db, br.grad) < TOL[dtype]
    # residual gradient = rows of ds summed per residual row (what layer_tail does with res_inv)
    inv = torch.full((R, (M + R - 1) // R), -1, dtype=torch.int32)
    ar = torch.arange(M)
    inv[ar % R, ar // R] = ar.to(torch.int32)
    inv = inv.to(dev())
    dres = ops.gather_sum_rows(ds, inv, R, inv.shape[1])
    assert rel_err(dres, rs.grad) < TOL[dtype]


def test_layernorm_no_residual_matches_reference_eps_placement():
    x = rnd(16, 768) * 1e-7          # tiny variance: eps=1e-12 inside the sqrt matters
    w, b = torch.ones(768, device=dev()), torch.zeros(768, device=dev())
    y, _, _ = ops.ln_fwd(x, None, None, w, b)
    assert rel_err(y, ln_ref(x, w, b)) < 1e-4


# ------------------------------------------------------------------------------------------- small row-wise kernels
def test_mask_additive_and_short_mask_error():
    m = (torch.rand(5, 60, device=dev()) < 0.7).long()
    add = ops.mask_additive(m, 49)
    assert torch.equal(add, (1.0 - m[:, :49].float()) * -10000.0)
    with pytest.raises(RuntimeError, match="shorter"):
        ops.mask_additive(m, 61)


@pytest.mark.parametrize("dtype", DTYPES)
def test_gather_sum_rows_and_dtanh_and_casts(dtype):
    src = rnd(50, 128, dtype=dtype)
    idx = torch.tensor([[0, 3, -1], [49, 49, 2], [-1, -1, -1], [7, 8, 9]], dtype=torch.int32, device=dev())
    out = ops.gather_sum_rows(src, idx, 4, 3)
    ref = torch.stack([src[[0, 3]].float().sum(0), src[[49, 49, 2]].float().sum(0), torch.zeros(128, device=dev()),
                       src[[7, 8, 9]].float().sum(0)])
    assert rel_err(out, ref) < TOL[dtype]
    wide = torch.zeros(4, 256, dtype=dtype, device=dev())
    ops.gather_sum_rows(src, idx, 4, 3, out=wide[:, 128:])            # strided destination (column block)
    assert rel_err(wide[:, 128:], ref) < TOL[dtype] and float(wide[:, :128].abs().max()) == 0.0
    y = torch.tanh(rnd(33, 40, dtype=dtype))
    dy = rnd(33, 40, dtype=dtype, seed=1)
    assert rel_err(ops.dtanh(dy, y), dy.float() * (1 - y.float() ** 2)) < TOL[dtype]
    w = rnd(70, 130)
    assert rel_err(ops.cast_matrix(w, dtype, transpose=True), w.t()) < TOL[dtype]
    assert rel_err(ops.cast_matrix(w, dtype), w) < TOL[dtype]
    assert rel_err(ops.cast_to_f32(src), src.float()) == 0.0


# ------------------------------------------------------------------------------------------- attention
def attn_ref(qs, ks, vs, idxq, idxk, mask_add, mask_div, bias, heads, dh):
    """qs/ks/vs: lists of (tensor [groups*rows, heads*dh] fp32, rows, idx) -> ctx [NP, Lq, heads*dh]."""
    NP = idxq[0].numel()

    def build(segs, idxs):
        parts = []
        for (t, rows), ix in zip(segs, idxs):
            parts.append(t.view(-1, rows, heads, dh)[ix.long()])             # [NP, rows, heads, dh]
        return torch.cat(parts, 1).permute(0, 2, 1, 3)                         # [NP, heads, L, dh]
    q, k, v = build(qs, idxq), build(ks, idxk), build(vs, idxk)
    s = q @ k.transpose(-1, -2) / math.sqrt(dh)
    if mask_add is not None:
        rows = torch.arange(NP, device=s.device) // mask_div
        s = s + mask_add[rows][:, None, None, : s.shape[-1]]
    if bias is not None:
        s = s + bias
    p = torch.softmax(s, -1)
    return (p @ v).permute(0, 2, 1, 3).reshape(NP, q.shape[2], heads * dh)


@pytest.mark.parametrize("dtype,engine", [(torch.float32, L.ENGINE_SIMT), (torch.bfloat16, L.ENGINE_SIMT),
                                          (torch.bfloat16, L.ENGINE_TCGEN05)])
@pytest.mark.parametrize("two_seg,use_bias,dh,heads,L1,L2", [
    (True, False, 64, 4, 9, 4), (False, True, 96, 8, 7, 0), (False, False, 64, 12, 7, 0),
    (True, False, 64, 2, 170, 4), (False, False, 64, 3, 49, 0), (False, False, 64, 1, 200, 0), (True, False, 64, 2, 130, 30),
    (False, False, 64, 2, 100, 0), (True, False, 64, 2, 90, 20), (False, False, 64, 2, 128, 0), (True, False, 64, 1, 32, 8)])
def test_folded_attention_fwd_bwd(dtype, engine, two_seg, use_bias, dh, heads, L1, L2):
    if engine == L.ENGINE_TCGEN05 and (use_bias or dh != 64 or L1 + L2 < 16):
        pytest.skip("tcgen05 attention covers bf16, head_dim 64, no bias, L >= 16")
    ops.set_attn_engine(engine)
    try:
        _attention_case(dtype, two_seg, use_bias, dh, heads, L1, L2)
    finally:
        ops.set_attn_engine(L.ENGINE_AUTO)


def _attention_case(dtype, two_seg, use_bias, dh, heads, L1, L2):
    HD = heads * dh
    NI, A, B = 3, 2, 2
    BA, NP = B * A, B * A * NI
    p = torch.arange(NP, device=dev(), dtype=torch.int32)
    p2ba, p2bi = (p // NI).contiguous(), ((p // (A * NI)) * NI + p % NI).contiguous()
    ba2p = (torch.arange(BA, device=dev(), dtype=torch.int32).view(BA, 1) * NI + torch.arange(NI, device=dev(), dtype=torch.int32)).contiguous()
    bi2p = torch.stack([torch.stack([torch.tensor((b * A + a) * NI + i) for a in range(A)]) for b in range(B) for i in range(NI)]).to(dev()).to(torch.int32).contiguous()
    t0 = rnd(BA * L1, 3 * HD, dtype=dtype).requires_grad_(True)                 # packed q|k|v, groups indexed by ba
    tensors = [t0]
    plan = Fn.AttnPlan(NP, heads, dh, mask_div=NI)
    for role, col in (("q", 0), ("k", HD), ("v", 2 * HD)):
        plan.add(role, 0, col, L1, p2ba, ba2p)
    if two_seg:
        t1 = rnd(B * NI * L2, 3 * HD, dtype=dtype, seed=9).requires_grad_(True)  # groups indexed by (b, i)
        tensors.append(t1)
        for role, col in (("q", 0), ("k", HD), ("v", 2 * HD)):
            plan.add(role, 1, col, L2, p2bi, bi2p)
    Lt = L1 + L2
    mask = (torch.rand(BA, Lt + 5, device=dev()) < 0.8).long()
    mask[:, 0] = 1
    mask_add = ops.mask_additive(mask, Lt + 5)
    bias = rnd(NP, heads, Lt, Lt, seed=11).requires_grad_(True) if use_bias else None
    out = Fn.folded_attention(plan, tensors, mask_add, bias)
    dout = rnd(NP * Lt, HD, dtype=dtype, seed=13)
    out.backward(dout)

    refs = [t.detach().float().clone().requires_grad_(True) for t in tensors]
    rb = bias.detach().clone().requires_grad_(True) if use_bias else None
    idxs = [p2ba, p2bi][: len(tensors)]
    rows = [L1, L2][: len(tensors)]
    ref = attn_ref([(r[:, :HD], n) for r, n in zip(refs, rows)], [(r[:, HD:2 * HD], n) for r, n in zip(refs, rows)],
                   [(r[:, 2 * HD:], n) for r, n in zip(refs, rows)], idxs, idxs, mask_add, NI, rb, heads, dh)
    assert rel_err(out.view(NP, Lt, HD), ref) < TOL[dtype]
    ref.backward(dout.float().view(NP, Lt, HD))
    for t, r in zip(tensors, refs):
        assert rel_err(t.grad, r.grad) < TOL[dtype]
    if use_bias:
        assert rel_err(bias.grad, rb.grad) < TOL[dtype]


# ------------------------------------------------------------------------------------------- geometry + head
def test_box_geometry_matches_oracle_fp64_embedding():
    from oracle import fcmf_oracle as O
    synth = pkg("synth")
    dims = synth.FusionDims(batch=3, num_imgs=2, num_roi=5)
    boxes = synth.make_batch(dims, seed=3)["roi_coors"].reshape(-1, 5, 4)
    g = torch.Generator().manual_seed(17)
    wg_w = (0.3 * torch.randn(8, 64, generator=g)).requires_grad_(True)
    wg_b = (0.5 + 0.3 * torch.randn(8, generator=g)).requires_grad_(True)
    emb = O.box_relational_embedding(boxes).float()
    z = torch.relu(torch.einsum("gijc,hc->ghij", emb, wg_w) + wg_b.view(1, 8, 1, 1))
    ref = torch.log(torch.clamp(z, min=1e-6))
    dbias = torch.randn(ref.shape, generator=g)
    ref.backward(dbias)
    w_d, b_d = wg_w.detach().to(dev()).requires_grad_(True), wg_b.detach().to(dev()).requires_grad_(True)
    got = Fn.box_geometry(boxes.to(dev()), w_d, b_d)
    assert rel_err(got, ref) < 1e-4
    got.backward(dbias.to(dev()))
    # d log(z)/dz = 1/z with z down to 1e-6 amplifies the fp32 summation-order noise of z = WG.emb + b: 3e-3, not 1e-4
    assert rel_err(w_d.grad, wg_w.grad) < 3e-3 and rel_err(b_d.grad, wg_b.grad) < 3e-3


@pytest.mark.parametrize("dtype", DTYPES)
def test_classifier_cross_entropy(dtype):
    R, H, Cn, B = 12, 768, 4, 3
    pooled = rnd(R, H, dtype=dtype).requires_grad_(True)
    wc, bc = rnd(Cn, H, scale=0.05).requires_grad_(True), rnd(Cn, scale=0.1).requires_grad_(True)
    labels = torch.randint(0, Cn, (R,), device=dev())
    logits, loss = Fn.classifier_ce(pooled, wc, bc, labels, 1.0 / B)
    pr = pooled.detach().float().clone().requires_grad_(True)
    wr, br = wc.detach().clone().requires_grad_(True), bc.detach().clone().requires_grad_(True)
    ref_logits = pr @ wr.t() + br
    ref_loss = torch.nn.functional.cross_entropy(ref_logits, labels, reduction="sum") / B
    assert rel_err(logits, ref_logits) < TOL[dtype] and abs(loss.item() - ref_loss.item()) < TOL[dtype] * max(1, abs(ref_loss.item()))
    loss.backward()
    ref_loss.backward()
    assert rel_err(pooled.grad, pr.grad) < TOL[dtype] and rel_err(wc.grad, wr.grad) < TOL[dtype] and rel_err(bc.grad, br.grad) < TOL[dtype]
    # logits-only path (loss computed by the caller, as run_multimodal_fcmf.py:474 does)
    p2 = pooled.detach().clone().requires_grad_(True)
    lg, _ = Fn.classifier_ce(p2, wc.detach(), bc.detach(), None)
    torch.nn.functional.cross_entropy(lg, labels).backward()
    p3 = pooled.detach().float().clone().requires_grad_(True)
    torch.nn.functional.cross_entropy(p3 @ wc.detach().t() + bc.detach(), labels).backward()
    assert rel_err(p2.grad, p3.grad) < TOL[dtype]


# ------------------------------------------------------------------------------------------- vocabulary softmax-CE
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("R,V", [(37, 1002), (8, 250002), (5, 7)])
def test_vocab_cross_entropy_matches_torch(dtype, R, V):
    """The IAOG loss (CrossEntropyLoss(ignore_index=-100) over the 250 002-entry vocabulary): rows start at odd
    alignments (V % 8 == 2), some rows are ignored, gradients flow through the mean over the counted rows."""
    logits = (rnd(R, V, dtype=dtype, scale=3.0)).requires_grad_(True)
    labels = torch.randint(0, V, (R,), device=dev())
    labels[1::4] = -100
    loss = Fn.vocab_cross_entropy(logits, labels)
    ref_in = logits.detach().float().clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in, labels, ignore_index=-100)
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    (loss * 1.7).backward()
    (ref * 1.7).backward()
    assert rel_err(logits.grad, ref_in.grad) < TOL[dtype]
    assert float(logits.grad[1].abs().max()) == 0.0                   # an ignored row gets no gradient
    # 3-D call as the training loop makes it ([B, T, V] logits, [B, T] labels)
    if R % 2 == 0:
        l3 = Fn.vocab_cross_entropy(logits.detach().view(2, R // 2, V), labels.view(2, R // 2))
        assert abs(l3.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("problems", [6, 400])
def test_layernorm_bwd_compact_upstream_gradient(dtype, problems):
    """dy_every: only rows 0, E, 2E, ... carry an upstream gradient (BertPooler keeps token 0 of each problem); the kernel
    takes the compact [M/E, H] gradient and must equal the dense call with zeros in the other rows."""
    M, H, E = problems * 29, 768, 29
    x, res = rnd(M, H, dtype=dtype), rnd(M, H, dtype=dtype, seed=2)
    w, b = (1 + 0.1 * rnd(H, seed=4)), 0.1 * rnd(H, seed=5)
    _, mean, rstd = ops.ln_fwd(x, res, None, w, b)
    dyc = rnd(M // E, H, dtype=dtype, seed=7)
    dense = torch.zeros(M, H, dtype=dtype, device=dev())
    dense[::E] = dyc
    ds0, _, dg0, db0 = ops.ln_bwd_drop(dense, None, x, res, None, w, mean, rstd)
    ds1, _, dg1, db1 = ops.ln_bwd_drop(dyc, None, x, res, None, w, mean, rstd, dy_every=E)
    assert torch.equal(ds0, ds1) and rel_err(dg1, dg0) < 1e-5 and rel_err(db1, db0) < 1e-5
