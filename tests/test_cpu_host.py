"""CPU-side checks: the C-ABI library loads and exports every symbol include/fcmf_b200.h declares, the host mirror
keeps the reference's state_dict contract, the fold/hoist index tables are right, and the product path fails
loudly without CUDA. No compute calls are made here."""
import ctypes
import os
import re

import pytest
import torch

from _util import ROOT, pkg, synth


def test_library_builds_and_exports_every_declared_symbol():
    p = pkg()
    path = p.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "fcmf_b200.h")).read()
    declared = set(re.findall(r"\b(fcmf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/fcmf_b200.h but not exported"
    assert declared == set(pkg("_lib").exported_symbols())
    assert lib.fcmf_abi_version() == pkg("_lib").ABI_VERSION == int(re.search(r"#define FCMF_ABI_VERSION (\d+)", header).group(1))


def test_state_dict_keys_match_reference_contract():
    dims = synth.FusionDims()
    model = pkg().FCMF(None, num_labels=4, num_imgs=7, num_roi=4)
    want = dict(synth.fusion_param_spec(dims))
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == want
    model.load_state_dict(synth.make_params(dims), strict=True)


def test_fold_index_tables():
    fusion = pkg("fusion")
    B, A, L, NI, NR = 2, 3, 5, 2, 2
    ix = fusion._Index(B, A, L, NI, NR, False, torch.device("cpu"))
    NP, S = B * A * NI, L + NR
    for p in range(NP):
        ba, i = divmod(p, NI)
        b = ba // A
        assert ix.p2ba[p] == ba and ix.p2bi[p] == b * NI + i
        assert p in ix.ba2p[ba].tolist() and p in ix.bi2p[b * NI + i].tolist()
    # residual row of every text+ROI row, and its inverse
    for m in range(NP * S):
        p, s = divmod(m, S)
        ba, i = divmod(p, NI)
        want = ba * L + s if s < L else B * A * L + ((ba // A) * NI + i) * NR + (s - L)
        assert ix.roi_res_idx[m] == want
        assert m in ix.roi_res_inv[want].tolist()
    assert sorted(x for x in ix.roi_res_inv.reshape(-1).tolist() if x >= 0) == list(range(NP * S))
    live = fusion._Index(B, A, L, NI, NR, True, torch.device("cpu"))
    assert live.t2i_res_idx.tolist() == [(p // NI) * L for p in range(NP)]
    assert sorted(x for x in live.t2i_res_inv.reshape(-1).tolist() if x >= 0) == list(range(NP))


def test_product_path_refuses_cpu_tensors():
    ops = pkg("ops")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm_tn(torch.zeros(4, 8), torch.zeros(4, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.ln_fwd(torch.zeros(4, 8), None, None, torch.ones(8), torch.zeros(8))


def test_flop_accounting_matches_survey_table():
    d = synth.FusionDims()
    assert abs(synth.flops_forward_per_sample(d, "exec") / 1e9 - 206.70) < 0.3
    assert abs(synth.flops_forward_per_sample(d, "full") / 1e9 - 188.93) < 0.3
    assert abs(synth.flops_forward_per_sample(d, "live") / 1e9 - 5.91) < 0.3


def test_dropout_mask_restatement_matches_the_library():
    """oracle/dropout_mask.py (numpy) == the inline functions the kernels use, evaluated on the host by the library;
    and the mask has the statistics of a Bernoulli(1-p) field."""
    import numpy as np
    from oracle import dropout_mask as DM
    lib = pkg("_lib").load()
    rng = np.random.RandomState(0)
    for p in (0.1, 0.5, 0.013):
        for _ in range(3):
            seed = int(rng.randint(0, 2 ** 62, dtype=np.int64)) * 4 + int(rng.randint(0, 4))
            rows = np.concatenate([rng.randint(0, 2 ** 40, size=5, dtype=np.int64), np.arange(3)])
            m = DM.keep_mask(seed, rows, 37, p)
            for ri, r in enumerate(rows):
                for c in range(37):
                    assert bool(lib.fcmf_dropout_keep(p, seed, int(r), c)) == bool(m[ri, c])
    big = DM.keep_mask(12345, np.arange(4096), 768, 0.1)
    assert abs(big.mean() - 0.9) < 2e-3
    assert abs(big.mean(0) - 0.9).max() < 0.03 and abs(big.mean(1) - 0.9).max() < 0.06     # no dead rows / columns
    a, b = big[:, 0::2], big[:, 1::2]                                                    # the two halves of a pair hash
    assert abs(np.corrcoef(a.ravel(), b.ravel())[0, 1]) < 5e-3
    assert abs(np.corrcoef(big[:-1].ravel(), big[1:].ravel())[0, 1]) < 5e-3              # adjacent rows
    # consecutive seeds (CUDA-graph replays advance the device counter by one) give unrelated masks, not row permutations
    m0, m1 = DM.keep_mask(777, np.arange(256), 768, 0.1), DM.keep_mask(778, np.arange(256), 768, 0.1)
    rows0 = {r.tobytes() for r in m0}
    assert sum(r.tobytes() in rows0 for r in m1) == 0
    assert abs(np.corrcoef(m0.ravel(), m1.ravel())[0, 1]) < 1e-2
    thr, inv = DM.threshold(0.1)
    assert thr == 6554 and abs(inv - 1.0 / (1.0 - 6554 / 65536)) < 1e-6


def test_ctypes_struct_layouts_match_the_compiled_library():
    """The by-value structs of the ABI (fcmf_dropout, fcmf_seg, fcmf_attn_desc) as mirrored in _lib.py have the sizes and
    field offsets the compiled library uses (a silent mismatch would corrupt every attention launch)."""
    L = pkg("_lib")
    lib = L.load()
    want = [ctypes.sizeof(L.Dropout), L.Dropout.seed.offset, L.Dropout.seed_dev.offset, ctypes.sizeof(L.Seg), L.Seg.idx.offset,
            ctypes.sizeof(L.AttnDesc), L.AttnDesc.mask_add.offset, L.AttnDesc.bias.offset, L.AttnDesc.scale.offset,
            L.AttnDesc.causal.offset, L.AttnDesc.drop.offset, L.AttnDesc.engine.offset, L.Seg.groups.offset]
    got = [lib.fcmf_abi_layout(i) for i in range(len(want))]
    assert got == want, (got, want)
    assert lib.fcmf_abi_layout(99) == -1


def test_attention_gradient_row_tables():
    """functional._identity_rows / _group_rows: the per-problem gradient rows that gather_sum_rows reduces into the rows of
    a shared tensor (text rows shared by the images of a (sample, aspect); patch / ROI rows shared by the aspects)."""
    Fn = pkg("functional")
    dev = torch.device("cpu")
    NP, L, off, rows = 6, 5, 2, 3
    ident = Fn._identity_rows(NP, L, off, rows, dev)
    assert ident.shape == (NP * rows, 1)
    assert ident.view(NP, rows).tolist() == [[p * L + off + r for r in range(rows)] for p in range(NP)]
    inv = torch.tensor([[0, 2, -1], [1, 3, 5]], dtype=torch.int32)              # 2 groups, problems of each (-1 padded)
    grp = Fn._group_rows(inv, L, off, rows)
    assert grp.shape == (2 * rows, 3)
    for g in range(2):
        for r in range(rows):
            want = [(-1 if p < 0 else p * L + off + r) for p in inv[g].tolist()]
            assert grp[g * rows + r].tolist() == want


def test_dropout_site_seeds_and_plan_objects():
    Fn, ops, fusion = pkg("functional"), pkg("ops"), pkg("fusion")
    from oracle import dropout_mask as DM
    assert Fn.site_drop(None, 3, 0.1) is None and Fn.site_drop(5, 3, 0.0) is None          # eval / p = 0: no Drop object
    d = Fn.site_drop(5, fusion.DROP_SITES["mm_attn"], 0.1)
    assert isinstance(d, ops.Drop) and d.seed == DM.site_seed(5, fusion.DROP_SITES["mm_attn"]) and d.struct().p == pytest.approx(0.1)
    torch.manual_seed(1)
    a = Fn.new_step_seed()
    torch.manual_seed(1)
    assert Fn.new_step_seed() == a and 0 <= a < 2 ** 62
    with pytest.raises(ValueError):
        ops.Drop(1.0, 1)
    plan = Fn.AttnPlan(4, 2, 64, drop=d).add("q", 0, 0, 7, None, None).add("k", 0, 128, 7, None, None).add("k", 1, 0, 3, None, None)
    assert plan.Lq == 7 and plan.Lk == 10 and plan.drop is d


def test_weight_gradient_split_k_fills_the_last_wave():
    """The split-K factor of the tcgen05 weight-gradient GEMM is chosen for the fill of the LAST wave of (tile, split)
    units (round 1: ceil(workers / tiles) left 27-45 % of the kernel idle). Every weight-gradient shape of the config-2
    step must run at >= 80 % wave efficiency. Host-only planning query: no GPU needed (148 SMs assumed without one)."""
    lib = pkg("_lib").load()
    i32 = ctypes.c_int32
    shapes = [(456960, 3072, 768), (456960, 768, 3072), (456960, 768, 768), (467712, 768, 768), (65280, 2304, 768),
              (21952, 768, 2048), (21952, 1536, 768), (65280, 768, 768), (5760, 768, 768), (1792, 2304, 768)]
    for M, N, K in shapes:
        pair, tiles, splits, workers = i32(), i32(), i32(), i32()
        assert lib.fcmf_gemm_wgrad_plan(M, N, K, ctypes.byref(pair), ctypes.byref(tiles), ctypes.byref(splits), ctypes.byref(workers)) == 0
        units = tiles.value * splits.value
        waves = -(-units // workers.value)
        assert 1 <= splits.value <= 32
        if M >= 5000:      # long reductions: a split's extra atomic pass is noise, the last wave's fill is what counts
            assert units / (waves * workers.value) >= 0.8, (M, N, K, tiles.value, splits.value)
    # short reduction, huge output (the vocabulary projection's weight gradient, 32 k-blocks, 768 MB of fp32): never split --
    # three splits measured 190 TFLOP/s behind their atomics
    pair, tiles, splits, workers = i32(), i32(), i32(), i32()
    assert lib.fcmf_gemm_wgrad_plan(2048, 250112, 768, ctypes.byref(pair), ctypes.byref(tiles), ctypes.byref(splits), ctypes.byref(workers)) == 0
    assert splits.value == 1 and tiles.value == 2931
    assert lib.fcmf_gemm_wgrad_plan(0, 1, 1, None, None, None, None) != 0


def test_gelu_tanh_fit_error_bounds():
    """The one-MUFU GELU of the bf16 GEMM epilogues (csrc/common.cuh: gelu_erf_fast / gelu_erf_grad_fast), evaluated in
    float32 numpy with the constants parsed from the header: the documented maximum errors against the exact erf-GELU hold."""
    import numpy as np
    from scipy.special import erf
    src = open(os.path.join(ROOT, "multimodal-aspect-category-sentiment-analysis_b200", "csrc", "common.cuh")).read()
    c0, c1, c2 = [np.float32(float(re.search(rf"kGeluC{i} = (-?[0-9.eE+-]+)f", src).group(1))) for i in range(3)]
    x = np.linspace(-12, 12, 480001).astype(np.float32)
    x2 = np.minimum(x * x, np.float32(64.0))
    t = np.tanh((x * ((c2 * x2 + c1) * x2 + c0)).astype(np.float64)).astype(np.float32)
    g = np.float32(0.5) * x * t + np.float32(0.5) * x
    du = (np.float32(5.0) * c2 * x2 + np.float32(3.0) * c1) * x2 + c0
    dg = np.float32(0.5) * x * (np.float32(1.0) - t * t) * du + (np.float32(0.5) * t + np.float32(0.5))
    xd = x.astype(np.float64)
    Phi = 0.5 * (1 + erf(xd / np.sqrt(2)))
    assert np.abs(g - xd * Phi).max() < 7e-5
    assert np.abs(dg - (Phi + xd * np.exp(-0.5 * xd * xd) / np.sqrt(2 * np.pi))).max() < 2e-4


def test_nvtx_tracing_hook_is_off_by_default_and_harmless():
    L = pkg("_lib")
    assert L.NVTX is False or os.environ.get("FCMF_NVTX", "0") not in ("", "0")
    with L.trace("unit-test range"):
        pass
    old = L.NVTX
    try:
        L.NVTX = True
        with L.trace("unit-test range (on)"):            # torch.cuda.nvtx is a no-op without a profiler attached
            pass
        assert L.load().fcmf_abi_version() == L.ABI_VERSION
    finally:
        L.NVTX = old
