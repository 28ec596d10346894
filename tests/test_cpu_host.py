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
    assert lib.fcmf_abi_version() == 1


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
