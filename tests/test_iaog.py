"""IAOG decoder hand-off: the oracle's decoder restatement is pinned against the
reference's own output (tests/golden/iaog_decoder.npz), and the kernel-backed decoder is compared with both."""
import os
import sys

import numpy as np
import pytest
import torch

from _util import GOLD, ROOT, pkg, rel_err
from oracle import fcmf_oracle as O

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import make_golden_iaog as G                                                   # noqa: E402


def _loss(logits, labels):
    return torch.nn.functional.cross_entropy(logits.permute(0, 2, 1), labels, ignore_index=-100)


def test_oracle_decoder_matches_reference_golden():
    z = np.load(os.path.join(GOLD, "iaog_decoder.npz"))
    params = {"decoder." + k: v.clone().requires_grad_(True) for k, v in G.decoder_params().items()}
    params["decoder.dense.weight"] = params["decoder.embedding.weight"]
    enc, dec_x, labels = G.inputs()
    enc = enc.requires_grad_(True)
    logits = O.iaog_decoder(dec_x, enc, params, num_blocks=12)
    loss = _loss(logits, labels)
    loss.backward()
    assert rel_err(logits, torch.from_numpy(z["logits"])) < 2e-5
    assert abs(loss.item() - float(z["loss"])) < 2e-5 * abs(float(z["loss"]))
    assert rel_err(enc.grad, torch.from_numpy(z["d_enc"])) < 5e-5
    assert rel_err(params["decoder.embedding.weight"].grad[:8], torch.from_numpy(z["g_embedding"])) < 5e-5


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
def test_kernel_decoder_matches_reference_golden(dtype, tol):
    z = np.load(os.path.join(GOLD, "iaog_decoder.npz"))
    iaog = pkg("iaog")
    dec = iaog.IAOGDecoder(vocab_size=G.VOCAB)
    missing, unexpected = dec.load_state_dict(G.decoder_params(), strict=False)
    assert not unexpected and all(k.startswith("pos_encoding") for k in missing)
    dec = dec.cuda().eval()
    dec.compute_dtype = dtype
    enc, dec_x, labels = G.inputs()
    enc = enc.cuda().requires_grad_(True)
    mask = torch.ones(enc.shape[0], enc.shape[1], dtype=torch.int64, device="cuda")
    logits = dec(dec_x.cuda(), [enc.to(dtype), mask, [None] * dec.num_blks], is_train=True)
    loss = pkg().FCMFSeq2Seq.loss(logits, labels.cuda())          # the vocabulary softmax-CE kernels (fwd + bwd)
    assert abs(loss.item() - float(z["loss"])) < (2e-4 if dtype == torch.float32 else 3e-2) * abs(float(z["loss"]))
    loss.backward()
    assert rel_err(logits, torch.from_numpy(z["logits"])) < tol
    assert rel_err(enc.grad, torch.from_numpy(z["d_enc"])) < (tol if dtype == torch.float32 else 0.15)
    assert rel_err(dec.embedding.weight.grad[:8], torch.from_numpy(z["g_embedding"])) < (tol if dtype == torch.float32 else 0.15)
