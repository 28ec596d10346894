"""Golden fixture for the IAOG decoder hand-off: runs the UNMODIFIED reference
``IAOGDecoder`` (mm_modeling.py:634-666) in training mode on a seeded memory [B,15,H] and stores logits plus the
gradient that flows back into the fusion output. ORACLE-side tooling (test infrastructure).

    python oracle/make_golden_iaog.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("FCMF_REFERENCE", "/root/reference")
VOCAB, T, B, F = 97, 8, 2, 15


def decoder_params(hidden=768, heads=12, blocks=12, vocab=VOCAB, seed=77):
    """Reference state_dict keys of IAOGDecoder (probed), deterministic values."""
    rs = np.random.RandomState(seed)
    dh = hidden // heads

    def t(*shape, scale=0.05):
        return torch.from_numpy((rs.standard_normal(int(np.prod(shape))) * scale).astype(np.float32).reshape(shape))
    p = {"embedding.weight": t(vocab, hidden, scale=0.5), "dense.bias": t(vocab, scale=0.02)}
    for i in range(blocks):
        b = f"blks.block{i}."
        for a in ("attention1", "attention2"):
            p[b + a + ".w_kx"], p[b + a + ".w_qx"] = t(heads, hidden, dh), t(heads, hidden, dh)
            p[b + a + ".proj.weight"], p[b + a + ".proj.bias"] = t(hidden, hidden), t(hidden, scale=0.02)
        for ln in ("addnorm1.ln", "addnorm2.ln", "add_norm3.ln"):
            p[b + ln + ".weight"], p[b + ln + ".bias"] = 1 + t(hidden, scale=0.1), t(hidden, scale=0.1)
        p[b + "ffn.dense1.weight"], p[b + "ffn.dense1.bias"] = t(hidden, hidden), t(hidden, scale=0.02)
        p[b + "ffn.dense2.weight"], p[b + "ffn.dense2.bias"] = t(hidden, hidden), t(hidden, scale=0.02)
    p["dense.weight"] = p["embedding.weight"]                      # tied (mm_modeling.py:645)
    return p


def inputs(hidden=768, seed=78):
    rs = np.random.RandomState(seed)
    enc = torch.from_numpy(rs.standard_normal(B * F * hidden).astype(np.float32).reshape(B, F, hidden))
    dec_x = torch.from_numpy(rs.randint(3, VOCAB, size=(B, T)).astype(np.int64))
    labels = torch.roll(dec_x, -1, dims=1)
    labels[:, -1] = -100                                           # iaog_dataset.py:94-96
    return enc, dec_x, labels


def main():
    sys.path.insert(0, REF)
    from fcmf_framework.mm_modeling import IAOGDecoder
    torch.manual_seed(0)
    dec = IAOGDecoder(vocab_size=VOCAB).eval()
    missing, unexpected = dec.load_state_dict(decoder_params(), strict=False)
    assert not unexpected and all(k.startswith("pos_encoding") for k in missing), (missing, unexpected)
    enc, dec_x, labels = inputs()
    enc = enc.requires_grad_(True)
    mask = torch.ones(B, F, dtype=torch.int64)                     # combined_mask of fcmf_pretraining.py:184-195
    logits = dec(dec_x, [enc, mask, [None] * dec.num_blks], is_train=True)
    loss = torch.nn.CrossEntropyLoss(ignore_index=-100)(logits.permute(0, 2, 1), labels)   # run_pretraining_fcmf.py:322-324
    loss.backward()
    out = os.path.join(ROOT, "tests", "golden", "iaog_decoder.npz")
    np.savez_compressed(out, logits=logits.detach().numpy(), loss=np.float64(loss.item()), d_enc=enc.grad.numpy(),
                        g_embedding=dec.embedding.weight.grad.numpy()[:8].copy(),
                        g_wkx0=dec.blks.block0.attention2.w_kx.grad.numpy()[0, ::37].copy())
    print(f"iaog_decoder: loss={loss.item():.6f} -> {out} ({os.path.getsize(out) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
