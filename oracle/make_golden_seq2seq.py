"""Golden fixture for the WHOLE ``FCMFSeq2Seq.forward``: the UNMODIFIED reference
wrapper (fcmf_pretraining.py:144-207) -- FCMFEncoder -> ``combined_mask`` / ``dec_state`` hand-off -> IAOGDecoder -> the
pre-training loss (run_pretraining_fcmf.py:320-324) -- run in this container on seeded inputs, with the text encoder
stubbed by a leaf ``sequence_output`` exactly as in make_golden.py. ORACLE-side tooling (test infrastructure).

    python oracle/make_golden_seq2seq.py

The vocabulary size is deliberately NOT a multiple of 8 (1003 = 3 mod 8, like the real 250 002 = 2 mod 8) so that the
tensor-core path of the vocabulary projection is exercised with the same alignment class as the real configuration;
target length T = 32 as in BASELINE config 4."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
REF = os.environ.get("FCMF_REFERENCE", "/root/reference")

import importlib                                                   # noqa: E402
import make_golden_iaog as GI                                      # noqa: E402

synth = importlib.import_module("multimodal-aspect-category-sentiment-analysis_b200.synth")

VOCAB, T = 1003, 32
DIMS = dict(batch=2, aspects=1, seq_len=24, num_imgs=7, num_roi=4)
PARAM_SEED, BATCH_SEED, DEC_SEED, TOK_SEED = 46, 1238, 79, 80
STRIDE = 997


def seq2seq_params():
    """Reference-keyed state_dict of FCMFSeq2Seq without the text encoder: fusion parameters + decoder parameters."""
    dims = synth.FusionDims(**DIMS)
    p = {k: v for k, v in synth.make_params(dims, seed=PARAM_SEED, with_head=False).items()}
    for k, v in GI.decoder_params(vocab=VOCAB, seed=DEC_SEED).items():
        p["decoder." + k] = v
    return p, dims


def inputs(dims):
    batch = synth.make_batch(dims, seed=BATCH_SEED, mask="bernoulli")
    rs = np.random.RandomState(TOK_SEED)
    dec_x = torch.from_numpy(rs.randint(3, VOCAB, size=(dims.batch, T)).astype(np.int64))
    labels = torch.roll(dec_x, -1, dims=1)
    labels[:, -1] = -100                                            # iaog_dataset.py:94-96
    attn = torch.ones(dims.batch, dims.seq_len, dtype=torch.int64)  # text attention_mask (only its first column reaches combined_mask)
    return batch, dec_x, labels, attn


def sample(t):
    f = t.detach().reshape(-1)
    return (f if f.numel() <= 4096 else f[::STRIDE]).numpy().copy()


def main():
    sys.path.insert(0, REF)
    from fcmf_framework.fcmf_pretraining import FCMFSeq2Seq
    from transformers import XLMRobertaConfig, XLMRobertaModel
    torch.manual_seed(0)
    params, dims = seq2seq_params()
    with tempfile.TemporaryDirectory() as d:       # hidden 768 so that the tied embedding has the decoder's width; replaced by a stub below
        XLMRobertaModel(XLMRobertaConfig(vocab_size=64, hidden_size=768, num_hidden_layers=1, num_attention_heads=12,
                                         intermediate_size=64, max_position_embeddings=40, type_vocab_size=1,
                                         pad_token_id=1)).save_pretrained(d)
        model = FCMFSeq2Seq(VOCAB, T, d, dims.num_imgs, dims.num_roi, 0.7).eval()

    class StubText(torch.nn.Module):
        def forward(self, input_ids, token_type_ids, attention_mask):
            return input_ids, None, None
    model.encoder.bert = StubText()
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    assert all("pos_encoding" in k for k in missing), missing

    batch, dec_x, labels, attn = inputs(dims)
    seq = batch["sequence_output"][:, 0].clone().requires_grad_(True)
    logits = model(seq, dec_x, batch["visual_embeds_att"], batch["roi_embeds_att"], batch["roi_coors"], None, attn,
                   batch["added_attention_mask"][:, 0], None, True)
    loss = torch.nn.CrossEntropyLoss(ignore_index=-100)(logits.permute(0, 2, 1), labels)    # run_pretraining_fcmf.py:322-324
    loss.backward()
    out = {"logits": logits.detach().numpy(), "loss": np.float64(loss.item()), "d_sequence_output": sample(seq.grad),
           "sample_stride": np.int64(STRIDE)}
    for k, v in model.named_parameters():
        if v.grad is None:
            continue
        out["gsample/" + k] = sample(v.grad)
    path = os.path.join(ROOT, "tests", "golden", "seq2seq_t32.npz")
    np.savez_compressed(path, **out)
    print(f"seq2seq_t32: loss={loss.item():.6f} logits {tuple(logits.shape)} -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
