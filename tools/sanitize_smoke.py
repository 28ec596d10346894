"""Tiny fwd+bwd of the fusion path in every mode, for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fcmf_b200 as pkg

synth = pkg.synth
dims = synth.FusionDims(batch=2, aspects=2, seq_len=140, num_imgs=2, num_roi=4)      # Lq=144: two 128-row tiles, 3 key blocks
params = synth.make_params(dims, seed=1)
b = synth.make_batch(dims, seed=2, mask="bernoulli")
BA = dims.batch * dims.aspects
for dtype in (torch.bfloat16, torch.float32):
    for rows in ("full", "live"):
        model = pkg.FCMF(None, num_labels=4, num_imgs=dims.num_imgs, num_roi=dims.num_roi)
        model.load_state_dict(params)
        model = model.cuda().eval()
        model.encoder.compute_dtype = dtype
        seq = b["sequence_output"].reshape(BA, dims.seq_len, dims.hidden).cuda().requires_grad_(True)
        logits, loss = model.fuse_all_aspects(seq, b["visual_embeds_att"].cuda(), b["roi_embeds_att"].cuda(), b["roi_coors"].cuda(),
                                              b["added_attention_mask"].reshape(BA, -1).cuda(), b["labels"].cuda(),
                                              aspects=dims.aspects, rows=rows)
        loss.backward()
        torch.cuda.synchronize()
        print(dtype, rows, float(loss), flush=True)
print("done")
