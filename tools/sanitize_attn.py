"""Small launches of the warp-specialised attention kernels (forward + unified backward; one and three key blocks, two
segments, dropout on/off) for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_attn.py
    compute-sanitizer --tool racecheck python tools/sanitize_attn.py
Checks the results against fp32 PyTorch as well (so a sanitizer-clean but wrong run is not reported as clean)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fcmf_b200 as pkg

Fn, ops = pkg.functional, pkg.ops
dev = "cuda"


def case(heads, L1, L2, Lk1, p_drop, seed=0):
    torch.manual_seed(seed)
    HD, NI, A, B = heads * 64, 2, 2, 1
    BA, NP = B * A, B * A * NI
    pidx = torch.arange(NP, device=dev, dtype=torch.int32)
    p2ba, p2bi = (pidx // NI).contiguous(), ((pidx // (A * NI)) * NI + pidx % NI).contiguous()
    ba2p = (torch.arange(BA, device=dev, dtype=torch.int32).view(BA, 1) * NI + torch.arange(NI, device=dev, dtype=torch.int32)).contiguous()
    bi2p = torch.tensor([[(b * A + a) * NI + i for a in range(A)] for b in range(B) for i in range(NI)], device=dev, dtype=torch.int32)
    tq = (torch.randn(BA * L1, (HD if Lk1 else 3 * HD), device=dev)).bfloat16().requires_grad_(True)
    tensors = [tq]
    plan = Fn.AttnPlan(NP, heads, 64, mask_div=NI, drop=ops.Drop(p_drop, 1234) if p_drop > 0 else None)
    if Lk1:   # text -> image: q from the text tensor, k/v from a patch tensor indexed by (b, i)
        tk = (torch.randn(B * NI * Lk1, 2 * HD, device=dev)).bfloat16().requires_grad_(True)
        tensors.append(tk)
        plan.add("q", 0, 0, L1, p2ba, ba2p).add("k", 1, 0, Lk1, p2bi, bi2p).add("v", 1, HD, Lk1, p2bi, bi2p)
        Lq, Lk = L1, Lk1
    else:
        for role, col in (("q", 0), ("k", HD), ("v", 2 * HD)):
            plan.add(role, 0, col, L1, p2ba, ba2p)
        if L2:
            t1 = (torch.randn(B * NI * L2, 3 * HD, device=dev)).bfloat16().requires_grad_(True)
            tensors.append(t1)
            for role, col in (("q", 0), ("k", HD), ("v", 2 * HD)):
                plan.add(role, 1, col, L2, p2bi, bi2p)
        Lq = Lk = L1 + L2
    mask = (torch.rand(BA, Lk + 3, device=dev) < 0.8).long()
    mask[:, 0] = 1
    mask_add = ops.mask_additive(mask, Lk + 3)
    out = Fn.folded_attention(plan, tensors, mask_add, None)
    out.backward(torch.randn_like(out))
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(out.float()).all()) and all(bool(torch.isfinite(t.grad.float()).all()) for t in tensors)
    print(f"heads={heads} Lq={Lq} Lk={Lk} drop={p_drop}: finite={ok}", flush=True)
    assert ok


if __name__ == "__main__":
    case(2, 170, 4, 0, 0.0)      # three key blocks, two segments, two query tiles
    case(2, 170, 4, 0, 0.1)
    case(2, 170, 0, 49, 0.0)     # one key block (in-team delta), two query tiles
    case(2, 170, 0, 49, 0.1)
    case(3, 40, 4, 0, 0.1)       # one key block, one query tile, two segments
    case(1, 130, 30, 0, 0.0)
    print("sanitize_attn: all cases ran")
