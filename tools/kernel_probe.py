"""Launch every hot kernel of the fusion path once at BASELINE config-2 sizes (B=64) between cudaProfilerStart/Stop.
ncu target:  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_kernels \
             python tools/kernel_probe.py
Without ncu it prints CUDA-event timings of the same launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fcmf_b200 as pkg

ops, Fn = pkg.ops, pkg.functional
L = __import__("importlib").import_module(pkg.__name__ + "._lib")
fusion = pkg.fusion
BF = torch.bfloat16
B, A, Lt, NI, NR, H, I, P = 64, 6, 170, 7, 4, 768, 3072, 49
BA, NP, S = B * A, B * A * NI, Lt + NR
M = NP * Lt
dev = "cuda"


def rnd(*s, scale=1.0, dtype=BF):
    return (torch.randn(*s, device=dev) * scale).to(dtype)


def build():
    t = {}
    t["x"], t["w_hh"], t["w_ih"], t["w_hi"] = rnd(M, H), rnd(H, H, scale=.05), rnd(I, H, scale=.05), rnd(H, I, scale=.05)
    t["g"], t["pre"] = rnd(M, I), rnd(M, I)
    t["dy"], t["dy2"] = rnd(M, H), rnd(M, H)          # distinct upstream gradients: the LayerNorm backward streams four tensors in the real step
    t["bias_h"], t["bias_i"] = torch.randn(H, device=dev) * .1, torch.randn(I, device=dev) * .1
    t["gamma"], t["beta"] = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    ix = fusion._index(B, A, Lt, NI, NR, False, torch.device(dev))
    t["ix"] = ix
    t["seq"] = rnd(BA * Lt, H)
    t["qkv_t"], t["qkv_r"] = rnd(BA * Lt, 3 * H), rnd(B * NI * NR, 3 * H)
    t["q_t"], t["kv_p"] = rnd(BA * Lt, H), rnd(B * NI * P, 2 * H)
    t["mask_add"] = torch.zeros(BA, Lt + P, device=dev)
    return t


def launches(t):
    ix = t["ix"]
    out = {}
    yield "gemm_tn K=3072 N=768 (FFN2 fwd / FFN1 dgrad)", 2.0 * M * H * I, lambda: ops.gemm_tn(t["g"], t["w_hi"], t["bias_h"], L.EPI_NONE)
    yield "gemm_tn K=768 N=3072 erf-GELU + pre (FFN1 fwd)", 2.0 * M * H * I, lambda: ops.gemm_tn(t["x"], t["w_ih"], t["bias_i"], L.EPI_GELU, want_aux=True)
    yield "gemm_tn K=768 N=3072 dGELU (FFN2 dgrad)", 2.0 * M * H * I, lambda: ops.gemm_tn(t["x"], t["w_ih"], None, L.EPI_DGELU, aux=t["pre"])
    yield "gemm_tn K=768 N=768 (attention output dense)", 2.0 * M * H * H, lambda: ops.gemm_tn(t["x"], t["w_hh"], t["bias_h"], L.EPI_NONE)
    yield "gemm_wgrad N=3072 K=768 + colsum", 2.0 * M * H * I, lambda: ops.gemm_wgrad(t["g"], t["x"])
    yield "gemm_wgrad N=768 K=768 + colsum", 2.0 * M * H * H, lambda: ops.gemm_wgrad(t["x"], t["x"])

    def ln_f():
        out["ln"] = ops.ln_fwd(t["x"], t["seq"], ix.t2i_res_idx, t["gamma"], t["beta"])
    yield "ln_fwd [456960,768] + gathered residual", 0, ln_f
    yield "ln_bwd [456960,768]", 0, lambda: ops.ln_bwd(t["dy"], t["dy2"], t["x"], t["seq"], ix.t2i_res_idx, t["gamma"], out["ln"][1], out["ln"][2])
    nh, dh = 12, 64
    plan1 = Fn.AttnPlan(NP, nh, dh, mask_div=NI).add("q", 0, 0, Lt, ix.p2ba, ix.ba2p).add("k", 1, 0, P, ix.p2bi, ix.bi2p).add("v", 1, H, P, ix.p2bi, ix.bi2p)
    plan2 = Fn.AttnPlan(NP, nh, dh, mask_div=NI)
    for role, col in (("q", 0), ("k", H), ("v", 2 * H)):
        plan2.add(role, 0, col, Lt, ix.p2ba, ix.ba2p).add(role, 1, col, NR, ix.p2bi, ix.bi2p)

    def attn(plan, tensors, Lq, Lk, key):
        d = Fn._desc(plan, tensors, t["mask_add"], None)
        ctx, lse = ops.attn_fwd(d, Lq, BF, torch.device(dev))
        out[key] = (d, ctx, lse)

    def attn_b(key, Lq, Lk):
        d, ctx, lse = out[key]
        ops.attn_bwd(d, Lq, Lk, ctx, ctx, lse, False)
    yield "attn fwd text->image (Lq=170, Lk=49)", 4.0 * NP * nh * Lt * P * dh, lambda: attn(plan1, (t["q_t"], t["kv_p"]), Lt, P, "a1")
    yield "attn bwd text->image (dQ + dK/dV)", 0, lambda: attn_b("a1", Lt, P)
    yield "attn fwd text+ROI (L=174)", 4.0 * NP * nh * S * S * dh, lambda: attn(plan2, (t["qkv_t"], t["qkv_r"]), S, S, "a2")
    yield "attn bwd text+ROI (dQ + dK/dV)", 0, lambda: attn_b("a2", S, S)
    yield "gather_sum_rows [65280 x 7 -> 768]", 0, lambda: ops.gather_sum_rows(t["x"], ix.t2i_res_inv, BA * Lt, NI)
    # the same kernels with in-kernel dropout (train() mode, p = 0.1)
    drop = ops.Drop(0.1, 12345)

    def ln_fd():
        out["lnd"] = ops.ln_fwd(t["x"], t["seq"], ix.t2i_res_idx, t["gamma"], t["beta"], drop=drop)
    yield "ln_fwd + dropout", 0, ln_fd
    yield "ln_bwd + dropout (two outputs)", 0, lambda: ops.ln_bwd_drop(t["dy"], t["dy2"], t["x"], t["seq"], ix.t2i_res_idx, t["gamma"], out["lnd"][1], out["lnd"][2], drop)
    plan1d = Fn.AttnPlan(NP, nh, dh, mask_div=NI, drop=drop).add("q", 0, 0, Lt, ix.p2ba, ix.ba2p).add("k", 1, 0, P, ix.p2bi, ix.bi2p).add("v", 1, H, P, ix.p2bi, ix.bi2p)
    plan2d = Fn.AttnPlan(NP, nh, dh, mask_div=NI, drop=drop)
    for role, col in (("q", 0), ("k", H), ("v", 2 * H)):
        plan2d.add(role, 0, col, Lt, ix.p2ba, ix.ba2p).add(role, 1, col, NR, ix.p2bi, ix.bi2p)
    yield "attn fwd text->image + dropout", 0, lambda: attn(plan1d, (t["q_t"], t["kv_p"]), Lt, P, "a1d")
    yield "attn bwd text->image + dropout", 0, lambda: attn_b("a1d", Lt, P)
    yield "attn fwd text+ROI + dropout", 0, lambda: attn(plan2d, (t["qkv_t"], t["qkv_r"]), S, S, "a2d")
    yield "attn bwd text+ROI + dropout", 0, lambda: attn_b("a2d", S, S)


if __name__ == "__main__":
    t = build()
    for _, _, fn in launches(t):          # warm-up (lazy attribute setting, allocator)
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    rows = []
    for name, flops, fn in launches(t):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        rows.append([name, flops, [(e0, e1)]])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    reps = int(os.environ.get("FCMF_PROBE_REPS", "1"))          # extra un-profiled repetitions: report the fastest
    for _ in range(reps - 1):
        for row, (_, _, fn) in zip(rows, launches(t)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            row[2].append((e0, e1))
    torch.cuda.synchronize()
    for name, flops, evs in rows:
        ms = min(a.elapsed_time(b) for a, b in evs)
        print(f"{ms:8.3f} ms  {(flops / ms / 1e9 if flops else 0):8.1f} TFLOP/s  {name}")
