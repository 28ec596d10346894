"""Time individual tcgen05 GEMM shapes/epilogues with CUDA events (and serve as a small ncu target).
    python tools/gemm_probe.py [M N K [epi]]         epi in none|gelu|gelu_aux|dgelu|tanh|wgrad"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fcmf_b200 as pkg

ops, L = pkg.ops, __import__("importlib").import_module(pkg.__name__ + "._lib")


def run(M, N, K, epi, iters=5):
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda") * 0.1
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    aux = torch.randn(M, N, device="cuda").bfloat16() if epi in ("gelu_aux", "dgelu") else None
    dy = torch.randn(M, N, device="cuda").bfloat16() if epi == "wgrad" else None
    dw = torch.empty(N, K, device="cuda") if epi == "wgrad" else None

    def call():
        if epi == "wgrad":
            ops.gemm_wgrad(dy, a, want_bias=False, engine=L.ENGINE_TCGEN05, dw=dw)
        elif epi == "none":
            ops.gemm_tn(a, b, bias, L.EPI_NONE, out=out, engine=L.ENGINE_TCGEN05)
        elif epi == "gelu":
            ops.gemm_tn(a, b, bias, L.EPI_GELU, out=out, engine=L.ENGINE_TCGEN05)
        elif epi == "gelu_aux":
            ops.gemm_tn(a, b, bias, L.EPI_GELU, out=out, aux=aux, engine=L.ENGINE_TCGEN05)
        elif epi == "dgelu":
            ops.gemm_tn(a, b, None, L.EPI_DGELU, out=out, aux=aux, engine=L.ENGINE_TCGEN05)
        elif epi == "tanh":
            ops.gemm_tn(a, b, bias, L.EPI_TANH, out=out, engine=L.ENGINE_TCGEN05)
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"M={M} N={N} K={K} {epi:9s} {ms:8.3f} ms  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) >= 4:
        run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else "none")
    else:
        for epi in ("none", "gelu", "gelu_aux", "dgelu"):
            run(456960, 3072, 768, epi)
        run(456960, 768, 3072, "none")
        run(456960, 768, 768, "none")
        run(456960, 768, 3072, "wgrad")
