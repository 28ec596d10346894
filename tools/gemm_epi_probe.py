"""Where does the time of the N=3072 / K=768 GEMMs go? Same main loop, different epilogues (CUDA-event timings):
plain + bias, GELU, GELU + saved pre-activation, dGELU, against a pure write and a copy of the same output bytes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fcmf_b200 as pkg

ops = pkg.ops
L = __import__("importlib").import_module(pkg.__name__ + "._lib")
BF = torch.bfloat16
M, H, I = 64 * 6 * 7 * 170, 768, 3072
dev = "cuda"
x = (torch.randn(M, H, device=dev)).to(BF)
w = (torch.randn(I, H, device=dev) * .05).to(BF)
bias = torch.randn(I, device=dev) * .1
pre = torch.randn(M, I, device=dev).to(BF)
out = torch.empty(M, I, dtype=BF, device=dev)
aux = torch.empty(M, I, dtype=BF, device=dev)


def timed(fn, reps=int(os.environ.get("FCMF_PROBE_REPS", "5"))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


flops = 2.0 * M * H * I
cases = [
    ("plain + bias", lambda: ops.gemm_tn(x, w, bias, L.EPI_NONE, out=out)),
    ("no bias", lambda: ops.gemm_tn(x, w, None, L.EPI_NONE, out=out)),
    ("tanh", lambda: ops.gemm_tn(x, w, bias, L.EPI_TANH, out=out)),
    ("GELU", lambda: ops.gemm_tn(x, w, bias, L.EPI_GELU, out=out)),
    ("GELU + pre", lambda: ops.gemm_tn(x, w, bias, L.EPI_GELU, out=out, aux=aux, want_aux=True)),
    ("dGELU", lambda: ops.gemm_tn(x, w, None, L.EPI_DGELU, out=out, aux=pre)),
]
for name, fn in cases:
    ms = timed(fn)
    print(f"{ms:8.3f} ms {flops / ms / 1e9:8.1f} TFLOP/s  {name}")
ms = timed(lambda: out.zero_())
print(f"{ms:8.3f} ms {out.numel() * 2 / ms / 1e6:8.1f} GB/s  memset of one [M, 3072] bf16 output")
ms = timed(lambda: out.copy_(pre))
print(f"{ms:8.3f} ms {2 * out.numel() * 2 / ms / 1e6:8.1f} GB/s  copy of one [M, 3072] bf16 tensor (read + write)")
