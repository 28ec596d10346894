"""The vocabulary projection's three GEMMs at BASELINE configs[3] size (B*T = 2048 rows, V = 250 002 -> 250 112, H = 768).
ncu target: ncu --set full --clock-control none -k regex:gemm_tc -o gpurun_out/x python tools/vocab_wgrad_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fcmf_b200 as pkg
ops = pkg.ops
M, Vp, H = 2048, 250112, 768
dy = (torch.randn(M, Vp, device="cuda") * 0.01).bfloat16()
x = torch.randn(M, H, device="cuda").bfloat16()
w = (torch.randn(Vp, H, device="cuda") * 0.02).bfloat16()
wt = w.t().contiguous()
for name, fn, fl in (("wgrad", lambda: ops.gemm_wgrad(dy, x, want_bias=False), 2.0 * M * Vp * H),
                     ("wgrad+colsum", lambda: ops.gemm_wgrad(dy, x, want_bias=True), 2.0 * M * Vp * H),
                     ("fwd", lambda: ops.gemm_tn(x, w), 2.0 * M * Vp * H),
                     ("dgrad splitK", lambda: ops.gemm_tn_f32(dy, wt), 2.0 * M * Vp * H)):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name:14s} {ms:8.3f} ms {fl / ms / 1e9:8.1f} TFLOP/s")
