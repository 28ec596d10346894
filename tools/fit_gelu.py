"""Fit of the one-MUFU erf-GELU used by the bf16 GEMM epilogues (csrc/common.cuh: gelu_erf_fast / gelu_erf_grad_fast):
    Phi(x) ~= 0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4)))
Iteratively re-weighted least squares (towards minimax) on the erf-GELU and its derivative over |x| <= 8; prints the
coefficients and the maximum absolute errors of value and derivative (with x^2 clamped at 64 as in the kernel)."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

x = np.linspace(-8, 8, 160001)
Phi = 0.5 * (1 + erf(x / np.sqrt(2)))
gelu, dgelu = x * Phi, Phi + x * np.exp(-0.5 * x * x) / np.sqrt(2 * np.pi)


def model(c, x):
    x2 = np.minimum(x * x, 64.0)
    t = np.tanh(x * (c[0] + c[1] * x2 + c[2] * x2 * x2))
    du = c[0] + 3 * c[1] * x2 + 5 * c[2] * x2 * x2
    return 0.5 * x * (1 + t), 0.5 * (1 + t) + 0.5 * x * (1 - t * t) * du


def res(c):
    g, dg = model(c, x)
    return np.concatenate([g - gelu, 0.3 * (dg - dgelu)])


c = least_squares(res, [np.sqrt(2 / np.pi), np.sqrt(2 / np.pi) * 0.044715, 0.0], method="lm", xtol=1e-15, ftol=1e-15).x
for _ in range(30):
    e = res(c)
    w = 0.2 + (np.abs(e) / np.abs(e).max()) ** 2
    c = least_squares(lambda cc: res(cc) * w, c, method="lm", xtol=1e-15, ftol=1e-15).x
g, dg = model(c, x)
print("c =", [float(v) for v in c])
print("max |gelu - fit| = %.3g   max |gelu' - fit'| = %.3g" % (np.abs(g - gelu).max(), np.abs(dg - dgelu).max()))
xx = np.linspace(-30, 30, 60001)
P = 0.5 * (1 + erf(xx / np.sqrt(2)))
g, dg = model(c, xx)
print("|x| <= 30: max |gelu - fit| = %.3g   max |gelu' - fit'| = %.3g" % (
    np.abs(g - xx * P).max(), np.abs(dg - (P + xx * np.exp(-0.5 * xx * xx) / np.sqrt(2 * np.pi))).max()))
