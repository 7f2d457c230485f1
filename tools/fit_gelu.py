"""Fits the inner polynomial of x*sigmoid(2u), u = x(a + b x^2 + c x^4), to the ERF-form GELU (minimax by
iteratively re-weighted least squares) and prints the coefficients used by csrc/gemm.cu (gelu_erf)."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

x = np.linspace(-10, 10, 200001)
ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))


def g(p, x):
    a, b, c = p
    return 0.5 * x * (1 + np.tanh(x * (a + x * x * (b + c * x * x))))


p = np.array([np.sqrt(2 / np.pi), np.sqrt(2 / np.pi) * 0.044715, 0.0])
w = np.ones_like(x)
for _ in range(60):
    p = least_squares(lambda q: (g(q, x) - ref) * w, p).x
    e = np.abs(g(p, x) - ref)
    w = w * (1 + 4 * e / e.max())
    w /= w.mean()
k = -2 * np.log2(np.e)
print("a, b, c =", p, " max |err| =", np.abs(g(p, x) - ref).max())
print("folded with -2*log2(e):", p * k)
