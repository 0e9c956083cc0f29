"""Fit x*sigmoid(x*(c1 + c3 x^2 + c5 x^4)) to the erf-form GELU (minimax via iteratively re-weighted least squares).
Prints the coefficients used by gelu_erf_fast in csrc/gemm_tc_common.cuh (pre-multiplied by -log2 e)."""
import warnings
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

warnings.filterwarnings("ignore")
x = np.linspace(-9, 9, 100001)
g = x * 0.5 * (1 + erf(x / np.sqrt(2)))


def model(c, x):
    x2 = x * x
    p = (c[2] * x2 + c[1]) * x2 + c[0]
    return x / (1 + np.exp(-np.clip(p * x, -80, 80)))


c = np.array([1.5957691216, 0.0713548, 0.0])
for _ in range(40):
    err = model(c, x) - g
    w = 1 + 100 * (np.abs(err) / np.abs(err).max()) ** 6
    c = least_squares(lambda cc: (model(cc, x) - g) * w, c, xtol=1e-15, ftol=1e-15).x
print("c1, c3, c5 =", list(c), " max |err| = %.3e" % np.abs(model(c, x) - g).max())
print("-log2(e) * c =", list(-c * np.log2(np.e)))
