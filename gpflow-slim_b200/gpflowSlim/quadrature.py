"""Gauss-Hermite nodes for the likelihood expectations (reference quadrature.py:24-27)."""
import numpy as np


def hermgauss(n):
    x, w = np.polynomial.hermite.hermgauss(n)
    return x.astype(np.float64), w.astype(np.float64)
