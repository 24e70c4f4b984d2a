"""Log densities on the hot path (reference densities.py:24-25, :73-95)."""
import numpy as np
import torch

from ._backend import ops as _ops

LOG2PI = float(np.log(2 * np.pi))


def gaussian(x, mu, var):
    """densities.py:24-25."""
    return -0.5 * (LOG2PI + torch.log(var) + (mu - x) ** 2 / var)


def multivariate_normal(x, mu, L):
    """log N(x | mu, L L^T), columns independent (densities.py:73-95):
    alpha = L^-1 (x - mu) (TRSM kernel); -N C/2 log 2pi - C sum log L_ii - 1/2 sum alpha^2."""
    d = x - mu
    vec = d.dim() == 1
    if vec:
        d = d.reshape(-1, 1)
    alpha_t = _ops.trsm_rlt(_ops.t(d), L)          # (L^-1 d)^T = d^T L^-T
    num_col = 1 if vec else x.shape[1]
    num_dims = x.shape[0]
    ret = -0.5 * num_dims * num_col * LOG2PI
    ret = ret - num_col * torch.log(torch.diagonal(L)).sum()
    ret = ret - 0.5 * (alpha_t ** 2).sum()
    return ret
