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


def multivariate_normal_feature(x, mu, C, var):
    """log N(x | mu, C C^T + var I) through the Woodbury identity: only the small
    [n_small, n_small] matrix C^T C + var I is factored (densities.py:98-123; the reference tests
    it against `multivariate_normal`, :159-174).  C is [n_big, n_small].  Op for op as the
    reference, including its `+ jitter` inside the log-determinant and its treatment of a
    multi-column x as one long vector."""
    from ._settings import SETTINGS as settings
    x = x - mu
    dim1, dim2 = C.shape[0], C.shape[1]
    Ct = _ops.t(C)
    CtC = _ops.matmul_nt(Ct, Ct) + var * torch.eye(dim2, dtype=C.dtype, device=C.device)
    L = _ops.cholesky(CtC)
    logdet_L = torch.log(torch.diagonal(L) + settings.jitter).sum()
    logdet_CCt = 2. * logdet_L + float(dim1 - dim2) * torch.log(var)
    x_norm = (x ** 2).sum()
    Ctx = _ops.matmul_nt(Ct, _ops.t(x))                # C^T x  [n_small, R]
    L_inv_x = _ops.solve_lower(L, Ctx)
    square_dist = (x_norm - (L_inv_x ** 2).sum()) / var
    logp = square_dist + float(dim1) * LOG2PI + logdet_CCt
    return -logp * 0.5


# ---- densities behind the non-Gaussian likelihoods (reference densities.py:28-70); elementwise
def lognormal(x, mu, var):
    lnx = torch.log(x)
    return gaussian(lnx, mu, var) - lnx


def bernoulli(p, y):
    return torch.log(torch.where(y == 1, p, 1 - p))


def poisson(lamb, y):
    return y * torch.log(lamb) - lamb - torch.lgamma(y + 1.0)


def exponential(lamb, y):
    return -y / lamb - torch.log(lamb)


def gamma(shape, scale, x):
    return -shape * torch.log(scale) - torch.lgamma(shape) + (shape - 1.0) * torch.log(x) - x / scale


def student_t(x, mean, scale, deg_free):
    df = torch.as_tensor(deg_free, dtype=x.dtype, device=x.device)
    const = torch.lgamma((df + 1.0) * 0.5) - torch.lgamma(df * 0.5) \
        - 0.5 * (torch.log(scale ** 2) + torch.log(df) + np.log(np.pi))
    return const - 0.5 * (df + 1.0) * torch.log(1.0 + (1.0 / df) * ((x - mean) / scale) ** 2)


def beta(alpha, beta, y):
    y = torch.clamp(y, 1e-6, 1 - 1e-6)
    return (alpha - 1.0) * torch.log(y) + (beta - 1.0) * torch.log(1.0 - y) \
        + torch.lgamma(alpha + beta) - torch.lgamma(alpha) - torch.lgamma(beta)


def laplace(mu, sigma, y):
    return -torch.abs(mu - y) / sigma - torch.log(2.0 * sigma)
