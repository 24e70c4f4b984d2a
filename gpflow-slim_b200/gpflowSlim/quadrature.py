"""Gauss-Hermite nodes for the likelihood expectations (reference quadrature.py:24-27)."""
import numpy as np


def hermgauss(n):
    x, w = np.polynomial.hermite.hermgauss(n)
    return x.astype(np.float64), w.astype(np.float64)


def mvhermgauss(H, D):
    """Tensor-product Gauss-Hermite rule: locations [H**D, D] and weights [H**D] for
    int exp(-x.x) f(x) dx (quadrature.py:30-45)."""
    import itertools
    gh_x, gh_w = hermgauss(H)
    x = np.array(list(itertools.product(*(gh_x,) * D)))
    w = np.prod(np.array(list(itertools.product(*(gh_w,) * D))), 1)
    return x, w


def mvnquad(func, means, covs, H, Din, Dout=()):
    """N Gaussian expectations E_{N(mean_n, cov_n)}[func] by Gauss-Hermite quadrature
    (quadrature.py:48-77).  means [N, Din], covs [N, Din, Din]; func maps [?, Din] ->
    [?, *Dout]; returns [N, *Dout].  The N small Cholesky factors (Din x Din, Din <= ~3 in
    practice) are a batched elementwise job, not hot-path work."""
    import torch
    xn, wn = mvhermgauss(H, Din)
    N = means.shape[0]
    xn_t = torch.as_tensor(xn, dtype=means.dtype, device=means.device)
    cholXcov = torch.linalg.cholesky(covs)                                  # [N, D, D]
    Xt = cholXcov @ xn_t.t().unsqueeze(0).expand(N, -1, -1)                 # [N, D, H**D]
    X = 2.0 ** 0.5 * Xt + means.unsqueeze(2)
    Xr = X.permute(2, 0, 1).reshape(-1, Din)                                # [(H**D * N), D]
    fX = func(Xr).reshape((H ** Din, N) + tuple(Dout))
    wr = np.reshape(wn * np.pi ** (-Din * 0.5), (-1,) + (1,) * (1 + len(Dout)))
    return (fX * torch.as_tensor(wr, dtype=means.dtype, device=means.device)).sum(0)
