"""KL[q || p] between Gaussians (reference kullback_leiblers.py:26-105); the Cholesky and the
triangular solves run in the CUDA library, the remaining terms are O(M K) reductions."""
import torch

from ._backend import ops as _ops
from .misc import to_tensor


def gauss_kl(q_mu, q_sqrt, K=None):
    q_mu, q_sqrt = to_tensor(q_mu), to_tensor(q_sqrt)
    white = K is None
    if white:
        alpha = q_mu
    else:
        K = to_tensor(K)
        Lp = _ops.cholesky(K)
        alpha = _ops.solve_lower(Lp, q_mu)                        # :54
    if q_sqrt.dim() == 2:
        diag = True
        num_latent = q_sqrt.shape[1]
        NM = q_sqrt.numel()
        Lq = Lq_diag = q_sqrt
    elif q_sqrt.dim() == 3:
        diag = False
        num_latent = q_sqrt.shape[2]
        NM = q_sqrt.shape[1] * q_sqrt.shape[2]
        Lq = torch.tril(q_sqrt.permute(2, 0, 1))                  # :61 force lower triangle
        Lq_diag = torch.diagonal(Lq, dim1=-2, dim2=-1)
    else:
        raise ValueError('Bad dimension for q_sqrt: {}'.format(q_sqrt.dim()))

    mahalanobis = (alpha ** 2).sum()                              # :67
    constant = -float(NM)
    logdet_qcov = torch.log(Lq_diag ** 2).sum()                   # :73

    if white:
        trace = (Lq ** 2).sum()                                   # :77
    elif diag:
        # diag(K^-1) = row sums of squares of U = Lp^-T   (:80-86)
        U = _ops._TriInvT.apply(Lp)
        kinv_diag = (U ** 2).sum(1)
        trace = (kinv_diag[:, None] * q_sqrt ** 2).sum()
    else:
        # sum_k |Lp^-1 Lq_k|^2 : (Lp^-1 Lq_k)^T = Lq_k^T Lp^-T   (:88-94)
        trace = 0.0
        for k in range(num_latent):
            X = _ops.trsm_rlt(_ops.t(Lq[k]), Lp)
            trace = trace + (X ** 2).sum()

    twoKL = mahalanobis + constant - logdet_qcov + trace
    if not white:
        twoKL = twoKL + num_latent * torch.log(torch.diagonal(Lp) ** 2).sum()   # :99-103
    return 0.5 * twoKL
