"""Sparse GP regression: Titsias' collapsed bound (reference models/sgpr.py:85-189), the FITC
approximation (:192-317) and the upper bound shared by both (SGPRUpperMixin, :30-82).  All in the
transposed orientation (row-major NT products, see conditionals.py)."""
import numpy as np
import torch

from .. import features, likelihoods
from .._backend import ops as _ops
from .._settings import SETTINGS as settings
from ..misc import to_tensor
from .model import GPModel


class SGPRUpperMixin(object):
    """Upper bound on the GP regression marginal likelihood with the model's own inducing points
    (models/sgpr.py:30-82; Titsias 2014, trace bound)."""

    def compute_upper_bound(self):
        num_data = float(self.Y.shape[0])
        Kdiag = self.kern.Kdiag(self.X)
        Kuu = self.feature.Kuu(self.kern, jitter=settings.numerics.jitter_level)
        Kfu = self.feature.Kfu(self.kern, self.X)                                  # Kuf^T  [N, M]
        Kuf = _ops.t(Kfu)
        var = self.likelihood.variance
        L = _ops.cholesky(Kuu)                                                     # :62
        KufKfu = _ops.matmul_nt(Kuf, Kuf)                                          # Kuf Kuf^T (M^2 N)
        LB = _ops.cholesky(Kuu + KufKfu / var)                                     # :63
        LinvKuf_t = _ops.trsm_rlt(Kfu, L)                                          # :65, transposed
        c = Kdiag.sum() - (LinvKuf_t ** 2).sum()                                   # :67
        corrected_noise = var + c                                                  # :73
        const = -0.5 * num_data * torch.log(2 * np.pi * var)                       # :75
        logdet = torch.log(torch.diagonal(L)).sum() - torch.log(torch.diagonal(LB)).sum()
        LC = _ops.cholesky(Kuu + KufKfu / corrected_noise)                         # :78
        KufY = _ops.matmul_nt(Kuf, _ops.t(self.Y))                                 # [M, R]
        v = _ops.solve_lower(LC, KufY / corrected_noise)                           # :79
        quad = -0.5 / corrected_noise * (self.Y ** 2).sum() + 0.5 * (v ** 2).sum()
        return const + logdet + quad


class SGPR(GPModel, SGPRUpperMixin):
    def __init__(self, X, Y, kern, feat=None, mean_function=None, Z=None, obs_var=0.1, num_data=None,
                 num_latent=None, **kwargs):
        likelihood = likelihoods.Gaussian(var=obs_var)
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function, **kwargs)
        self.feature = features.inducingpoint_wrapper(feat, Z)
        self.num_data = self.X.shape[0] if num_data is None else num_data
        self.num_latent = self.Y.shape[1] if num_latent is None else num_latent

    def _common(self):
        """Shared by the bound and the predictor (sgpr.py:132-145 / :165-174), in the transposed
        orientation: At = Kuf^T L^-T / sigma  [N, M]."""
        M = len(self.feature)
        err = self.Y - self.mean_function(self.X)
        Kfu = self.feature.Kfu(self.kern, self.X)                                  # Kuf^T
        Kuu = self.feature.Kuu(self.kern, jitter=settings.numerics.jitter_level)
        L = _ops.cholesky(Kuu)
        var = self.likelihood.variance
        sigma = torch.sqrt(var)
        At = _ops.trsm_rlt(Kfu, L) / sigma                                         # (L^-1 Kuf / sigma)^T
        Att = _ops.t(At)                                                           # A  [M, N]
        AAT = _ops.matmul_nt(Att, Att)                                             # A A^T  (SYRK, M^2 N)
        B = AAT + torch.eye(M, dtype=AAT.dtype, device=AAT.device)
        LB = _ops.cholesky(B)
        Aerr = _ops.matmul_nt(Att, _ops.t(err))                                    # A err  [M, R]
        c = _ops.solve_lower(LB, Aerr) / sigma
        return err, L, LB, AAT, c, var

    def _build_likelihood(self):
        """sgpr.py:121-156."""
        err, L, LB, AAT, c, var = self._common()
        num_data, output_dim = float(self.Y.shape[0]), float(self.Y.shape[1])
        Kdiag = self.kern.Kdiag(self.X)
        bound = -0.5 * num_data * output_dim * np.log(2 * np.pi)
        bound = bound - output_dim * torch.log(torch.diagonal(LB)).sum()
        bound = bound - 0.5 * num_data * output_dim * torch.log(var)
        bound = bound - 0.5 * (err ** 2).sum() / var
        bound = bound + 0.5 * (c ** 2).sum()
        bound = bound - 0.5 * output_dim * Kdiag.sum() / var
        bound = bound + 0.5 * output_dim * torch.diagonal(AAT).sum()
        return bound

    def _build_predict(self, Xnew, full_cov=False):
        """sgpr.py:158-189."""
        Xnew = to_tensor(Xnew)
        err, L, LB, AAT, c, var = self._common()
        Ksu = self.feature.Kfu(self.kern, Xnew)                                    # Kus^T [N*, M]
        tmp1t = _ops.trsm_rlt(Ksu, L)                                              # (L^-1 Kus)^T
        tmp2t = _ops.trsm_rlt(tmp1t, LB)                                           # (LB^-1 tmp1)^T
        mean = _ops.matmul_nt(tmp2t, _ops.t(c))                                    # tmp2^T c
        r = self.Y.shape[1]
        if full_cov:
            v = self.kern.K(Xnew) + _ops.matmul_nt(tmp2t, tmp2t) - _ops.matmul_nt(tmp1t, tmp1t)
            v = v.unsqueeze(2).expand(-1, -1, r)
        else:
            v = self.kern.Kdiag(Xnew) + (tmp2t ** 2).sum(1) - (tmp1t ** 2).sum(1)
            v = v.unsqueeze(1).expand(-1, r)
        return mean + self.mean_function(Xnew), v


class GPRFITC(GPModel, SGPRUpperMixin):
    """GP regression with the FITC approximation (models/sgpr.py:192-326).  (The reference's
    constructor reads `self.name` before it is set, :221, and cannot run as shipped; the
    arguments and everything after construction follow it.)"""

    def __init__(self, X, Y, kern, feat=None, mean_function=None, Z=None, obs_var=0.1, num_data=None,
                 num_latent=None, **kwargs):
        likelihood = likelihoods.Gaussian(var=obs_var)
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function, **kwargs)
        self.feature = features.inducingpoint_wrapper(feat, Z)
        self.num_data = self.X.shape[0] if num_data is None else num_data
        self.num_latent = self.Y.shape[1] if num_latent is None else num_latent

    def _build_common_terms(self):
        """:227-247 with V kept as V^T [N, M]."""
        M = len(self.feature)
        err = self.Y - self.mean_function(self.X)
        Kdiag = self.kern.Kdiag(self.X)
        Kfu = self.feature.Kfu(self.kern, self.X)                                  # Kuf^T
        Kuu = self.feature.Kuu(self.kern, jitter=settings.numerics.jitter_level)
        Luu = _ops.cholesky(Kuu)
        Vt = _ops.trsm_rlt(Kfu, Luu)                                               # (Luu^-1 Kuf)^T
        diagQff = (Vt ** 2).sum(1)
        nu = Kdiag - diagQff + self.likelihood.variance
        V = _ops.t(Vt)
        B = torch.eye(M, dtype=V.dtype, device=V.device) + _ops.matmul_nt(_ops.t(Vt / nu[:, None]), V)
        L = _ops.cholesky(B)
        beta = err / nu[:, None]
        alpha = _ops.matmul_nt(V, _ops.t(beta))                                    # V beta  [M, R]
        gamma = _ops.solve_lower(L, alpha)
        return err, nu, Luu, L, alpha, beta, gamma

    def _build_likelihood(self):
        """:249-291."""
        err, nu, Luu, L, alpha, beta, gamma = self._build_common_terms()
        mahalanobis = -0.5 * (err ** 2 / nu[:, None]).sum() + 0.5 * (gamma ** 2).sum()
        constant = -0.5 * float(self.num_data) * np.log(2.0 * np.pi)
        logdet = -0.5 * torch.log(nu).sum() - torch.log(torch.diagonal(L)).sum()
        return mahalanobis + (constant + logdet) * float(self.num_latent)

    def _build_predict(self, Xnew, full_cov=False):
        """:293-317."""
        Xnew = to_tensor(Xnew)
        _, _, Luu, L, _, _, gamma = self._build_common_terms()
        Ksu = self.feature.Kfu(self.kern, Xnew)                                    # Kus^T [N*, M]
        wt = _ops.trsm_rlt(Ksu, Luu)                                               # (Luu^-1 Kus)^T
        tmp = _ops.solve_upper_t(L, gamma)                                         # L^-T gamma [M, R]
        mean = _ops.matmul_nt(wt, _ops.t(tmp)) + self.mean_function(Xnew)
        it = _ops.trsm_rlt(wt, L)                                                  # (L^-1 w)^T
        r = self.num_latent
        if full_cov:
            var = self.kern.K(Xnew) - _ops.matmul_nt(wt, wt) + _ops.matmul_nt(it, it)
            var = var.unsqueeze(2).expand(-1, -1, r)
        else:
            var = self.kern.Kdiag(Xnew) - (wt ** 2).sum(1) + (it ** 2).sum(1)
            var = var.unsqueeze(1).expand(-1, r)
        return mean, var
