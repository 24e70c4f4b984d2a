"""Sparse GP regression, Titsias' collapsed bound (reference models/sgpr.py:85-189)."""
import numpy as np
import torch

from .. import features, likelihoods
from .._backend import ops as _ops
from .._settings import SETTINGS as settings
from ..misc import to_tensor
from .model import GPModel


class SGPR(GPModel):
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
        Kfu = self.kern.K(self.X, self.feature.Z)                                  # Kuf^T
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
        Ksu = self.kern.K(Xnew, self.feature.Z)                                    # Kus^T [N*, M]
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
