"""Gaussian-process regression (reference models/gpr.py:25-132).

Fast path: the whole objective -- Gram, + noise I, Cholesky, alpha = L^-1 (Y - m), log-det,
quadratic form AND its gradient w.r.t. kernel parameters, noise and (Y - m) -- is ONE library
call (`gps_gpr_nlml_fwd_bwd`, csrc/gpr.cu); prediction is `gps_gpr_predict`.  The op-by-op
path (`fused=False`) goes through the same CUDA kernels one autograd op at a time and is
what the tests compare the fused path against."""
import torch

from .. import likelihoods
from .._backend import ops as _ops
from ..densities import multivariate_normal
from ..misc import to_tensor
from .model import GPModel


class GPR(GPModel):
    def __init__(self, X, Y, kern, mean_function=None, obs_var=0.1, num_latent=None, min_var=None,
                 fused=True, **kwargs):
        likelihood = likelihoods.Gaussian(var=obs_var, min_var=min_var)
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function, **kwargs)
        self.num_latent = self.Y.shape[1] if num_latent is None else num_latent
        self.fused = fused

    def _fusable(self, *tensors):
        if not self.fused or self.Y.shape[1] > 16:
            return False
        if any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
            return False
        try:
            self.kern.program()
        except NotImplementedError:
            return False
        return True

    def _build_likelihood(self):
        """log p(Y | theta) (models/gpr.py:55-72).  NOTE: no jitter, only + noise I (:69)."""
        m = self.mean_function(self.X)
        if self._fusable(self.X):
            return _ops.gpr_loglik(self.kern.program(), self.X, self.Y - m, self.likelihood.variance)
        n = self.X.shape[0]
        K = self.kern.K(self.X) + torch.eye(n, dtype=self.X.dtype, device=self.X.device) \
            * self.likelihood.variance
        L = _ops.cholesky(K)
        return multivariate_normal(self.Y, m, L)

    def _build_predict(self, Xnew, full_cov=False):
        """p(F* | Y) (models/gpr.py:118-131)."""
        Xnew = to_tensor(Xnew)
        r = self.Y.shape[1]
        if self._fusable(self.X, Xnew) and not torch.is_grad_enabled():
            mean, var = _ops.gpr_predict(self.kern.program(), self.X, self.Y - self.mean_function(self.X),
                                         self.likelihood.variance, Xnew, full_cov=full_cov)
            fmean = mean + self.mean_function(Xnew)
            if full_cov:
                return fmean, var.unsqueeze(2).expand(-1, -1, r)
            return fmean, var.reshape(-1, 1).expand(-1, r)
        n = self.X.shape[0]
        Kxt = self.kern.K(Xnew, self.X)                       # K(X, Xnew)^T
        K = self.kern.K(self.X) + torch.eye(n, dtype=self.X.dtype, device=self.X.device) \
            * self.likelihood.variance
        L = _ops.cholesky(K)
        At = _ops.trsm_rlt(Kxt, L)                            # (L^-1 Kx)^T
        Vt = _ops.trsm_rlt(_ops.t(self.Y - self.mean_function(self.X)), L)
        fmean = _ops.matmul_nt(At, Vt) + self.mean_function(Xnew)
        if full_cov:
            fvar = self.kern.K(Xnew) - _ops.matmul_nt(At, At)
            fvar = fvar.unsqueeze(2).expand(-1, -1, r)
        else:
            fvar = self.kern.Kdiag(Xnew) - (At ** 2).sum(1)
            fvar = fvar.reshape(-1, 1).expand(-1, r)
        return fmean, fvar
