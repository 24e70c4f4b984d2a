"""Gaussian-process regression (reference models/gpr.py:25-132).

Fast path: the whole objective -- Gram, + noise I, Cholesky, alpha = L^-1 (Y - m), log-det,
quadratic form AND its gradient w.r.t. kernel parameters, noise and (Y - m) -- is ONE library
call (`gps_gpr_nlml_fwd_bwd`, csrc/gpr.cu); prediction is `gps_gpr_predict`.  The op-by-op
path (`fused=False`) goes through the same CUDA kernels one autograd op at a time and is
what the tests compare the fused path against."""
import torch

from .. import likelihoods
from .._backend import ops as _ops
from ..densities import multivariate_normal, multivariate_normal_feature
from ..misc import to_tensor
from .model import GPModel


class GPR(GPModel):
    def __init__(self, X, Y, kern, mean_function=None, obs_var=0.1, num_latent=None, min_var=None,
                 fused=True, **kwargs):
        likelihood = likelihoods.Gaussian(var=obs_var, min_var=min_var)
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function, **kwargs)
        self.num_latent = self.Y.shape[1] if num_latent is None else num_latent
        self.fused = fused
        self._split = None

    def _core_and_white(self):
        """(covariance without its top-level White terms, summed White variance or None):
        `RBF + White` is the RBF with the white variance added to the noise on the diagonal."""
        core, whites = self.kern, []
        if hasattr(self.kern, 'split_white'):
            if self._split is None:
                self._split = self.kern.split_white()
            core, whites = self._split
        extra = None
        for w in whites:
            extra = w.variance.reshape(()) if extra is None else extra + w.variance.reshape(())
        return core, extra

    def _fusable(self, *tensors):
        if not self.fused or self.Y.shape[1] > 16:
            return False
        if any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
            return False
        core, _ = self._core_and_white()
        return bool(core is not None and getattr(core, 'fusable', False))     # cached by the kernel object

    def _feature_map(self):
        """`kern.features` if the covariance is an explicit feature expansion K = C C^T
        (kernel_kitchen_sink.SamplerKernel), else None (models/gpr.py:62-63, :85-86)."""
        fn = getattr(self.kern, 'features', None)
        return fn if fn is not None and callable(fn) else None

    def _build_likelihood(self):
        """log p(Y | theta) (models/gpr.py:55-72).  NOTE: no jitter, only + noise I (:69)."""
        m = self.mean_function(self.X)
        features = self._feature_map()
        if features is not None:                     # Woodbury form, no N x N matrix (:62-66)
            var = self.likelihood.variance
            var = var if isinstance(var, torch.Tensor) else torch.as_tensor(var, dtype=self.X.dtype,
                                                                            device=self.X.device)
            return multivariate_normal_feature(self.Y, m, features(self.X), var)
        if self._fusable(self.X):
            core, white = self._core_and_white()
            noise = self.likelihood.variance if white is None else self.likelihood.variance.reshape(()) + white
            return _ops.gpr_loglik(core.program(), self.X, self.Y - m, noise)
        n = self.X.shape[0]
        K = self.kern.K(self.X) + torch.eye(n, dtype=self.X.dtype, device=self.X.device) \
            * self.likelihood.variance
        L = _ops.cholesky(K)
        return multivariate_normal(self.Y, m, L)

    def _build_predict(self, Xnew, full_cov=False):
        """p(F* | Y) (models/gpr.py:118-131)."""
        Xnew = to_tensor(Xnew)
        r = self.Y.shape[1]
        features = self._feature_map()
        if features is not None:
            return self._build_predict_features(features, Xnew, full_cov)
        from .. import parallel
        if parallel.active() and not full_cov and self._fusable(self.X, Xnew) and not torch.is_grad_enabled():
            # all ranks of the group factor K + noise I together and share out the test points
            from .._backend import dist_gpr
            core, white = self._core_and_white()
            noise = self.likelihood.variance if white is None else self.likelihood.variance.reshape(()) + white
            prog = core.program()
            mean, var = dist_gpr.predict(prog, prog.theta(self.X.device).detach(), float(noise), self.X,
                                         (self.Y - self.mean_function(self.X)).detach(), Xnew,
                                         core.Kdiag(Xnew).detach(), block=parallel.block(), group=parallel.group(),
                                         lookahead=parallel.lookahead())
            if white is not None:
                var = var + white
            return mean + self.mean_function(Xnew), var.reshape(-1, 1).expand(-1, r)
        if self._fusable(self.X, Xnew) and not torch.is_grad_enabled():
            core, white = self._core_and_white()
            noise = self.likelihood.variance if white is None else self.likelihood.variance.reshape(()) + white
            mean, var = _ops.gpr_predict(core.program(), self.X, self.Y - self.mean_function(self.X),
                                         noise, Xnew, full_cov=full_cov)
            if white is not None:        # K(X*, X*) / Kdiag(X*) of the full covariance include the white term
                var = var + (torch.eye(var.shape[0], dtype=var.dtype, device=var.device) * white if full_cov
                             else white)
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

    def _build_predict_features(self, features, Xnew, full_cov):
        """Prediction through the feature expansion (models/gpr.py:84-114): with C = features(X),
        B = features(Xnew) and S = C^T C + sigma^2 I (small), G = (C^T - C^T C S^-1 C^T) / sigma^2
        = C^T (C C^T + sigma^2 I)^-1, fmean = B G (Y - m) + m*, fvar = B B^T - B G C B^T.
        S^-1 comes from the Cholesky factor (U U^T, U = L^-T) instead of the reference's
        tf.matrix_inverse; all products run on the tensor-core GEMM."""
        from .._backend.lib import TRI_UPPER
        r = self.Y.shape[1]
        var = self.likelihood.variance
        mX, m_new = self.mean_function(self.X), self.mean_function(Xnew)
        feat, feat_new = features(self.X), features(Xnew)
        Ct = _ops.t(feat)                                         # [d, N]
        CtC = _ops.matmul_nt(Ct, Ct)                              # [d, d]
        S = CtC + torch.eye(CtC.shape[0], dtype=CtC.dtype, device=CtC.device) * var
        U = _ops._TriInvT.apply(_ops.cholesky(S))
        S_inv = _ops.matmul_nt(U, U, a_tri=TRI_UPPER, b_tri=TRI_UPPER)
        tmp = _ops.matmul(CtC, _ops.matmul(S_inv, Ct))
        G = (Ct - tmp) / var                                      # [d, N]
        fmean = _ops.matmul(feat_new, _ops.matmul(G, self.Y - mX)) + m_new
        GC = _ops.matmul(G, feat)                                 # [d, d]
        if full_cov:
            BBt = _ops.matmul_nt(feat_new, feat_new)
            fvar = BBt - _ops.matmul(feat_new, _ops.matmul_nt(GC, feat_new))
            return fmean, fvar.unsqueeze(2).expand(-1, -1, r)
        tmp = _ops.matmul(feat_new, GC)
        fvar = (feat_new ** 2.).sum(-1) - (tmp * feat_new).sum(-1)
        return fmean, fvar.reshape(-1, 1).expand(-1, r)

