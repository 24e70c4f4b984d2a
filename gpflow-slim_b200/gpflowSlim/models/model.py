"""Model base classes (reference models/model.py:28-196)."""
import torch

from .._settings import SETTINGS as settings
from ..mean_functions import Zero
from ..misc import to_tensor


class Model(object):
    def __init__(self, name='model'):
        self._name = name
        self._parameters = []

    @property
    def name(self):
        return self._name

    @property
    def parameters(self):
        return self._parameters

    @property
    def trainable_tensors(self):
        """The unconstrained leaves an optimiser steps (the reference trains
        tf.trainable_variables(), which also holds feature.Z: examples/svgp.py:160-163)."""
        out = []
        for p in self.parameters:       # a parameter may be listed twice (Polynomial, kernels.py:541)
            if p.trainable and not any(p.unconstrained_tensor is q for q in out):
                out.append(p.unconstrained_tensor)
        feat = getattr(self, 'feature', None)
        for p in (getattr(feat, 'parameters', None) or []):      # Z; Multiscale: Z and scales
            if p.trainable and not any(p.unconstrained_tensor is q for q in out):
                out.append(p.unconstrained_tensor)
        return out

    def compute_log_prior(self):
        return self.prior_tensor

    def compute_log_likelihood(self):
        return self.likelihood_tensor

    @property
    def likelihood_tensor(self):
        return self._build_likelihood()

    @property
    def prior_tensor(self):
        """models/model.py:57-65."""
        priors = [p._build_prior(p.unconstrained_tensor, p.constrained_tensor)
                  for p in self.parameters if p.prior is not None]
        if not priors:
            return torch.zeros((), dtype=torch.float64, device=settings.device)
        return sum(priors)

    @property
    def objective(self):
        """-(log likelihood + log prior), recomputed on every access exactly like the
        reference's eager mode (models/model.py:67-73)."""
        return -(self.likelihood_tensor + self.prior_tensor)

    def _build_likelihood(self):
        raise NotImplementedError


class GPModel(Model):
    """models/model.py:76-170."""

    def __init__(self, X, Y, kern, likelihood, mean_function, name='GPModel'):
        super().__init__(name=name)
        self.mean_function = mean_function or Zero()
        self.kern = kern
        self.likelihood = likelihood
        self.X, self.Y = to_tensor(X), to_tensor(Y)
        self._parameters = self.mean_function.parameters + self.kern.parameters + self.likelihood.parameters

    def predict_f(self, Xnew):
        return self._build_predict(to_tensor(Xnew))

    def predict_f_full_cov(self, Xnew):
        return self._build_predict(to_tensor(Xnew), full_cov=True)

    def predict_f_samples(self, Xnew, num_samples):
        """models/model.py:135-148 (posterior samples via a jittered Cholesky of the full
        covariance)."""
        from .._backend import ops as _ops
        mu, var = self._build_predict(to_tensor(Xnew), full_cov=True)
        jitter = torch.eye(mu.shape[0], dtype=mu.dtype, device=mu.device) * settings.numerics.jitter_level
        samples = []
        for i in range(self.num_latent):
            L = _ops.cholesky(var[:, :, i].contiguous() + jitter)
            V = torch.randn(L.shape[0], num_samples, dtype=mu.dtype, device=mu.device)
            samples.append(mu[:, i:i + 1] + _ops.matmul_nt(L, _ops.t(V), a_tri=1))
        return torch.stack(samples).permute(2, 1, 0)

    def predict_y(self, Xnew):
        mu, var = self._build_predict(to_tensor(Xnew))
        return self.likelihood.predict_mean_and_var(mu, var)

    def predict_density(self, Xnew, Ynew):
        mu, var = self._build_predict(to_tensor(Xnew))
        return self.likelihood.predict_density(mu, var, to_tensor(Ynew))

    def _build_predict(self, *args, **kwargs):
        raise NotImplementedError

    def optimize(self, max_iter=1000):
        """Eager L-BFGS on the objective (models/model.py:172-195): a fresh LBFGS with 20
        correction pairs per call and the optimiser's own default of 100 iterations -- like the
        reference, `max_iter` is accepted and not used.  If the run raises (a failed Cholesky,
        'Very unstable, exit'), the objective history is printed and the parameters are reset to
        the third-last recorded iterate (:180-187)."""
        from ..LBFGS import LBFGS, model_opfunc
        self.LBFGS_opt = LBFGS(model_opfunc(self), nCorrection=20)
        try:
            self.LBFGS_opt.run()
        except Exception:
            opt = self.LBFGS_opt
            print([float(h[0]) for h in opt.history])
            opt.update_vars(opt.history[0][3], opt.history[-3][2])
