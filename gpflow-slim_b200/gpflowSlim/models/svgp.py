"""Sparse variational GP (reference models/svgp.py:28-130)."""
import numpy as np
import torch

from .. import conditionals, features, kullback_leiblers, transforms
from .._settings import SETTINGS as settings
from ..misc import to_tensor
from ..params import Parameter, param_value
from .model import GPModel


class SVGP(GPModel):
    def __init__(self, X, Y, kern, likelihood, feat=None, mean_function=None, num_latent=None,
                 q_diag=False, whiten=True, minibatch_size=None, Z=None, num_data=None, **kwargs):
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function, **kwargs)
        self.num_data = num_data or self.X.shape[0]
        self.q_diag, self.whiten = q_diag, whiten
        self.feature = features.inducingpoint_wrapper(feat, Z)
        self.num_latent = num_latent or self.Y.shape[1]
        num_inducing = len(self.feature)
        # svgp.py:81-89: q_mu = 0, q_sqrt = 1 (diag) or identity matrices (full)
        self._q_mu = Parameter(np.zeros((num_inducing, self.num_latent)), name='q_mu')
        if self.q_diag:
            self._q_sqrt = Parameter(np.ones((num_inducing, self.num_latent)), transforms.positive,
                                     name='q_sqrt')
        else:
            q_sqrt = np.array([np.eye(num_inducing) for _ in range(self.num_latent)]).swapaxes(0, 2)
            self._q_sqrt = Parameter(q_sqrt, transform=transforms.LowerTriangular(num_inducing, self.num_latent),
                                     name='q_sqrt')
        self._parameters = self._parameters + [self._q_mu, self._q_sqrt]

    q_mu = param_value('q_mu')
    q_sqrt = param_value('q_sqrt')

    def build_prior_KL(self):
        """svgp.py:101-106."""
        K = None if self.whiten else self.feature.Kuu(self.kern, jitter=settings.numerics.jitter_level)
        return kullback_leiblers.gauss_kl(self.q_mu, self.q_sqrt, K)

    def _build_likelihood(self):
        """ELBO = sum var_exp * N / B - KL (svgp.py:108-125)."""
        KL = self.build_prior_KL()
        fmean, fvar = self._build_predict(self.X, full_cov=False)
        var_exp = self.likelihood.variational_expectations(fmean, fvar, self.Y)
        scale = float(self.num_data) / float(self.X.shape[0])
        return var_exp.sum() * scale - KL

    def _build_predict(self, Xnew, full_cov=False):
        Xnew = to_tensor(Xnew)
        mu, var = features.conditional(self.feature, self.kern, Xnew, self.q_mu, q_sqrt=self.q_sqrt,
                                       full_cov=full_cov, white=self.whiten)
        return mu + self.mean_function(Xnew), var
