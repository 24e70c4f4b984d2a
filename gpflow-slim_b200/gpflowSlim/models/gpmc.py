"""GP with a general likelihood, latent values whitened for MCMC / MAP (reference
models/gpmc.py:28-110): v ~ N(0, I), f = L v + m(x), L L^T = K + jitter I.  The Cholesky, the
triangular product L V and the conditional run on the library kernels."""
import numpy as np
import torch

from .._backend import ops as _ops
from .._backend.lib import TRI_LOWER
from .._settings import SETTINGS as settings
from ..conditionals import conditional
from ..params import Parameter
from ..priors import Gaussian
from .model import GPModel


class GPMC(GPModel):
    def __init__(self, X, Y, kern, likelihood, mean_function=None, num_latent=None, **kwargs):
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function, **kwargs)
        self.num_data = self.X.shape[0]
        self.num_latent = num_latent or self.Y.shape[1]
        self._V = Parameter(np.zeros((self.num_data, self.num_latent)), name='V')
        self._V.prior = Gaussian(0., 1.)
        self._parameters = self._parameters + [self._V]

    @property
    def V(self):
        return self._V.value

    def _build_likelihood(self):
        """log p(Y | F = L V + m) summed over the data (gpmc.py:64-78); the N(0, I) prior on V
        enters through `prior_tensor`."""
        L = _ops.cholesky(self.kern.K_jittered(self.X, settings.numerics.jitter_level))
        F = _ops.matmul_nt(L, _ops.t(self.V), a_tri=TRI_LOWER) + self.mean_function(self.X)
        return self.likelihood.logp(F, self.Y).sum()

    def _build_predict(self, Xnew, full_cov=False):
        """p(F* | F = L V) (gpmc.py:80-95)."""
        mu, var = conditional(Xnew, self.X, self.kern, self.V, full_cov=full_cov, q_sqrt=None,
                              white=True)
        return mu + self.mean_function(Xnew), var
