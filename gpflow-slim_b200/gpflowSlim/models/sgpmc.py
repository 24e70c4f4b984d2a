"""Sparse variational GP with MCMC over the whitened inducing values (Hensman et al. 2015;
reference models/sgpmc.py:25-105): u = L v, v ~ N(0, I)."""
import numpy as np

from ..features import conditional, inducingpoint_wrapper
from ..params import Parameter
from ..priors import Gaussian
from .model import GPModel


class SGPMC(GPModel):
    def __init__(self, X, Y, kern, likelihood, feat=None, mean_function=None, num_latent=None,
                 Z=None, **kwargs):
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function, **kwargs)
        self.num_data = self.X.shape[0]
        self.num_latent = num_latent or self.Y.shape[1]
        self.feature = inducingpoint_wrapper(feat, Z)
        self._V = Parameter(np.zeros((len(self.feature), self.num_latent)), name='V')
        self._V.prior = Gaussian(0., 1.)
        self._parameters = self._parameters + [self._V]

    @property
    def V(self):
        return self._V.value

    def _build_likelihood(self):
        """Optimal q*(v) up to a constant: the variational expectations under the exact
        conditional marginals (sgpmc.py:82-88)."""
        fmean, fvar = self._build_predict(self.X, full_cov=False)
        return self.likelihood.variational_expectations(fmean, fvar, self.Y).sum()

    def _build_predict(self, Xnew, full_cov=False):
        """p(F* | U = L V) (sgpmc.py:90-104)."""
        mu, var = conditional(self.feature, self.kern, Xnew, self.V, full_cov=full_cov, q_sqrt=None,
                              white=True)
        return mu + self.mean_function(Xnew), var
