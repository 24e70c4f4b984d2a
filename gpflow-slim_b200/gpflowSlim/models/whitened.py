"""The two whitened-latent models of the reference (models/gpmc.py:28-95, models/sgpmc.py:25-104):
latent values are represented as v ~ N(0, I) with  f = L v + m(x),  L L^T = K + jitter I
(GPMC: L over the data, one v per data point) or  u = L v  at the inducing inputs (SGPMC).
Both put a standard-normal prior on V, which enters the objective through `prior_tensor`; both
predict through the whitened conditional.  The Cholesky, the triangular product L V and the
conditional run on the library kernels."""
import numpy as np

from .. import conditionals, features
from .._backend import ops as _ops
from .._backend.lib import TRI_LOWER
from .._settings import SETTINGS as settings
from ..params import Parameter
from ..priors import Gaussian
from .model import GPModel


class _WhitenedLatentModel(GPModel):
    def _init_latents(self, rows, num_latent):
        self.num_data = self.X.shape[0]
        self.num_latent = num_latent or self.Y.shape[1]
        self._V = Parameter(np.zeros((rows, self.num_latent)), name='V')
        self._V.prior = Gaussian(0., 1.)
        self._parameters = self._parameters + [self._V]

    @property
    def V(self):
        return self._V.value

    def _conditional(self, Xnew, full_cov):
        raise NotImplementedError

    def _build_predict(self, Xnew, full_cov=False):
        mu, var = self._conditional(Xnew, full_cov)
        return mu + self.mean_function(Xnew), var


class GPMC(_WhitenedLatentModel):
    """log p(Y | F = L V + m) summed over the data; p(F* | F = L V) for predictions."""

    def __init__(self, X, Y, kern, likelihood, mean_function=None, num_latent=None, **kwargs):
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function, **kwargs)
        self._init_latents(self.X.shape[0], num_latent)

    def _build_likelihood(self):
        L = _ops.cholesky(self.kern.K_jittered(self.X, settings.numerics.jitter_level))
        F = _ops.matmul_nt(L, _ops.t(self.V), a_tri=TRI_LOWER) + self.mean_function(self.X)
        return self.likelihood.logp(F, self.Y).sum()

    def _conditional(self, Xnew, full_cov):
        return conditionals.conditional(Xnew, self.X, self.kern, self.V, full_cov=full_cov,
                                        q_sqrt=None, white=True)


class SGPMC(_WhitenedLatentModel):
    """Hensman et al. 2015: the optimal q*(v) up to a constant -- the variational expectations
    under the exact conditional marginals given U = L V at the inducing inputs."""

    def __init__(self, X, Y, kern, likelihood, feat=None, mean_function=None, num_latent=None,
                 Z=None, **kwargs):
        GPModel.__init__(self, X, Y, kern, likelihood, mean_function, **kwargs)
        self.feature = features.inducingpoint_wrapper(feat, Z)
        self._init_latents(len(self.feature), num_latent)

    def _build_likelihood(self):
        fmean, fvar = self._build_predict(self.X, full_cov=False)
        return self.likelihood.variational_expectations(fmean, fvar, self.Y).sum()

    def _conditional(self, Xnew, full_cov):
        return features.conditional(self.feature, self.kern, Xnew, self.V, full_cov=full_cov,
                                    q_sqrt=None, white=True)
