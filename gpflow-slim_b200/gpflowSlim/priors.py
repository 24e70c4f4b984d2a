"""Parameter priors (reference priors.py:27-124): log densities summed over the parameter's
entries, added to the objective through `Model.prior_tensor` (models/model.py:57-65) together
with the transform's log-Jacobian (params.py:176-194).  Elementwise device work."""
import numpy as np
import torch

from . import densities


def _on(x, a):
    return torch.as_tensor(a, dtype=x.dtype, device=x.device)


class Prior(object):
    def logp(self, x):
        raise NotImplementedError

    def sample(self, shape=(1,)):
        raise NotImplementedError


class Gaussian(Prior):
    def __init__(self, mu, var):
        self.mu = np.atleast_1d(np.array(mu, np.float64))
        self.var = np.atleast_1d(np.array(var, np.float64))

    def logp(self, x):
        return densities.gaussian(x, _on(x, self.mu), _on(x, self.var)).sum()

    def sample(self, shape=(1,)):
        return self.mu + np.sqrt(self.var) * np.random.randn(*shape)

    def __str__(self):
        return 'N(' + str(self.mu) + ',' + str(self.var) + ')'


class LogNormal(Prior):
    def __init__(self, mu, var):
        self.mu = np.atleast_1d(np.array(mu, np.float64))
        self.var = np.atleast_1d(np.array(var, np.float64))

    def logp(self, x):
        return densities.lognormal(x, _on(x, self.mu), _on(x, self.var)).sum()

    def sample(self, shape=(1,)):
        return np.exp(self.mu + np.sqrt(self.var) * np.random.randn(*shape))

    def __str__(self):
        return 'logN(' + str(self.mu) + ',' + str(self.var) + ')'


class Gamma(Prior):
    def __init__(self, shape, scale):
        self.shape = np.atleast_1d(np.array(shape, np.float64))
        self.scale = np.atleast_1d(np.array(scale, np.float64))

    def logp(self, x):
        return densities.gamma(_on(x, self.shape), _on(x, self.scale), x).sum()

    def sample(self, shape=(1,)):
        return np.random.gamma(self.shape, self.scale, size=shape)

    def __str__(self):
        return 'Ga(' + str(self.shape) + ',' + str(self.scale) + ')'


class Laplace(Prior):
    def __init__(self, mu, sigma):
        self.mu = np.atleast_1d(np.array(mu, np.float64))
        self.sigma = np.atleast_1d(np.array(sigma, np.float64))

    def logp(self, x):
        return densities.laplace(_on(x, self.mu), _on(x, self.sigma), x).sum()

    def sample(self, shape=(1,)):
        return np.random.laplace(self.mu, self.sigma, size=shape)

    def __str__(self):
        return 'Lap.(' + str(self.mu) + ',' + str(self.sigma) + ')'


class Beta(Prior):
    def __init__(self, a, b):
        self.a = np.atleast_1d(np.array(a, np.float64))
        self.b = np.atleast_1d(np.array(b, np.float64))

    def logp(self, x):
        return densities.beta(_on(x, self.a), _on(x, self.b), x).sum()

    def sample(self, shape=(1,)):
        return np.random.beta(self.a, self.b, size=shape)

    def __str__(self):
        return 'Beta(' + str(self.a) + ',' + str(self.b) + ')'


class Uniform(Prior):
    def __init__(self, lower=0., upper=1.):
        self.log_height = - np.log(upper - lower)
        self.lower, self.upper = lower, upper

    def logp(self, x):
        return _on(x, self.log_height * float(x.numel()))

    def sample(self, shape=(1,)):
        return self.lower + (self.upper - self.lower) * np.random.rand(*shape)

    def __str__(self):
        return 'U(' + str(self.lower) + ',' + str(self.upper) + ')'
