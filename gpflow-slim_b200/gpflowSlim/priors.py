"""Parameter priors (the reference's priors.py:27-124 surface: Gaussian, LogNormal, Gamma, Laplace,
Beta, Uniform with `.logp(x)`, `.sample(shape)`, `str()`).

Design here: ONE generic class driven by a small table -- per family the hyper-parameter names,
the elementwise log density (from `densities`), a numpy sampler and a label.  `logp` moves the
hyper-parameters to the device of its argument and returns the SUM of the elementwise log
densities, which `Model.prior_tensor` adds to the objective together with the transform's
log-Jacobian (models/model.py:57-65, params.py:176-194)."""
import numpy as np
import torch

from . import densities


class Prior(object):
    """Base of every family below; `_names` are the hyper-parameters in constructor order."""
    _names = ()
    _label = 'Prior'

    def __init__(self, *values, **named):
        given = dict(zip(self._names, values))
        given.update(named)
        missing = [n for n in self._names if n not in given]
        if missing:
            raise TypeError('%s needs %s' % (type(self).__name__, ', '.join(missing)))
        for n in self._names:
            setattr(self, n, np.atleast_1d(np.array(given[n], np.float64)))

    def _on_device_of(self, x):
        return [torch.as_tensor(getattr(self, n), dtype=x.dtype, device=x.device) for n in self._names]

    def _elementwise(self, x, *hyper):
        raise NotImplementedError

    def _draw(self, shape):
        raise NotImplementedError

    def logp(self, x):
        return self._elementwise(x, *self._on_device_of(x)).sum()

    def sample(self, shape=(1,)):
        return self._draw(tuple(shape))

    def __str__(self):
        return '%s(%s)' % (self._label, ','.join(str(getattr(self, n)) for n in self._names))


def _family(name, names, label, elementwise, draw):
    return type(name, (Prior,), dict(_names=names, _label=label, __doc__='%s prior over %s.' % (name, names),
                                     _elementwise=staticmethod(elementwise),
                                     _draw=lambda self, shape: draw(self, shape)))


Gaussian = _family('Gaussian', ('mu', 'var'), 'N',
                   lambda x, mu, var: densities.gaussian(x, mu, var),
                   lambda p, shape: p.mu + np.sqrt(p.var) * np.random.randn(*shape))
LogNormal = _family('LogNormal', ('mu', 'var'), 'logN',
                    lambda x, mu, var: densities.lognormal(x, mu, var),
                    lambda p, shape: np.exp(p.mu + np.sqrt(p.var) * np.random.randn(*shape)))
Gamma = _family('Gamma', ('shape', 'scale'), 'Ga',
                lambda x, shape, scale: densities.gamma(shape, scale, x),
                lambda p, shape: np.random.gamma(p.shape, p.scale, size=shape))
Laplace = _family('Laplace', ('mu', 'sigma'), 'Lap.',
                  lambda x, mu, sigma: densities.laplace(mu, sigma, x),
                  lambda p, shape: np.random.laplace(p.mu, p.sigma, size=shape))
Beta = _family('Beta', ('a', 'b'), 'Beta',
               lambda x, a, b: densities.beta(a, b, x),
               lambda p, shape: np.random.beta(p.a, p.b, size=shape))


class Uniform(Prior):
    """Flat density on [lower, upper]: logp = -log(upper - lower) per entry, whatever the value
    (the reference does not test the support either, priors.py:117-118)."""
    _label = 'U'

    def __init__(self, lower=0., upper=1.):
        self.lower, self.upper = lower, upper
        self.log_height = -np.log(upper - lower)

    def logp(self, x):
        return torch.as_tensor(self.log_height * float(x.numel()), dtype=x.dtype, device=x.device)

    def sample(self, shape=(1,)):
        return self.lower + (self.upper - self.lower) * np.random.rand(*shape)

    def __str__(self):
        return 'U(%s,%s)' % (self.lower, self.upper)
