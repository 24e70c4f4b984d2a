"""Likelihoods (reference likelihoods.py).  Gaussian (:158-188) is the one on the hot path: its
closed-form variational expectation feeds the SVGP bound.  The others -- the Gauss-Hermite
defaults of the base class (:47-151), Bernoulli / probit (:262-293), Poisson, Exponential,
StudentT, Gamma, Beta (:190-376) and MultiClass with the RobustMax link (:379-489, the
likelihood of examples/svgp.py) -- are O(B * 20) elementwise torch maths around the same SVGP
kernels (`SURVEY.md` section 8(f) rank 3)."""
import numpy as np
import torch

from . import densities, transforms
from .params import Parameter, Parameterized, param_value
from .quadrature import hermgauss


def _const(a, like):
    return torch.as_tensor(a, dtype=like.dtype, device=like.device)


class Likelihood(Parameterized):
    def __init__(self, name=None):
        self.name = name or type(self).__name__
        self.num_gauss_hermite_points = 20
        self._parameters = []

    @property
    def parameters(self):
        return self._parameters

    def conditional_mean(self, F):
        raise NotImplementedError

    def conditional_variance(self, F):
        raise NotImplementedError

    def logp(self, F, Y):
        raise NotImplementedError

    def _gh_grid(self, Fmu, Fvar):
        """X[n, h] = Fmu_n + x_h sqrt(2 Fvar_n) and the weights w_h / sqrt(pi) as a column."""
        gh_x, gh_w = hermgauss(self.num_gauss_hermite_points)
        Fmu, Fvar = Fmu.reshape(-1, 1), Fvar.reshape(-1, 1)
        X = _const(gh_x, Fmu)[None, :] * torch.sqrt(2.0 * Fvar) + Fmu
        return X, _const(gh_w / np.sqrt(np.pi), Fmu).reshape(-1, 1)

    def predict_mean_and_var(self, Fmu, Fvar):
        """E[y], Var[y] under q(f) = N(Fmu, Fvar) by Gauss-Hermite quadrature (:47-87)."""
        shape = Fmu.shape
        X, w = self._gh_grid(Fmu, Fvar)
        E_y = (self.conditional_mean(X) @ w).reshape(shape)
        integrand = self.conditional_variance(X) + self.conditional_mean(X) ** 2
        V_y = (integrand @ w).reshape(shape) - E_y ** 2
        return E_y, V_y

    def predict_density(self, Fmu, Fvar, Y):
        """log int p(y = Y | f) q(f) df (:89-118)."""
        shape = Fmu.shape
        X, w = self._gh_grid(Fmu, Fvar)
        Yt = Y.reshape(-1, 1).expand(-1, self.num_gauss_hermite_points)
        return torch.log(torch.exp(self.logp(X, Yt)) @ w).reshape(shape)

    def variational_expectations(self, Fmu, Fvar, Y):
        """int log p(y | f) q(f) df (:120-151)."""
        shape = Fmu.shape
        X, w = self._gh_grid(Fmu, Fvar)
        Yt = Y.reshape(-1, 1).expand(-1, self.num_gauss_hermite_points)
        return (self.logp(X, Yt) @ w).reshape(shape)


class Gaussian(Likelihood):
    """y = f + N(0, variance): every expectation in closed form (:158-188); the likelihood of
    the GPR / SGPR / SVGP hot path."""
    variance = param_value('variance')

    def __init__(self, var=1.0, min_var=None):
        super().__init__()
        self._param('variance', var, transforms.positive if min_var is None else transforms.Log1pe(min_var))

    def logp(self, F, Y):
        return densities.gaussian(F, Y, self.variance)

    def conditional_mean(self, F):
        return F

    def conditional_variance(self, F):
        return torch.ones_like(F) * self.variance

    def predict_mean_and_var(self, Fmu, Fvar):
        return Fmu, Fvar + self.variance

    def predict_density(self, Fmu, Fvar, Y):
        return densities.gaussian(Fmu, Y, Fvar + self.variance)

    def variational_expectations(self, Fmu, Fvar, Y):
        s2 = self.variance
        return -0.5 * np.log(2 * np.pi) - 0.5 * torch.log(s2) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / s2


class _LinkLikelihood(Likelihood):
    """Likelihoods of the form p(y | rate), rate = invlink(f).  A subclass states three rules in
    terms of the rate -- `_logdensity(rate, Y)`, `_mean(rate)`, `_var(rate)` -- and, optionally,
    `_varexp_exp(Fmu, Fvar, Y)`: the closed-form variational expectation that exists when the
    link is exp (used only then; every other case falls back to Gauss-Hermite quadrature)."""
    _varexp_exp = None

    def __init__(self, invlink):
        super().__init__()
        self.invlink = invlink

    def logp(self, F, Y):
        return self._logdensity(self.invlink(F), Y)

    def conditional_mean(self, F):
        return self._mean(self.invlink(F))

    def conditional_variance(self, F):
        return self._var(self.invlink(F))

    def variational_expectations(self, Fmu, Fvar, Y):
        if self._varexp_exp is not None and self.invlink is torch.exp:
            return self._varexp_exp(Fmu, Fvar, Y)
        return Likelihood.variational_expectations(self, Fmu, Fvar, Y)


class Poisson(_LinkLikelihood):
    """Counts in bins of width `binsize`: y ~ Poisson(invlink(f) binsize)  (:190-222)."""

    def __init__(self, invlink=torch.exp, binsize=1.0):
        super().__init__(invlink)
        self.binsize = float(binsize)

    def _logdensity(self, rate, Y):
        return densities.poisson(rate * self.binsize, Y)

    def _mean(self, rate):
        return rate * self.binsize

    _var = _mean

    def _varexp_exp(self, Fmu, Fvar, Y):
        return Y * Fmu - torch.exp(Fmu + Fvar / 2) * self.binsize - torch.lgamma(Y + 1) \
            + Y * float(np.log(self.binsize))


class Exponential(_LinkLikelihood):
    """y ~ Exponential with mean invlink(f)  (:224-241)."""

    def __init__(self, invlink=torch.exp):
        super().__init__(invlink)

    def _logdensity(self, rate, Y):
        return densities.exponential(rate, Y)

    def _mean(self, rate):
        return rate

    def _var(self, rate):
        return rate ** 2

    def _varexp_exp(self, Fmu, Fvar, Y):
        return -torch.exp(-Fmu + Fvar / 2) * Y - Fmu


class Gamma(_LinkLikelihood):
    """y ~ Gamma(shape, scale = invlink(f))  (:300-332)."""
    shape = param_value('shape')

    def __init__(self, invlink=torch.exp):
        super().__init__(invlink)
        self._param('shape', 1.0, transforms.positive)

    def _logdensity(self, rate, Y):
        return densities.gamma(self.shape, rate, Y)

    def _mean(self, rate):
        return self.shape * rate

    def _var(self, rate):
        return self.shape * rate ** 2

    def _varexp_exp(self, Fmu, Fvar, Y):
        k = self.shape
        return -k * Fmu - torch.lgamma(k) + (k - 1.0) * torch.log(Y) - Y * torch.exp(-Fmu + Fvar / 2.0)


def probit(x):
    """Standard normal CDF squeezed into [1e-3, 1 - 1e-3]  (:268-269)."""
    return 0.5 * (1.0 + torch.erf(x / np.sqrt(2.0))) * (1 - 2e-3) + 1e-3


class Bernoulli(_LinkLikelihood):
    """y in {0, 1} with p(y = 1) = invlink(f); probit link: closed-form predictions (:272-297)."""

    def __init__(self, invlink=probit):
        super().__init__(invlink)

    def _logdensity(self, p, Y):
        return densities.bernoulli(p, Y)

    def _mean(self, p):
        return p

    def _var(self, p):
        return p - p ** 2

    def predict_mean_and_var(self, Fmu, Fvar):
        if self.invlink is not probit:
            return Likelihood.predict_mean_and_var(self, Fmu, Fvar)
        p = probit(Fmu / torch.sqrt(1 + Fvar))
        return p, p - p ** 2

    def predict_density(self, Fmu, Fvar, Y):
        return densities.bernoulli(self.predict_mean_and_var(Fmu, Fvar)[0], Y)


class Beta(_LinkLikelihood):
    """y in (0, 1) ~ Beta(scale m, scale (1 - m)) with mean m = invlink(f)  (:335-376)."""
    scale = param_value('scale')

    def __init__(self, invlink=probit, scale=1.0):
        super().__init__(invlink)
        self._param('scale', scale, transforms.positive)

    def _logdensity(self, m, Y):
        a = m * self.scale
        return densities.beta(a, self.scale - a, Y)

    def _mean(self, m):
        return m

    def _var(self, m):
        return (m - m ** 2) / (self.scale + 1.0)


class StudentT(Likelihood):
    """y = f + scale * t_{deg_free}  (:244-265)."""
    scale = param_value('scale')

    def __init__(self, deg_free=3.0):
        super().__init__()
        self.deg_free = deg_free
        self._param('scale', 1.0, transforms.positive)

    def logp(self, F, Y):
        return densities.student_t(Y, F, self.scale, self.deg_free)

    def conditional_mean(self, F):
        return F

    def conditional_variance(self, F):
        return F * 0.0 + (self.deg_free / (self.deg_free - 2.0))


class RobustMax(object):
    """Multi-class inverse link: y_i = 1 - eps if i = argmax f, else eps / (k - 1)  (:379-424)."""

    def __init__(self, num_classes, epsilon=1e-3):
        self.epsilon = epsilon
        self.num_classes = num_classes
        self._eps_K1 = self.epsilon / (self.num_classes - 1.0)

    def __call__(self, F):
        i = torch.argmax(F, 1)
        hot = torch.nn.functional.one_hot(i, self.num_classes).to(F.dtype)
        return hot * (1.0 - self.epsilon) + (1.0 - hot) * self._eps_K1

    def prob_is_largest(self, Y, mu, var, gh_x, gh_w):
        """P(f_Y is the largest latent) under independent N(mu_k, var_k): Gauss-Hermite over
        f_Y of the product of the other latents' CDFs (:397-424)."""
        Yi = Y.reshape(-1).to(torch.int64)
        oh_on = torch.nn.functional.one_hot(Yi, self.num_classes).to(mu.dtype)
        mu_selected = (oh_on * mu).sum(1)
        var_selected = (oh_on * var).sum(1)
        gx = _const(gh_x, mu)
        X = mu_selected.reshape(-1, 1) + gx * torch.sqrt(torch.clamp(2.0 * var_selected, 1e-10, np.inf)).reshape(-1, 1)
        dist = (X.unsqueeze(1) - mu.unsqueeze(2)) / torch.sqrt(torch.clamp(var, 1e-10, np.inf)).unsqueeze(2)
        cdfs = 0.5 * (1.0 + torch.erf(dist / np.sqrt(2.0)))
        cdfs = cdfs * (1 - 2e-4) + 1e-4
        oh_off = 1.0 - oh_on
        cdfs = cdfs * oh_off.unsqueeze(2) + oh_on.unsqueeze(2)
        return cdfs.prod(1) @ _const(gh_w / np.sqrt(np.pi), mu).reshape(-1, 1)


class MultiClass(Likelihood):
    """Multi-way classification with the RobustMax link (:427-489)."""

    def __init__(self, num_classes, invlink=None):
        super().__init__()
        self.num_classes = num_classes
        if invlink is None:
            invlink = RobustMax(self.num_classes)
        elif not isinstance(invlink, RobustMax):
            raise NotImplementedError
        self.invlink = invlink

    def logp(self, F, Y):
        hits = torch.argmax(F, 1).unsqueeze(1) == Y.to(torch.int64)
        yes = torch.ones(Y.shape, dtype=F.dtype, device=F.device) - self.invlink.epsilon
        no = torch.zeros(Y.shape, dtype=F.dtype, device=F.device) + self.invlink._eps_K1
        return torch.log(torch.where(hits, yes, no))

    def variational_expectations(self, Fmu, Fvar, Y):
        gh_x, gh_w = hermgauss(self.num_gauss_hermite_points)
        p = self.invlink.prob_is_largest(Y, Fmu, Fvar, gh_x, gh_w)
        return p * float(np.log(1 - self.invlink.epsilon)) + (1.0 - p) * float(np.log(self.invlink._eps_K1))

    def predict_mean_and_var(self, Fmu, Fvar):
        n = Fmu.shape[0]
        ps = [self._predict_non_logged_density(
            Fmu, Fvar, torch.full((n, 1), i, dtype=torch.int64, device=Fmu.device)).reshape(-1)
            for i in range(self.num_classes)]
        ps = torch.stack(ps).t()
        return ps, ps - ps ** 2

    def predict_density(self, Fmu, Fvar, Y):
        return torch.log(self._predict_non_logged_density(Fmu, Fvar, Y))

    def _predict_non_logged_density(self, Fmu, Fvar, Y):
        gh_x, gh_w = hermgauss(self.num_gauss_hermite_points)
        p = self.invlink.prob_is_largest(Y, Fmu, Fvar, gh_x, gh_w)
        return p * (1 - self.invlink.epsilon) + (1.0 - p) * self.invlink._eps_K1

    def conditional_mean(self, F):
        return self.invlink(F)

    def conditional_variance(self, F):
        p = self.conditional_mean(F)
        return p - p ** 2


class SwitchedLikelihood(Likelihood):
    """Declared but not implemented in the reference (likelihoods.py:491-551)."""
    pass


class Ordinal(Likelihood):
    """Ordinal regression (Chu & Ghahramani 2005; likelihoods.py:554-631): labels 0..K with
    K-1... bin edges a_k, p(Y = k | F) = phi((a_k - F) / sigma) - phi((a_{k-1} - F) / sigma)."""

    def __init__(self, bin_edges):
        Likelihood.__init__(self)
        self.bin_edges = np.asarray(bin_edges, dtype=np.float64)
        self.num_bins = self.bin_edges.size + 1
        self._param('sigma', 1.0, transforms.positive)

    sigma = param_value('sigma')

    def _scaled_bins(self, like):
        edges = _const(self.bin_edges, like) / self.sigma
        inf = _const(np.array([np.inf]), like)
        return torch.cat([edges, inf], 0), torch.cat([-inf, edges], 0)

    def logp(self, F, Y):
        left, right = self._scaled_bins(F)
        idx = Y.to(torch.int64)
        return torch.log(probit(left[idx] - F / self.sigma) - probit(right[idx] - F / self.sigma) + 1e-6)

    def _make_phi(self, F):
        """[numel(F), num_bins] label probabilities, F flattened (:600-611)."""
        left, right = self._scaled_bins(F)
        Fs = F.reshape(-1, 1) / self.sigma
        return probit(left - Fs) - probit(right - Fs)

    def conditional_mean(self, F):
        phi = self._make_phi(F)
        Ys = _const(np.arange(self.num_bins, dtype=np.float64), F).reshape(-1, 1)
        return (phi @ Ys).reshape(F.shape)

    def conditional_variance(self, F):
        phi = self._make_phi(F)
        Ys = _const(np.arange(self.num_bins, dtype=np.float64), F).reshape(-1, 1)
        E_y = phi @ Ys
        E_y2 = phi @ Ys ** 2
        return (E_y2 - E_y ** 2).reshape(F.shape)
