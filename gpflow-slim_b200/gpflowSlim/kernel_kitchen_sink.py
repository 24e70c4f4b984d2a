"""Random-feature ("kitchen sink") kernels: the part of the reference's kernel_kitchen_sink.py
that feeds GPR's feature path (models/gpr.py:62-66, :84-114): a `Sampler` maps X [N, D] to
features C [N, n_components] with K ~= C C^T, and `SamplerKernel` exposes `features(X)`, which
switches GPR to the Woodbury form (an n_components x n_components Cholesky instead of N x N).

Built here: Sampler, RBFSampler (:56-118), LinearSampler (:193-207), CosineSampler (:210-240),
ConstantSampler (:304-318), SamplerKernel (:694-710).  The projections X W and the Gram C C^T
run on the library's FP64 tensor-core GEMM; cos / scaling are elementwise.  The approximate
sum / product / sketch combinators of the reference are not built.

Like the reference, random draws come from numpy's GLOBAL RNG at construction.  The reference
divides the RBF draw by the lengthscale tensor once, at graph-construction time, so the
lengthscale stays differentiable; here the raw draw is kept and divided on every transform,
which is the same function of the current lengthscale."""
import numpy as np
import torch

from . import transforms
from ._backend import ops as _ops
from .kernels import Kernel
from .misc import to_tensor
from .params import Parameter


class Sampler(object):
    def __init__(self, input_dim, n_components):
        self.input_dim = input_dim
        self.n_components = n_components

    def check_dim(self, X):
        if X.shape[1] != self.input_dim:
            raise ValueError('input dimension not compatible with the init value')

    def transform(self, X):
        """[N, input_dim] -> [N, n_components]."""
        raise NotImplementedError

    @property
    def parameters(self):
        """Every Parameter held by the sampler (the reference leaves them to TensorFlow's
        global trainable-variable collection)."""
        return [v for _, v in sorted(vars(self).items()) if isinstance(v, Parameter)]


class RBFSampler(Sampler):
    """Random Fourier features of sigma^2 exp(-|x-x'|^2 / (2 l^2)) (Rahimi & Recht):
    sqrt(2 sigma^2 / C) cos(X W / l + b), W ~ N(0, 1), b ~ U(0, 2 pi)."""

    def __init__(self, input_dim, ls=1., var=1., n_components=100, scope='RBFSampler'):
        self._ls = Parameter(ls, transform=transforms.positive, name='ls')
        self._variance = Parameter(var, transform=transforms.positive, name='variance')
        self._normal_draw = to_tensor(np.random.normal(size=(input_dim, n_components)))
        self.random_offset_ = to_tensor(np.random.uniform(0, 2 * np.pi, size=n_components))
        super().__init__(input_dim, n_components)

    @property
    def ls(self):
        return self._ls.value

    @property
    def variance(self):
        return self._variance.value

    @property
    def random_weights_(self):
        return self._normal_draw / self.ls.reshape(-1, 1)

    def transform(self, X):
        X = to_tensor(X)
        self.check_dim(X)
        projection = _ops.matmul(X, self.random_weights_) + self.random_offset_
        feature = torch.cos(projection) * np.sqrt(2. / self.n_components)
        return feature * (self.variance ** 0.5)


class LinearSampler(Sampler):
    """Exact features of the linear kernel: X tiled to n_components columns, rescaled."""

    def __init__(self, input_dim, var=1., n_components=None, scope='LinearSampler'):
        self._variance = Parameter(var, transform=transforms.positive, name='variance')
        super().__init__(input_dim, n_components or input_dim)

    @property
    def variance(self):
        return self._variance.value

    def transform(self, X):
        X = to_tensor(X)
        self.check_dim(X)
        reps = int(np.ceil(self.n_components / self.input_dim))
        X_tile = torch.cat([X for _ in range(reps)], dim=-1)
        return X_tile[:, :self.n_components] * (self.variance * self.input_dim / float(self.n_components)) ** 0.5


class CosineSampler(Sampler):
    """[cos(X w / l), sin(X w / l)] tiled: exact features of the Cosine kernel."""

    def __init__(self, input_dim, ls=1., var=1., n_components=2, scope='CosineSampler'):
        self._ls = Parameter(ls, transform=transforms.positive, name='ls')
        self._variance = Parameter(var, transform=transforms.positive, name='variance')
        self._weights = Parameter(np.random.normal(size=[input_dim, 1]), name='weights')
        super().__init__(input_dim, n_components)

    @property
    def variance(self):
        return self._variance.value

    @property
    def ls(self):
        return self._ls.value

    @property
    def weights(self):
        return self._weights.value

    def transform(self, X):
        X = to_tensor(X)
        self.check_dim(X)
        mul = ((X / self.ls) * self.weights.reshape(1, -1)).sum(1, keepdim=True)
        feat = torch.cat([torch.cos(mul), torch.sin(mul)], dim=-1)
        feat = torch.cat([feat for _ in range(int(np.ceil(self.n_components / 2)))], dim=-1)
        return feat[:, :self.n_components] * (self.variance * 2. / self.n_components) ** 0.5


class ConstantSampler(Sampler):
    def __init__(self, input_dim, var=1., n_components=1, scope='ConstantSampler'):
        self._variance = Parameter(var, transform=transforms.positive, name='variance')
        super().__init__(input_dim, n_components)

    @property
    def variance(self):
        return self._variance.value

    def transform(self, X):
        X = to_tensor(X)
        self.check_dim(X)
        feat = X.new_ones((X.shape[0], self.n_components))
        return feat * (self.variance / self.n_components) ** 0.5


class SamplerKernel(Kernel):
    """K = features(X) features(X2)^T (kernel_kitchen_sink.py:694-710).  A composed kernel
    (never part of a fused Gram program); its `features` method is what GPR looks for."""

    def __init__(self, sampler):
        self.sampler = sampler
        super().__init__(input_dim=sampler.input_dim)
        self._parameters = self._parameters + list(getattr(sampler, 'parameters', []))

    def K(self, X, X2=None, presliced=False):
        feat1 = self.sampler.transform(X)
        feat2 = self.sampler.transform(X if X2 is None else X2)
        return _ops.matmul_nt(feat1, feat2)

    def Kdiag(self, X, presliced=False):
        feat1 = self.sampler.transform(X)
        return (feat1 ** 2.).sum(-1)

    def features(self, X):
        return self.sampler.transform(X)
