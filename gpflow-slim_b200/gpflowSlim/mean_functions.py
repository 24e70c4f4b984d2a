"""Mean functions on the hot path (reference mean_functions.py:28-59, :62-106)."""
import numpy as np
import torch

from .params import Parameter


class MeanFunction(object):
    def __init__(self):
        self._parameters = []

    def __call__(self, X):
        raise NotImplementedError

    @property
    def parameters(self):
        return self._parameters


class Zero(MeanFunction):
    """mean_functions.py:57-59: zeros [N, 1]."""

    def __init__(self, output_dim=1):
        super().__init__()
        self.output_dim = output_dim

    def __call__(self, X):
        return torch.zeros((X.shape[0], self.output_dim), dtype=X.dtype, device=X.device)


class Constant(MeanFunction):
    def __init__(self, c=None, name='constant_mean'):
        super().__init__()
        self.c = Parameter(np.zeros(1) if c is None else c, name=name)
        self._parameters = [self.c]

    def __call__(self, X):
        return self.c.value.reshape(1, -1).expand(X.shape[0], -1)


class Linear(MeanFunction):
    """y = X A + b."""

    def __init__(self, A=None, b=None, name='linear_mean'):
        super().__init__()
        self.A = Parameter(np.ones((1, 1)) if A is None else A, name=name + '_A')
        self.b = Parameter(np.zeros(1) if b is None else b, name=name + '_b')
        self._parameters = [self.A, self.b]

    def __call__(self, X):
        return X @ self.A.value + self.b.value
