"""Mean functions (reference mean_functions.py:24-193): m(X) evaluated next to the GP hot path
(`Y - m(X)` feeds the Cholesky solve, models/gpr.py:66-72, :123; `+ m(Xnew)` the predictions).
O(N D) elementwise work on the device; gradients by torch autograd."""
import numpy as np
import torch

from .params import Parameterized, param_value


class MeanFunction(Parameterized):
    """Base class: `__call__(X)` maps [N, D] inputs to [N, Q] means (mean_functions.py:24-54)."""

    def __init__(self, name='MeanFunction'):
        self._parameters = []
        self._name = name

    def __call__(self, X):
        raise NotImplementedError('Implement the __call__ method for this mean function')

    def __add__(self, other):
        return Additive(self, other)

    def __mul__(self, other):
        return Product(self, other)

    @property
    def parameters(self):
        return self._parameters

    @property
    def name(self):
        return self._name


class Zero(MeanFunction):
    """zeros [N, 1] (mean_functions.py:57-59); `output_dim` is an extension for multi-column
    broadcasting and defaults to the reference's single column."""

    def __init__(self, output_dim=1):
        super().__init__()
        self.output_dim = output_dim

    def __call__(self, X):
        return torch.zeros((X.shape[0], self.output_dim), dtype=X.dtype, device=X.device)


class Customized(MeanFunction):
    """The reference's hard-wired 1 -> 20 -> 20 -> 1 ReLU network mean (mean_functions.py:62-88):
    all-ones weights, zero biases."""

    def __init__(self):
        super().__init__()
        for i, (fan_in, fan_out) in enumerate(((1, 20), (20, 20), (20, 1)), start=1):
            self._param('W%d' % i, np.ones((fan_in, fan_out)))
            self._param('b%d' % i, np.zeros(fan_out))

    def __call__(self, X):
        h = torch.relu(X @ self._W1.value + self._b1.value)
        h = torch.relu(h @ self._W2.value + self._b2.value)
        return h @ self._W3.value + self._b3.value


class Linear(MeanFunction):
    """y_i = A x_i + b with A [D, Q], b [Q] (mean_functions.py:91-121)."""

    def __init__(self, A=None, b=None):
        A = np.ones((1, 1)) if A is None else A
        b = np.zeros(1) if b is None else b
        super().__init__()
        self._param('A', np.atleast_2d(A))
        self._param('b', b)

    A = param_value('A')
    b = param_value('b')

    def __call__(self, X):
        return X @ self.A + self.b


class Constant(MeanFunction):
    """y_i = c (mean_functions.py:124-141)."""

    def __init__(self, c=None):
        super().__init__()
        c = np.zeros(1) if c is None else c
        self._param('c', c)

    c = param_value('c')

    def __call__(self, X):
        return self.c.reshape(1, -1).repeat(X.shape[0], 1)


class SwitchedMeanFunction(MeanFunction):
    """Declared but not implemented in the reference (mean_functions.py:144-171)."""
    pass


class _Pair(MeanFunction):
    """The reference's Additive / Product keep an empty parameter list (:174-193) and rely on
    TensorFlow's global trainable-variable collection for training; here the parts' parameters
    are listed so that optimisers and `model.parameters` see them."""

    def __init__(self, first_part, second_part):
        super().__init__()
        self._first, self._second = first_part, second_part
        self._parameters = self._parameters + list(first_part.parameters) + list(second_part.parameters)


class Additive(_Pair):
    def __init__(self, first_part, second_part):
        super().__init__(first_part, second_part)
        self.add_1, self.add_2 = first_part, second_part

    def __call__(self, X):
        return self.add_1(X) + self.add_2(X)


class Product(_Pair):
    def __init__(self, first_part, second_part):
        super().__init__(first_part, second_part)
        self.prod_1, self.prod_2 = first_part, second_part

    def __call__(self, X):
        return self.prod_1(X) * self.prod_2(X)
