"""Parameter transforms (reference transforms.py).  `forward_tensor` maps the unconstrained
torch leaf to the constrained value; `forward` / `backward` are the numpy maps used at
construction time (params.py:142).  Host-side, O(#parameters) work."""
import numpy as np
import torch

from ._settings import SETTINGS as settings
from .misc import vec_to_tri


def _softplus(x):
    # log(1 + exp(x)) without torch's linearisation threshold (tf.nn.softplus is exact)
    return torch.clamp(x, min=0.0) + torch.log1p(torch.exp(-torch.abs(x)))


class Transform(object):
    def forward(self, x):
        raise NotImplementedError

    def backward(self, y):
        raise NotImplementedError

    def forward_tensor(self, x):
        raise NotImplementedError

    def log_jacobian_tensor(self, x):
        raise NotImplementedError

    def __call__(self, other):
        if not isinstance(other, Transform):
            raise TypeError('transforms can only be chained with transforms')
        return Chain(self, other)


class Identity(Transform):
    """transforms.py:62-84."""

    def forward(self, x):
        return x

    def backward(self, y):
        return y

    def forward_tensor(self, x):
        return x

    def log_jacobian_tensor(self, x):
        return torch.zeros((), dtype=x.dtype, device=x.device)

    def __str__(self):
        return '(none)'


class Chain(Transform):
    """t1(t2(x)) (transforms.py:87-114)."""

    def __init__(self, t1, t2):
        self.t1, self.t2 = t1, t2

    def forward(self, x):
        return self.t1.forward(self.t2.forward(x))

    def backward(self, y):
        return self.t2.backward(self.t1.backward(y))

    def forward_tensor(self, x):
        return self.t1.forward_tensor(self.t2.forward_tensor(x))

    def log_jacobian_tensor(self, x):
        return self.t1.log_jacobian_tensor(self.t2.forward_tensor(x)) + self.t2.log_jacobian_tensor(x)

    def __str__(self):
        return '{} {}'.format(self.t1, self.t2)


class Exp(Transform):
    """y = exp(x) + lower (transforms.py:117-142 of the reference numbering: Exp class)."""

    def __init__(self, lower=1e-6):
        self._lower = lower

    def forward(self, x):
        return np.exp(x) + self._lower

    def backward(self, y):
        return np.log(y - self._lower)

    def forward_tensor(self, x):
        return torch.exp(x) + self._lower

    def log_jacobian_tensor(self, x):
        return x.sum()

    def __str__(self):
        return 'Exp'


class Log1pe(Transform):
    """Softplus: y = log(1 + exp(x)) + lower (transforms.py:117-181).  The +1e-6 is a parity
    trap: every `positive` parameter of the reference carries it."""

    def __init__(self, lower=1e-6):
        self._lower = lower

    def forward(self, x):
        return np.logaddexp(0.0, x) + self._lower

    def forward_tensor(self, x):
        return _softplus(x) + self._lower

    def log_jacobian_tensor(self, x):
        return -_softplus(-x).sum()

    def backward(self, y):
        ys = np.maximum(np.asarray(y, dtype=np.float64) - self._lower, np.finfo(np.float64).eps)
        return ys + np.log(-np.expm1(-ys))

    def __str__(self):
        return '+ve'


class Logistic(Transform):
    """y = a + (b - a) sigmoid(x)."""

    def __init__(self, a=0., b=1.):
        if a >= b:
            raise ValueError('a must be smaller than b')
        self.a, self.b = float(a), float(b)

    def forward(self, x):
        ex = np.exp(-x)
        return self.a + (self.b - self.a) / (1. + ex)

    def backward(self, y):
        return -np.log((self.b - self.a) / (y - self.a) - 1.)

    def forward_tensor(self, x):
        return self.a + (self.b - self.a) * torch.sigmoid(x)

    def log_jacobian_tensor(self, x):
        return (x - 2. * _softplus(x) + np.log(self.b - self.a)).sum()

    def __str__(self):
        return '[{}, {}]'.format(self.a, self.b)


class Rescale(Transform):
    """y = factor * x (transforms.py:215-251).  Chained like any other transform:
    `Rescale(s)(positive)` = Chain(Rescale(s), positive): y = s * (softplus(x) + 1e-6)."""

    def __init__(self, factor=1.0):
        self.factor = float(factor)

    def forward(self, x):
        return x * self.factor

    def backward(self, y):
        return y / self.factor

    def forward_tensor(self, x):
        return x * self.factor

    def log_jacobian_tensor(self, x):
        return x.numel() * torch.log(torch.as_tensor(self.factor, dtype=x.dtype, device=x.device))

    def __str__(self):
        return '{}*'.format(self.factor)


class DiagMatrix(Transform):
    """Vector <-> stack of diagonal matrices."""

    def __init__(self, dim=1):
        self.dim = dim

    def forward(self, x):
        x = np.reshape(x, (-1, self.dim))
        return np.stack([np.diag(r) for r in x])

    def backward(self, y):
        return np.stack([np.diag(m) for m in y]).reshape(-1)

    def forward_tensor(self, x):
        return torch.diag_embed(x.reshape(-1, self.dim))

    def log_jacobian_tensor(self, x):
        return torch.zeros((1,), dtype=x.dtype, device=x.device)

    def __str__(self):
        return 'DiagMatrix'


class LowerTriangular(Transform):
    """Free vector <-> [N, N, num_matrices] lower-triangular matrices (transforms.py:294-374,
    misc.vec_to_tri): row-major tril order, one row of the free state per matrix."""

    def __init__(self, N, num_matrices=1, squeeze=False):
        self.N, self.num_matrices, self.squeeze = N, num_matrices, squeeze

    def forward(self, x):
        x = np.asarray(x)
        xr = x.reshape(self.num_matrices, -1)
        n = int(np.floor(0.5 * np.sqrt(xr.shape[1] * 8. + 1.) - 0.5))
        if n * (n + 1) // 2 != xr.shape[1]:
            raise ValueError('The free state must be a triangle number.')
        var = np.zeros((n, n, self.num_matrices), settings.float_type)
        r, c = np.tril_indices(n)
        for i in range(self.num_matrices):
            var[r, c, i] = xr[i]
        return var.squeeze() if self.squeeze else var

    def backward(self, y):
        y = np.asarray(y)
        N = int(np.sqrt(y.size / self.num_matrices))
        reshaped = np.reshape(y, (N, N, self.num_matrices))
        return reshaped[np.tril_indices(N, 0)].T          # [num_matrices, N(N+1)/2]

    def forward_tensor(self, x):
        fwd = vec_to_tri(x.reshape(self.num_matrices, -1), self.N).permute(1, 2, 0)
        return fwd.squeeze() if self.squeeze else fwd

    def log_jacobian_tensor(self, x):
        return torch.zeros((1,), dtype=x.dtype, device=x.device)

    def __str__(self):
        return 'LoTri->vec'


positive = Log1pe()


def positiveRescale(scale):
    return Rescale(scale)(positive)
