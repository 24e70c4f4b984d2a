"""TEST INFRASTRUCTURE ONLY -- a minimal TensorFlow-1.x API shim backed by torch CPU float64.

Why this exists
---------------
The reference (ssydasheng/GPflow-Slim, /root/reference) is pure Python that only *sequences*
TensorFlow 1.x ops.  TensorFlow 1.x cannot be installed here (no
network, no wheel for Python 3.12), so the reference cannot run as shipped.  This package is
named ``tensorflow`` and implements just the ~70 ``tf.*`` entry points the reference's GP hot
path calls, each as the documented TF-1.x semantics of that op, evaluated eagerly with torch
CPU float64.  With ``oracle/tf_shim`` first on ``sys.path`` the UNMODIFIED reference package
imports and runs; ``oracle/gen_golden.py`` uses that to write the golden vectors under
``tests/golden/``.  Gradients come from ``torch.autograd`` (the Cholesky / triangular-solve
adjoints are mathematically the ones TF registers).

What this pins and what it does not: every line of gpflowSlim's own code (kernels, models,
conditionals, transforms, NKN wrappers ...) is executed as written; the arithmetic inside the
individual tf ops is torch/LAPACK's, not Eigen's.  Nothing in the product imports this.
"""
import contextlib
import types

import numpy as np
import torch

torch.set_default_dtype(torch.float64)

Tensor = torch.Tensor
Variable = torch.Tensor
float64 = np.float64
float32 = np.float32
int32 = np.int32
int64 = np.int64


class DType(object):  # only used in isinstance() checks (misc.py:75)
    pass


_NP2T = {np.float64: torch.float64, np.float32: torch.float32, np.int32: torch.int32,
         np.int64: torch.int64, float: torch.float64, int: torch.int64}


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, torch.dtype):
        return dtype
    if dtype in _NP2T:
        return _NP2T[dtype]
    return _NP2T[np.dtype(dtype).type]


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(_dt(dtype))
    if isinstance(x, (list, tuple)) and any(isinstance(v, torch.Tensor) for v in x):
        return torch.stack([_t(v) for v in x])
    a = np.asarray(x)
    if dtype is None and a.dtype.kind == 'f':
        dtype = np.float64
    return torch.as_tensor(a, dtype=_dt(dtype))


def _ints(shape):
    if isinstance(shape, torch.Tensor):
        return [int(v) for v in shape.reshape(-1).tolist()]
    if isinstance(shape, (int, np.integer)):
        return [int(shape)]
    return [int(v) for v in shape]


class _StaticShape(object):
    def __init__(self, shp):
        self._s = [int(v) for v in shp]
        self.ndims = len(self._s)

    def as_list(self):
        return list(self._s)


# TF tensors expose a static shape; the reference calls q_sqrt.get_shape().ndims
# (conditionals.py:105, kullback_leiblers.py:56).
torch.Tensor.get_shape = lambda self: _StaticShape(self.shape)




def _assign(self, value):
    """tf.Variable.assign in eager mode (LBFGS.py:74): overwrite the variable's storage."""
    with torch.no_grad():
        self.copy_(_t(value).reshape(self.shape))
    return self


torch.Tensor.assign = _assign


# ---------------------------------------------------------------- variables / scopes
_VARIABLES = []  # every tf.get_variable, in creation order


class GraphKeys(object):
    TRAINABLE_VARIABLES = 'trainable_variables'
    GLOBAL_VARIABLES = 'variables'


def get_variable(name, initializer=None, trainable=True, **kw):
    v = _t(initializer).detach().clone().to(torch.float64)
    v.requires_grad_(bool(trainable))
    v.tf_name = name
    _VARIABLES.append(v)
    return v


def shim_variables():
    return list(_VARIABLES)


def shim_reset():
    del _VARIABLES[:]


@contextlib.contextmanager
def variable_scope(name=None, *a, **kw):
    yield


@contextlib.contextmanager
def name_scope(name=None, *a, **kw):
    yield


@contextlib.contextmanager
def control_dependencies(deps):
    yield


# ---------------------------------------------------------------- constructors
def constant(value, dtype=None, name=None, shape=None):
    return _t(value, dtype)


def convert_to_tensor(value, dtype=None, name=None):
    return _t(value, dtype)


def cast(x, dtype, name=None):
    return _t(x).to(_dt(dtype))


def eye(n, dtype=None, **kw):
    return torch.eye(int(n), dtype=_dt(dtype) or torch.float64)


def zeros(shape, dtype=None, **kw):
    return torch.zeros(_ints(shape), dtype=_dt(dtype) or torch.float64)


def ones(shape, dtype=None, **kw):
    return torch.ones(_ints(shape), dtype=_dt(dtype) or torch.float64)


def zeros_like(x, dtype=None, **kw):
    return torch.zeros_like(_t(x), dtype=_dt(dtype))


def ones_like(x, dtype=None, **kw):
    return torch.ones_like(_t(x), dtype=_dt(dtype))


def fill(dims, value, name=None):
    return torch.ones(_ints(dims), dtype=torch.float64) * _t(value)


def range(*a, **kw):  # noqa: A001
    return torch.arange(*a)


def identity(x, name=None):
    return _t(x)


# ---------------------------------------------------------------- shape ops
def shape(x, **kw):
    return torch.tensor(list(_t(x).shape), dtype=torch.int64)


def size(x, **kw):
    return torch.tensor(_t(x).numel(), dtype=torch.int64)


def rank(x, **kw):
    return _t(x).dim()


def reshape(x, shp, name=None):
    return _t(x).reshape(_ints(shp))


def transpose(x, perm=None, name=None):
    x = _t(x)
    if perm is None:
        perm = list(reversed(builtins_range(x.dim())))
    return x.permute(*[int(p) for p in perm])


def matrix_transpose(x, name=None):
    return _t(x).transpose(-1, -2)


def expand_dims(x, axis, name=None):
    return _t(x).unsqueeze(int(axis))


def squeeze(x, axis=None, name=None):
    x = _t(x)
    return x.squeeze() if axis is None else x.squeeze(int(axis))


def stack(values, axis=0, name=None):
    def _int_scalar(v):
        if isinstance(v, torch.Tensor):
            return v.dim() == 0 and v.dtype in (torch.int64, torch.int32)
        return isinstance(v, (int, np.integer))
    if all(_int_scalar(v) for v in values):
        return torch.tensor([int(v) for v in values], dtype=torch.int64)  # shape vectors
    return torch.stack([_t(v) for v in values], dim=int(axis))


def concat(values, axis, name=None):
    vals = [_t(v) for v in values]
    return torch.cat(vals, dim=int(axis))


def tile(x, multiples, name=None):
    return _t(x).repeat(*_ints(multiples))


def gather(x, indices, axis=0, name=None):
    # tf.gather: result shape = x.shape[:axis] + indices.shape + x.shape[axis+1:]
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.int64)
    x, axis = _t(x), int(axis)
    out = torch.index_select(x, axis, idx.reshape(-1))
    return out.reshape(tuple(x.shape[:axis]) + tuple(idx.shape) + tuple(x.shape[axis + 1:]))


def scatter_nd(indices, updates, shape, name=None):  # noqa: A002
    out = torch.zeros(_ints(shape), dtype=_t(updates).dtype)
    idx = _t(indices).to(torch.int64)
    return out.index_put(tuple(idx[:, k] for k in builtins_range(idx.shape[1])), _t(updates),
                         accumulate=True)


def map_fn(fn, elems, **kw):
    return torch.stack([fn(e) for e in _t(elems)])


# ---------------------------------------------------------------- elementwise
def add(a, b, name=None):
    return _t(a) + _t(b)


def add_n(xs, name=None):
    out = _t(xs[0])
    for v in xs[1:]:
        out = out + _t(v)
    return out


def multiply(a, b, name=None):
    return _t(a) * _t(b)


def negative(x, name=None):
    return -_t(x)


def square(x, name=None):
    return _t(x) ** 2


def sqrt(x, name=None):
    return torch.sqrt(_t(x))


def exp(x, name=None):
    return torch.exp(_t(x))


def log(x, name=None):
    return torch.log(_t(x))


def sin(x, name=None):
    return torch.sin(_t(x))


def cos(x, name=None):
    return torch.cos(_t(x))


def acos(x, name=None):
    return torch.acos(_t(x))


def abs(x, name=None):  # noqa: A001
    return torch.abs(_t(x))


def pow(x, y, name=None):  # noqa: A001
    return torch.pow(_t(x), _t(y))


def lgamma(x, name=None):
    return torch.lgamma(_t(x))


def erf(x, name=None):
    return torch.erf(_t(x))


def equal(a, b, name=None):
    return _t(a) == _t(b)


def where(c, a, b, name=None):
    return torch.where(c, _t(a), _t(b))


def clip_by_value(x, lo, hi, name=None):
    # TF semantics: gradient is passed only where lo <= x <= hi (torch.clamp agrees).
    return torch.clamp(_t(x), min=float(lo), max=float(hi))


# ---------------------------------------------------------------- reductions
def _axis(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return [int(a) for a in axis]
    return int(axis)


def reduce_sum(x, axis=None, keepdims=False, name=None, keep_dims=None):
    x = _t(x)
    kd = bool(keepdims or keep_dims)
    return x.sum() if axis is None else x.sum(dim=_axis(axis), keepdim=kd)


def reduce_prod(x, axis=None, keepdims=False, name=None, reduction_indices=None):
    x = _t(x)
    if axis is None and reduction_indices is not None:      # TF-1.x alias (likelihoods.py:424)
        axis = reduction_indices[0] if isinstance(reduction_indices, (list, tuple)) else reduction_indices
    return x.prod() if axis is None else x.prod(dim=int(axis), keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=_axis(axis), keepdim=keepdims)


def argmax(x, axis=0, output_type=None, name=None):
    """tf.argmax: index of the largest entry along `axis` (first one on ties, like torch)."""
    out = torch.argmax(_t(x), dim=int(axis))
    return out if output_type is None else out.to(_dt(output_type))


def one_hot(indices, depth, on_value=1.0, off_value=0.0, axis=-1, dtype=None, name=None):
    """tf.one_hot with the new axis last: on_value at the index, off_value elsewhere."""
    idx = _t(indices).to(torch.int64)
    hot = torch.nn.functional.one_hot(idx, int(depth)).to(torch.float64)
    out = hot * float(on_value) + (1.0 - hot) * float(off_value)
    return out if dtype is None else out.to(_dt(dtype))


def reduce_max(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return x.max() if axis is None else x.max(dim=int(axis), keepdim=keepdims)[0]


def norm(x, **kw):
    return torch.sqrt((_t(x) ** 2).sum())


def trace(x, name=None):
    return torch.diagonal(_t(x), dim1=-2, dim2=-1).sum(-1)


# ---------------------------------------------------------------- linear algebra
def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


def cholesky(x, name=None):
    return torch.linalg.cholesky(_t(x))


def matrix_triangular_solve(matrix, rhs, lower=True, adjoint=False, name=None):
    m, r = _t(matrix), _t(rhs)
    if adjoint:
        m, lower = m.transpose(-1, -2), not lower
    # TF reads only the `lower` (or upper) triangle of `matrix`.
    m = torch.tril(m) if lower else torch.triu(m)
    return torch.linalg.solve_triangular(m, r, upper=not lower)


def matrix_inverse(x, name=None):
    return torch.linalg.inv(_t(x))


def matrix_diag_part(x, name=None):
    return torch.diagonal(_t(x), dim1=-2, dim2=-1)


diag_part = matrix_diag_part


def matrix_diag(x, name=None):
    return torch.diag_embed(_t(x))


def matrix_band_part(x, num_lower, num_upper, name=None):
    x = _t(x)
    if num_lower == -1 and num_upper == 0:
        return torch.tril(x)
    if num_lower == 0 and num_upper == -1:
        return torch.triu(x)
    raise NotImplementedError('matrix_band_part(%r, %r)' % (num_lower, num_upper))


def einsum(eq, *ops):
    return torch.einsum(eq, *[_t(o) for o in ops])


def assert_equal(x, y, message=None, **kw):
    if not bool(torch.all(_t(x) == _t(y))):
        raise ValueError('tf.assert_equal failed: %r' % (message,))
    return None


def cond(pred, true_fn, false_fn, **kw):
    return true_fn() if bool(pred) else false_fn()


def random_normal(shape, dtype=None, **kw):  # noqa: A002
    return torch.randn(_ints(shape), dtype=_dt(dtype) or torch.float64)


# ---------------------------------------------------------------- namespaces
nn = types.SimpleNamespace(softplus=lambda x, name=None: torch.nn.functional.softplus(_t(x)),
                           relu=lambda x, name=None: torch.relu(_t(x)))


class _TestCase(object):
    pass


test = types.SimpleNamespace(TestCase=_TestCase, main=lambda: None)


class _Unavailable(object):
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **kw):
        raise NotImplementedError('tf shim: %s is outside the GP hot path' % self._name)

    def __getattr__(self, item):
        return _Unavailable(self._name + '.' + item)


train = _Unavailable('tf.train')

import builtins as _b  # noqa: E402

builtins_range = _b.range


def __getattr__(name):  # module-level fallback: import-time references must not fail
    if name.startswith('__'):
        raise AttributeError(name)
    return _Unavailable('tf.' + name)
