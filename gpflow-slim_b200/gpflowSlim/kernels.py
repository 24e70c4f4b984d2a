"""Covariance functions with the reference's API (gpflowSlim/kernels.py), re-designed for the
fused B200 Gram kernel.

A kernel object does not compute anything in Python: `K` / `Kdiag` COMPILE the kernel
expression (primitives + Sum/Product/NKN composition) into a `gps_kernel_desc` once, gather
the constrained parameter values into one device vector `theta`, and make a single library
call that evaluates the whole expression per matrix element in registers
(csrc/gram.cu).  Gradients w.r.t. every parameter (and the inputs) come from the matching
backward kernel through `torch.autograd`.

Supported on the fused path: RBF, Exponential, Matern12/32/52 (ARD or
isotropic lengthscales), Linear (ARD or isotropic variance), Periodic, Sum, Product (with scalar
/ tensor constants), and NeuralKernelNetwork (neural_kernel_network/).

The remaining covariances of the reference (White, Constant/Bias,
RatQuad, Polynomial, Cosine, ArcCosine, Coregion, TPS) are COMPOSED kernels: their O(N M D)
inner products run on the library's FP64 tensor-core GEMM (`_ops.matmul_nt`) or reuse the fused
Linear Gram, and the O(N M) elementwise body is torch glue on the device with torch autograd for
the gradients.  They do not compile into a `gps_kernel_desc` (`_emit` raises
NotImplementedError), so models containing them take the op-by-op path (models/gpr.py) and a
Sum / Product evaluates its fusable children in ONE fused launch and combines the composed
children elementwise.
"""
from functools import reduce

import numpy as np
import torch

from . import transforms
from ._backend import lib as _lib
from ._backend import ops as _ops
from ._settings import SETTINGS as settings
from .misc import to_tensor
from .params import Parameter


class KernelTooLarge(NotImplementedError):
    """The expression does not fit the fused Gram kernel's static tables (include/gpslim_b200.h:
    GPS_MAX_PRIMS / DIMS / OPS / SLOTS, 160 features per point, 191 hyper-parameters, column
    indices <= 255).  A NotImplementedError on purpose: callers then evaluate the expression on
    the composed path (tensor-core GEMM + elementwise) instead of failing."""


_MAX_FEATURES, _MAX_THETA_FUSED, _MAX_COLUMN = 160, 191, 255


class _Builder(object):
    """Collects primitives / theta pieces / ops while a kernel expression is walked."""

    def __init__(self):
        self.prims = []      # (type, dims, ard, theta_off)
        self.pieces = []     # callables -> tensors
        self.n_theta = 0
        self.ops = []        # [op, result_id, a, b, c, d, n]; operands are ('p', i) / ('o', id)
        self.n_results = 0

    def theta(self, getter, size):
        off = self.n_theta
        self.pieces.append(getter)
        self.n_theta += int(size)
        return off

    def prim(self, ptype, dims, ard, getters_sizes):
        off = self.n_theta
        for g, s in getters_sizes:
            self.theta(g, s)
        self.prims.append((ptype, [int(d) for d in dims], int(bool(ard)), off))
        return ('p', len(self.prims) - 1)

    def op(self, op, a=None, b=0, c=0, d=0, n=1):
        rid = self.n_results
        self.n_results += n
        self.ops.append([op, rid, a, b, c, d, n])
        return ('o', rid)

    def finalize(self, out):
        P = len(self.prims)
        if P < 1:
            # e.g. a Constant on its own: nothing for the Gram kernel to evaluate per element
            raise KernelTooLarge('an expression without any input-dependent kernel is not worth a fused launch')
        if P > _lib.GPS_MAX_PRIMS or len(self.ops) > _lib.GPS_MAX_OPS \
                or P + self.n_results > _lib.GPS_MAX_SLOTS or self.n_theta > _MAX_THETA_FUSED:
            raise KernelTooLarge('kernel expression too large for the fused Gram kernel')
        nfeat = 0
        for ptype, dims, _, _ in self.prims:
            nfeat += len(dims) + 1 if ptype <= _lib.GPS_MATERN52 else \
                (len(dims) if ptype == _lib.GPS_LINEAR else 3 * len(dims))
            if len(dims) > _lib.GPS_MAX_DIMS or (dims and max(dims) > _MAX_COLUMN):
                raise KernelTooLarge('more than %d active dimensions (or a column index above %d) in one '
                                     'primitive' % (_lib.GPS_MAX_DIMS, _MAX_COLUMN))
        if nfeat > _MAX_FEATURES:
            raise KernelTooLarge('more than %d features per point' % _MAX_FEATURES)

        def slot(ref):
            return ref[1] if ref[0] == 'p' else P + ref[1]

        desc = _lib.gps_kernel_desc()
        desc.n_prims, desc.n_ops, desc.n_theta = P, len(self.ops), self.n_theta
        desc.out_slot = slot(out)
        for i, (ptype, dims, ard, off) in enumerate(self.prims):
            pr = desc.prims[i]
            pr.type, pr.ndims, pr.ard, pr.theta_off = ptype, len(dims), ard, off
            for k, dd in enumerate(dims):
                pr.dims[k] = dd
        for i, (op, rid, a, b, c, d, n) in enumerate(self.ops):
            o = desc.ops[i]
            o.op, o.dst, o.n = op, P + rid, n
            if op == _lib.GPS_OP_CONST:
                o.a = a
            elif op in (_lib.GPS_OP_ADD, _lib.GPS_OP_MUL):
                o.a, o.b = slot(a), slot(b)
            elif op == _lib.GPS_OP_COPY:
                o.a = slot(a)
            else:
                o.a, o.b, o.c, o.d = slot(a), b, c, d
        return _ops.KernelProgram(desc, self.pieces, self.n_theta)


class Kernel(object):
    """Base class: input_dim / active_dims handling (reference kernels.py:33-66, 217-253)."""

    def __init__(self, input_dim, active_dims=None, name=None):
        self._name = name
        self.input_dim = int(input_dim)
        if active_dims is None:
            self.active_dims = slice(input_dim)
        elif isinstance(active_dims, slice):
            self.active_dims = active_dims
        else:
            self.active_dims = np.array(active_dims, dtype=np.int32)
            assert len(active_dims) == input_dim
        self._parameters = []
        self._program = None
        self._fusable = None

    # -- reference surface ------------------------------------------------------------
    @property
    def parameters(self):
        return self._parameters

    def __add__(self, other):
        return Sum([self, other])

    def __mul__(self, other):
        return Product([self, other])

    def compute_K(self, X, Z):
        return self.K(X, Z)

    def compute_K_symm(self, X):
        return self.K(X)

    def compute_Kdiag(self, X):
        return self.Kdiag(X)

    def _use_fused(self, X, X2=None):
        """The fused Gram kernel applies when the expression compiles AND, if a gradient w.r.t.
        the inputs is wanted, X has at most GPS_MAX_DIMS columns (the d/dX accumulators of the
        backward kernel are per column)."""
        if not self.fusable:
            return False
        wants_dx = X.requires_grad or (X2 is not None and X2.requires_grad)
        return not (wants_dx and X.shape[1] > _lib.GPS_MAX_DIMS)

    def K(self, X, X2=None, presliced=False):
        """Gram matrix [N, M]: one fused kernel launch (`presliced` is accepted for API
        compatibility -- slicing happens inside the CUDA kernel via the active-dims table), or
        the composed evaluation when the expression does not fit the fused kernel."""
        X = to_tensor(X)
        X2 = None if X2 is None else to_tensor(X2)
        if self._use_fused(X, X2):
            return _ops.gram(self.program(presliced), X, X2)
        return self._K_composed(X, X2, presliced)

    def Kdiag(self, X, presliced=False):
        X = to_tensor(X)
        if self._use_fused(X):
            return _ops.kdiag(self.program(presliced), X)
        return self._Kdiag_composed(X, presliced)

    def _K_composed(self, X, X2, presliced):
        raise NotImplementedError('%s has no composed evaluation' % type(self).__name__)

    def _Kdiag_composed(self, X, presliced):
        raise NotImplementedError('%s has no composed evaluation' % type(self).__name__)

    def K_jittered(self, X, jitter):
        """K(X) + jitter I (features.py:74-77, conditionals.py:60).  Fused kernels add the
        jitter in the Gram kernel's diagonal epilogue; composed kernels add it elementwise."""
        X = to_tensor(X)
        if self._use_fused(X):
            return _ops.gram(self.program(), X, None, diag_add=float(jitter))
        return self.K(X) + torch.eye(X.shape[0], dtype=X.dtype, device=X.device) * float(jitter)

    # -- compilation ---------------------------------------------------------------------
    def _dims(self, presliced=False):
        """Columns of X this kernel reads (kernels.py:217-253)."""
        if presliced:
            return list(range(self.input_dim))
        if isinstance(self.active_dims, slice):
            sl = self.active_dims
            start, step = sl.start or 0, sl.step or 1
            stop = sl.stop if sl.stop is not None else start + self.input_dim * step
            return list(range(start, stop, step))
        return [int(d) for d in self.active_dims]

    def _emit(self, b, presliced=False):
        raise NotImplementedError('%s is outside the fused Gram path' % type(self).__name__)

    @property
    def fusable(self):
        """True when the whole expression compiles into one fused Gram program."""
        if self._fusable is None:
            try:
                self.program()
                self._fusable = True
            except NotImplementedError:
                self._fusable = False
        return self._fusable

    def program(self, presliced=False):
        """Compiled descriptor (cached: the expression structure is static)."""
        key = bool(presliced)
        if self._program is None:
            self._program = {}
        if key not in self._program:
            b = _Builder()
            out = self._emit(b, presliced)
            self._program[key] = b.finalize(out)
        return self._program[key]

    def _slice(self, X, X2):
        X = X[:, self._dims()]
        if X2 is not None:
            X2 = X2[:, self._dims()]
        return X, X2

    def Kdim(self, dim, X, X2=None):
        """Covariance along one input dimension: X [n, 1] is embedded into column `dim` of an
        all-zero [n, input_dim] matrix (kernels.py:287-306)."""
        def embed(x):
            x = to_tensor(x)
            out = x.new_zeros((x.shape[0], self.input_dim))
            out[:, dim:dim + 1] = x
            return out
        return self.K(embed(X), None if X2 is None else embed(X2))


class Static(Kernel):
    """Covariances that do not look at the input values; one `variance` parameter
    (kernels.py:308-325).  Composed kernels: O(N M) fills, no arithmetic worth a CUDA kernel."""

    def __init__(self, input_dim, variance=1.0, active_dims=None, name=None):
        super().__init__(input_dim, active_dims, name=name)
        self._variance = Parameter(variance, transform=transforms.positive, name='variance')
        self._parameters = self._parameters + [self._variance]

    @property
    def variance(self):
        return self._variance.value

    def Kdiag(self, X, presliced=False):
        X = to_tensor(X)
        return torch.ones_like(X[:, 0]) * self.variance


class White(Static):
    """variance * I for K(X), zeros for K(X, X2) (kernels.py:328-338)."""

    def K(self, X, X2=None, presliced=False):
        X = to_tensor(X)
        if X2 is None:
            d = torch.ones_like(X[:, 0]) * self.variance.squeeze()
            return torch.diag_embed(d)
        X2 = to_tensor(X2)
        return X.new_zeros((X.shape[0], X2.shape[0]))


class Constant(Static):
    """variance everywhere (kernels.py:341-350).  Fusable: it compiles to the constant op of the
    Gram program, exactly like the scalar in `kern + 0.37`."""

    def _emit(self, b, presliced=False):
        return b.op(_lib.GPS_OP_CONST, b.theta(lambda: self.variance, 1))

    def _K_composed(self, X, X2, presliced):
        m = X.shape[0] if X2 is None else X2.shape[0]
        return X.new_ones((X.shape[0], m)) * self.variance.squeeze()

    def _Kdiag_composed(self, X, presliced):
        return torch.ones_like(X[:, 0]) * self.variance

    def K(self, X, X2=None, presliced=False):
        return Kernel.K(self, X, X2, presliced)

    def Kdiag(self, X, presliced=False):
        return Kernel.Kdiag(self, X, presliced)


class Bias(Constant):
    """Another name for Constant (kernels.py:353-357)."""
    pass


class Stationary(Kernel):
    """variance (positive) + lengthscales (Log1pe(min_ls)); reference kernels.py:360-429."""
    _ptype = None

    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False,
                 min_ls=1e-6, name='kernel'):
        super().__init__(input_dim, active_dims, name=name)
        self._variance = Parameter(variance, transform=transforms.positive, name='variance')
        if ARD:
            if lengthscales is None:
                lengthscales = np.ones(input_dim, dtype=np.float64)
            else:
                lengthscales = lengthscales * np.ones(input_dim, dtype=np.float64)
        else:
            lengthscales = 1.0 if lengthscales is None else lengthscales
        self.ARD = ARD
        self._ls = Parameter(lengthscales, transform=transforms.Log1pe(min_ls), name='ls')
        self._parameters = self._parameters + [self._variance, self._ls]

    @property
    def variance(self):
        return self._variance.value

    @property
    def lengthscales(self):
        return self._ls.value

    def _emit(self, b, presliced=False):
        if self._ptype is None:
            return super()._emit(b, presliced)
        nls = self.input_dim if self.ARD else 1
        return b.prim(self._ptype, self._dims(presliced), self.ARD,
                      [(lambda: self.variance, 1), (lambda: self.lengthscales, nls)])

    # The two helpers below are the reference's public building blocks (kernels.py:408-426).
    # The fused primitives never call them (the distance lives in registers there); the
    # composed stationary kernels (RatQuad, TPS) and user code do.
    def square_dist(self, X, X2):
        """clip(-2 X X'^T + |x|^2 + |x'|^2, 0, inf) of X / lengthscales; the inner products
        run on the FP64 tensor-core GEMM."""
        X = X / self.lengthscales
        Xs = (X ** 2).sum(1)
        if X2 is None:
            dist = -2.0 * _ops.matmul_nt(X, X) + Xs.reshape(-1, 1) + Xs.reshape(1, -1)
            return torch.clamp(dist, min=0.0)
        X2 = X2 / self.lengthscales
        X2s = (X2 ** 2).sum(1)
        dist = -2.0 * _ops.matmul_nt(X, X2) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)
        return torch.clamp(dist, min=0.0)

    def euclid_dist(self, X, X2):
        return torch.sqrt(self.square_dist(X, X2) + 1e-12)

    # composed evaluation of the fused primitives (expressions that do not fit the fused kernel,
    # e.g. the 100 network features of the reference's examples/svgp.py): the squared distance via
    # the tensor-core GEMM, the radial profile elementwise -- op for op kernels.py:408-439, :562-610
    _profile = None          # k(d2) / variance

    def _K_composed(self, X, X2, presliced):
        X, X2 = self._sliced(X, X2, presliced)
        return self.variance * type(self)._profile(self.square_dist(X, X2))

    def _Kdiag_composed(self, X, presliced):
        return torch.ones_like(X[:, 0]) * self.variance

    def _sliced(self, X, X2, presliced):
        X = to_tensor(X)
        X2 = None if X2 is None else to_tensor(X2)
        return (X, X2) if presliced else self._slice(X, X2)

    def _dimwise(self, cls, dim):
        ls = self.lengthscales[dim] if self.ARD else self.lengthscales
        return cls(input_dim=1, variance=self.variance ** (1. / self.input_dim), lengthscales=ls,
                   name='%s_dimwise_%d' % (cls.__name__, dim))


class RBF(Stationary):
    """sigma^2 exp(-d^2/2) (kernels.py:432-439)."""
    _ptype = _lib.GPS_RBF
    _profile = staticmethod(lambda d2: torch.exp(-d2 / 2))

    def dimwise(self, dim):
        """One-dimensional factor of the product form (kernels.py:441-444)."""
        return self._dimwise(RBF, dim)


class RatQuad(Stationary):
    """sigma^2 (1 + d^2 / (2 alpha))^(-alpha) (kernels.py:447-471).  Composed kernel: the
    squared distance comes from `square_dist` (tensor-core GEMM), the power is elementwise."""

    def __init__(self, input_dim, alpha=1., variance=1.0, lengthscales=None, active_dims=None,
                 ARD=False, min_ls=1e-6, name='kernel'):
        super().__init__(input_dim=input_dim, variance=variance, lengthscales=lengthscales,
                         active_dims=active_dims, ARD=ARD, min_ls=min_ls, name=name)
        self._alpha = Parameter(alpha, transform=transforms.positive, name='alpha')
        self._parameters = self._parameters + [self._alpha]

    @property
    def alpha(self):
        return self._alpha.value

    def K(self, X, X2=None, presliced=False):
        X, X2 = self._sliced(X, X2, presliced)
        base = 1.0 + 0.5 * self.square_dist(X, X2) * (1.0 / self.alpha)
        return self.variance * torch.pow(base, -1.0 * self.alpha)

    def Kdiag(self, X, presliced=False):
        X = to_tensor(X)
        return torch.ones_like(X[:, 0]) * self.variance


class Exponential(Stationary):
    """sigma^2 exp(-r/2) (kernels.py:555-566)."""
    _ptype = _lib.GPS_EXPONENTIAL
    _profile = staticmethod(lambda d2: (lambda r: torch.exp(-0.5 * r))(torch.sqrt(d2 + 1e-12)))


class Matern12(Stationary):
    """sigma^2 exp(-r) (kernels.py:569-577)."""
    _ptype = _lib.GPS_MATERN12
    _profile = staticmethod(lambda d2: (lambda r: torch.exp(-r))(torch.sqrt(d2 + 1e-12)))

    def dimwise(self, dim):
        return self._dimwise(Matern12, dim)


class Matern32(Stationary):
    """sigma^2 (1 + sqrt3 r) exp(-sqrt3 r) (kernels.py:585-594)."""
    _ptype = _lib.GPS_MATERN32
    _profile = staticmethod(lambda d2: (lambda r: (1. + np.sqrt(3.) * r) * torch.exp(-np.sqrt(3.) * r))(torch.sqrt(d2 + 1e-12)))

    def dimwise(self, dim):
        return self._dimwise(Matern32, dim)


class Matern52(Stationary):
    """sigma^2 (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r) (kernels.py:601-610)."""
    _ptype = _lib.GPS_MATERN52
    _profile = staticmethod(lambda d2: (lambda r: (1.0 + np.sqrt(5.) * r + 5. / 3. * r ** 2) * torch.exp(-np.sqrt(5.) * r))(torch.sqrt(d2 + 1e-12)))

    def dimwise(self, dim):
        return self._dimwise(Matern52, dim)


class Cosine(Stationary):
    """sigma^2 cos(sum_d w_d (x_d - x'_d) / l_d) with free weights w drawn from numpy's global
    RNG (kernels.py:617-646).  Composed kernel: O(N D) projections, O(N M) cosine.  The
    reference divides the UNSLICED X by the lengthscales and slices afterwards (:634-636) but
    slices X2 first (:641); both orders are reproduced."""

    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False,
                 min_ls=1e-6, name='kernel'):
        super().__init__(input_dim, variance=variance, lengthscales=lengthscales,
                         active_dims=active_dims, ARD=ARD, min_ls=min_ls, name=name)
        self._weights = Parameter(np.random.normal(size=[input_dim, 1]), name='weights')
        self._parameters = self._parameters + [self._weights]

    @property
    def weights(self):
        return self._weights.value

    def K(self, X, X2=None, presliced=False):
        X = to_tensor(X) / self.lengthscales
        X2 = None if X2 is None else to_tensor(X2)
        if not presliced:
            X, X2 = self._slice(X, X2)
        w = self.weights.reshape(1, -1)
        prod = (X * w).sum(1)
        prod2 = prod if X2 is None else ((X2 / self.lengthscales) * w).sum(1)
        r = prod.reshape(-1, 1) - prod2.reshape(1, -1)
        return self.variance * torch.cos(r)

    def Kdiag(self, X, presliced=False):
        X = to_tensor(X)
        return torch.ones_like(X[:, 0]) * self.variance


class ArcCosine(Kernel):
    """Arc-cosine kernel of order 0 / 1 / 2 (Cho & Saul 2009; kernels.py:649-766).  Composed
    kernel: the weighted inner products run on the FP64 tensor-core GEMM, the angle / J
    function is elementwise."""

    implemented_orders = {0, 1, 2}

    def __init__(self, input_dim, order=0, variance=1.0, weight_variances=1., bias_variance=1.0,
                 active_dims=None, ARD=False, name='kernel'):
        super().__init__(input_dim, active_dims, name=name)
        if order not in self.implemented_orders:
            raise ValueError('Requested kernel order is not implemented.')
        self.order = order
        self._variance = Parameter(variance, transform=transforms.positive, name='variance')
        self._bias_variance = Parameter(bias_variance, transform=transforms.positive,
                                        name='bias_variance')
        if ARD:
            if weight_variances is None:
                weight_variances = np.ones(input_dim, dtype=np.float64)
            else:
                weight_variances = weight_variances * np.ones(input_dim, dtype=np.float64)
        elif weight_variances is None:
            weight_variances = 1.0
        self.ARD = ARD
        self._weight_variances = Parameter(weight_variances, transform=transforms.positive,
                                           name='weight_variances')
        self._parameters = self._parameters + [self._variance, self._bias_variance,
                                               self._weight_variances]

    @property
    def variance(self):
        return self._variance.value

    @property
    def bias_variance(self):
        return self._bias_variance.value

    @property
    def weight_variances(self):
        return self._weight_variances.value

    def _weighted_product(self, X, X2=None):
        if X2 is None:
            return (self.weight_variances * X ** 2).sum(1) + self.bias_variance
        return _ops.matmul_nt(self.weight_variances * X, X2) + self.bias_variance

    def _J(self, theta):
        """Equations 4-7 of the paper (kernels.py:727-738)."""
        if self.order == 0:
            return np.pi - theta
        if self.order == 1:
            return torch.sin(theta) + (np.pi - theta) * torch.cos(theta)
        return 3. * torch.sin(theta) * torch.cos(theta) + \
            (np.pi - theta) * (1. + 2. * torch.cos(theta) ** 2)

    def K(self, X, X2=None, presliced=False):
        X = to_tensor(X)
        X2 = None if X2 is None else to_tensor(X2)
        if not presliced:
            X, X2 = self._slice(X, X2)
        X_denominator = torch.sqrt(self._weighted_product(X))
        if X2 is None:
            X2 = X
            X2_denominator = X_denominator
        else:
            X2_denominator = torch.sqrt(self._weighted_product(X2))
        numerator = self._weighted_product(X, X2)
        cos_theta = numerator / X_denominator[:, None] / X2_denominator[None, :]
        jitter = 1e-15
        theta = torch.acos(jitter + (1 - 2 * jitter) * cos_theta)
        return self.variance * (1. / np.pi) * self._J(theta) * \
            X_denominator[:, None] ** self.order * X2_denominator[None, :] ** self.order

    def Kdiag(self, X, presliced=False):
        X = to_tensor(X)
        if not presliced:
            X, _ = self._slice(X, None)
        X_product = self._weighted_product(X)
        theta = torch.zeros((), dtype=X.dtype, device=X.device)
        return self.variance * (1. / np.pi) * self._J(theta) * X_product ** self.order


class Linear(Kernel):
    """sum_d v_d x_d x'_d (kernels.py:474-510)."""

    def __init__(self, input_dim, variance=1.0, active_dims=None, ARD=False, name='kernel'):
        super().__init__(input_dim, active_dims, name=name)
        self.ARD = ARD
        variance = np.ones(self.input_dim, dtype=np.float64) * variance if ARD else variance
        self._variance = Parameter(variance, transform=transforms.positive, name='variance')
        self._parameters = self._parameters + [self._variance]

    @property
    def variance(self):
        return self._variance.value

    def _emit(self, b, presliced=False):
        return b.prim(_lib.GPS_LINEAR, self._dims(presliced), self.ARD,
                      [(lambda: self.variance, self.input_dim if self.ARD else 1)])

    def _K_composed(self, X, X2, presliced):
        if not presliced:
            X, X2 = self._slice(X, X2)
        return _ops.matmul_nt(X * self.variance, X if X2 is None else X2)       # kernels.py:503-505

    def _Kdiag_composed(self, X, presliced):
        if not presliced:
            X, _ = self._slice(X, None)
        return (X ** 2 * self.variance).sum(1)                                   # kernels.py:510

    def dimwise(self, dim):
        """kernels.py:512-515."""
        var = self.variance[dim] if self.ARD else self.variance ** (1. / self.input_dim)
        return Linear(input_dim=1, variance=var, name='Linear_dimwise_%d' % dim)


class Polynomial(Linear):
    """(Linear.K + offset) ** degree (kernels.py:518-554).  The linear part is the fused Linear
    Gram; offset and power are elementwise.  Like the reference (:541), `parameters` lists the
    variance twice (once from Linear.__init__, once here)."""

    def __init__(self, input_dim, degree=3.0, variance=1.0, offset=1.0, active_dims=None,
                 ARD=False, name='kernel'):
        super().__init__(input_dim, variance, active_dims, ARD, name=name)
        self.degree = degree
        self._offset = Parameter(offset, transform=transforms.positive, name='offset')
        self._parameters = self._parameters + [self._variance, self._offset]

    @property
    def offset(self):
        return self._offset.value

    def _emit(self, b, presliced=False):
        return Kernel._emit(self, b, presliced)

    def _linear_program(self, presliced):
        key = ('linear', bool(presliced))
        if self._program is None:
            self._program = {}
        if key not in self._program:
            b = _Builder()
            self._program[key] = b.finalize(Linear._emit(self, b, presliced))
        return self._program[key]

    def K(self, X, X2=None, presliced=False):
        X = to_tensor(X)
        X2 = None if X2 is None else to_tensor(X2)
        lin = _ops.gram(self._linear_program(presliced), X, X2)
        return (lin + self.offset) ** self.degree

    def Kdiag(self, X, presliced=False):
        lin = _ops.kdiag(self._linear_program(presliced), to_tensor(X))
        return (lin + self.offset) ** self.degree


class Periodic(Kernel):
    """sigma^2 exp(-1/2 sum_d sin^2(pi (x_d - x'_d)/p) / l^2), scalar l and p
    (kernels.py:769-819)."""

    def __init__(self, input_dim, period=1.0, variance=1.0, lengthscales=1.0, active_dims=None,
                 name='kernel'):
        super().__init__(input_dim, active_dims, name=name)
        self._variance = Parameter(variance, transform=transforms.positive, name='variance')
        self._ls = Parameter(lengthscales, transform=transforms.positive, name='ls')
        self._period = Parameter(period, transform=transforms.positive, name='period')
        self._parameters = self._parameters + [self._variance, self._ls, self._period]

    @property
    def variance(self):
        return self._variance.value

    @property
    def lengthscales(self):
        return self._ls.value

    @property
    def period(self):
        return self._period.value

    def _K_composed(self, X, X2, presliced):
        """sum_d sin^2(pi (x_d - x'_d) / p) = (D - sum_d cos(a_d - a'_d)) / 2 with a = 2 pi x / p
        (the identity the fused kernel uses): two tensor-core GEMMs on cos / sin features
        instead of the reference's N x M x D broadcast (kernels.py:806-819)."""
        if not presliced:
            X, X2 = self._slice(X, X2)
        a = 2.0 * np.pi * X / self.period
        a2 = a if X2 is None else 2.0 * np.pi * X2 / self.period
        cs = _ops.matmul_nt(torch.cos(a), torch.cos(a2)) + _ops.matmul_nt(torch.sin(a), torch.sin(a2))
        r = 0.5 * (float(X.shape[1]) - cs) / self.lengthscales ** 2
        return self.variance * torch.exp(-0.5 * r)

    def _Kdiag_composed(self, X, presliced):
        return torch.ones_like(X[:, 0]) * self.variance

    def _emit(self, b, presliced=False):
        return b.prim(_lib.GPS_PERIODIC, self._dims(presliced), False,
                      [(lambda: self.variance, 1), (lambda: self.lengthscales, 1),
                       (lambda: self.period, 1)])


class Coregion(Kernel):
    """K(x, y) = B[x, y], B = W W^T + diag(kappa), integer-valued 1-D inputs
    (kernels.py:822-881).  Composed kernel: a table lookup."""

    def __init__(self, input_dim, output_dim, rank, active_dims=None, name='kernel'):
        assert input_dim == 1, 'Coregion kernel in 1D only'
        super().__init__(input_dim, active_dims, name=name)
        self.output_dim = output_dim
        self.rank = rank
        self._W = Parameter(np.zeros((self.output_dim, self.rank), dtype=np.float64), name='W')
        self._kappa = Parameter(np.ones(self.output_dim, dtype=np.float64),
                                transform=transforms.positive, name='kappa')
        self._parameters = self._parameters + [self._W, self._kappa]

    @property
    def W(self):
        return self._W.value

    @property
    def kappa(self):
        return self._kappa.value

    def K(self, X, X2=None, presliced=False):
        X = to_tensor(X)
        X2 = None if X2 is None else to_tensor(X2)
        X, X2 = self._slice(X, X2)
        i = X[:, 0].to(torch.int64)                       # tf.cast(..., tf.int32) truncates
        j = i if X2 is None else X2[:, 0].to(torch.int64)
        B = self.W @ self.W.t() + torch.diag_embed(self.kappa)
        return B[j].t()[i]

    def Kdiag(self, X, presliced=False):
        X, _ = self._slice(to_tensor(X), None)
        i = X[:, 0].to(torch.int64)
        Bdiag = (self.W ** 2).sum(1) + self.kappa
        return Bdiag[i]


class TPS(Stationary):
    """Thin-plate-spline kernel sigma^2 (D^3 - 1.5 R D^2 + 0.5 R^3), R = 2, D the UNSCALED
    Euclidean distance: its square_dist ignores the lengthscales (kernels.py:943-970)."""

    def square_dist(self, X, X2):
        Xs = (X ** 2).sum(1)
        if X2 is None:
            dist = -2.0 * _ops.matmul_nt(X, X) + Xs.reshape(-1, 1) + Xs.reshape(1, -1)
            return torch.clamp(dist, min=0.0)
        X2s = (X2 ** 2).sum(1)
        dist = -2.0 * _ops.matmul_nt(X, X2) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)
        return torch.clamp(dist, min=0.0)

    @property
    def R(self):
        return torch.tensor(2., dtype=torch.float64, device=self._variance.vf_val.device)

    def K(self, X, X2=None, presliced=False):
        X, X2 = self._sliced(X, X2, presliced)
        D = torch.sqrt(self.square_dist(X, X2))
        return self.variance * (torch.pow(D, 3.) - 1.5 * self.R * D ** 2 + 0.5 * torch.pow(self.R, 3.))

    def Kdiag(self, X, presliced=False):
        X = to_tensor(X)
        return self.variance * 0.5 * torch.pow(self.R, 3.) * torch.ones_like(X[:, 0])


def make_kernel_names(kern_list):
    """Lower-cased class names, duplicates numbered from _1 (kernels.py:973-997)."""
    names, seen = [], {}
    for k in kern_list:
        base = k.__class__.__name__.lower()
        if base in seen:
            if seen[base] == 1:
                names[names.index(base)] = base + '_1'
            seen[base] += 1
            names.append('%s_%d' % (base, seen[base]))
        else:
            seen[base] = 1
            names.append(base)
    return names


class Combination(Kernel):
    """Sum / Product of kernels and scalar or tensor constants (kernels.py:1000-1063): nested
    combinations of the same class are flattened, constants are appended after the kernels."""

    def __init__(self, kern_list, name='kernel'):
        active = reduce(np.union1d,
                        (np.r_[k.active_dims] for k in kern_list if isinstance(k, Kernel)),
                        np.asarray([], dtype=int))
        super().__init__(input_dim=active.size, name=name, active_dims=active)
        self.kern_list, self.const_list = [], []
        for k in kern_list:
            if isinstance(k, self.__class__):
                self.kern_list.extend(k.kern_list)
                self.const_list.extend(k.const_list)
            elif isinstance(k, (int, float, np.floating, np.integer, torch.Tensor)):
                self.const_list.append(k)
            else:
                self.kern_list.append(k)
        # the reference also exposes the children as attributes named after their classes
        # (kernels.py:1030-1032)
        for nm, k in zip(make_kernel_names(self.kern_list), self.kern_list):
            setattr(self, nm, k)
        for kern in kern_list:
            if isinstance(kern, Kernel):
                self._parameters = self._parameters + kern.parameters
        self._fused_part = None

    @property
    def on_separate_dimensions(self):
        if np.any([isinstance(k.active_dims, slice) for k in self.kern_list]):
            return False
        dimlist = [k.active_dims for k in self.kern_list]
        for i, di in enumerate(dimlist):
            for dj in dimlist[i + 1:]:
                if np.any(di.reshape(-1, 1) == dj.reshape(1, -1)):
                    return False
        return True

    _op = None

    def _emit(self, b, presliced=False):
        # children always slice the FULL X themselves (Sum.K calls k.K(X, X2), kernels.py:1073)
        refs = [k._emit(b, False) for k in self.kern_list]
        for c in self.const_list:
            if isinstance(c, torch.Tensor):
                off = b.theta((lambda c=c: c), 1)
            else:
                off = b.theta((lambda c=c: float(c)), 1)
            refs.append(b.op(_lib.GPS_OP_CONST, off))
        out = refs[0]
        for r in refs[1:]:
            out = b.op(self._op, out, r)
        return out


    # -- mixed evaluation: some children are composed kernels ----------------------------
    def _split(self):
        """(fused sub-combination of all fusable children or None, composed children)."""
        if self._fused_part is None:
            fus = [k for k in self.kern_list if k.fusable]
            rest = [k for k in self.kern_list if not k.fusable]
            part = None
            if len(fus) == 1:
                part = fus[0]
            elif fus and not rest:
                # children fuse one by one: their combination does not fit, or it does but the
                # composed path was chosen because d/dX is wanted on inputs wider than the fused
                # kernel's column table (_use_fused) -- building Combination(fus) again would
                # reproduce this very expression and recurse without end
                part = self
            elif fus:
                part = self.__class__(fus)
                if not part.fusable:    # the fusable children together are too large: one by one
                    part, rest = None, list(self.kern_list)
            self._fused_part = (part, rest)
        return self._fused_part

    def _combine(self, vals, like):
        for c in self.const_list:
            vals.append(c if isinstance(c, torch.Tensor) else
                        torch.as_tensor(float(c), dtype=like.dtype, device=like.device))
        return reduce(self._torch_op, vals)

    def _K_composed(self, X, X2, presliced):
        part, rest = self._split()
        if X2 is not None and self._op == _lib.GPS_OP_ADD and len(self.kern_list) > 1:
            # a White term of a SUM contributes an all-zero cross-covariance (kernels.py:336-338):
            # skip the N x M zeros instead of allocating and adding them
            keep = [k for k in rest if type(k) is not White]
            if part is not None or keep:
                rest = keep
        if part is self:          # everything is fusable in principle, but the whole is too large / needs d/dX
            vals = [k.K(X, X2) for k in self.kern_list]
        else:
            vals = ([part.K(X, X2)] if part is not None else []) + [k.K(X, X2) for k in rest]
        return self._combine(vals, X)

    def _Kdiag_composed(self, X, presliced):
        part, rest = self._split()
        if part is self:
            vals = [k.Kdiag(X) for k in self.kern_list]
        else:
            vals = ([part.Kdiag(X)] if part is not None else []) + [k.Kdiag(X) for k in rest]
        return self._combine(vals, X)


class Sum(Combination):
    _op = _lib.GPS_OP_ADD
    _torch_op = staticmethod(torch.add)

    def split_white(self):
        """(kernel of everything but the top-level White terms or None, [White kernels]).  A
        White term only adds its variance to the diagonal of K(X), so the fused GPR objective can
        fold it into the noise it already adds there (models/gpr.py)."""
        whites = [k for k in self.kern_list if type(k) is White]
        rest = [k for k in self.kern_list if type(k) is not White]
        if not whites or not rest:
            return (self if not whites else None), whites
        if len(rest) == 1 and not self.const_list:
            return rest[0], whites
        return Sum(rest + list(self.const_list)), whites


class Product(Combination):
    _op = _lib.GPS_OP_MUL
    _torch_op = staticmethod(torch.mul)
