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
            raise ValueError('a kernel expression needs at least one kernel')
        if P > _lib.GPS_MAX_PRIMS or len(self.ops) > _lib.GPS_MAX_OPS \
                or P + self.n_results > _lib.GPS_MAX_SLOTS:
            raise ValueError('kernel expression too large for the fused Gram kernel')

        def slot(ref):
            return ref[1] if ref[0] == 'p' else P + ref[1]

        desc = _lib.gps_kernel_desc()
        desc.n_prims, desc.n_ops, desc.n_theta = P, len(self.ops), self.n_theta
        desc.out_slot = slot(out)
        for i, (ptype, dims, ard, off) in enumerate(self.prims):
            pr = desc.prims[i]
            pr.type, pr.ndims, pr.ard, pr.theta_off = ptype, len(dims), ard, off
            if len(dims) > _lib.GPS_MAX_DIMS:
                raise ValueError('at most %d active dimensions per primitive' % _lib.GPS_MAX_DIMS)
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

    def K(self, X, X2=None, presliced=False):
        """Gram matrix [N, M] (one fused kernel launch; `presliced` is accepted for API
        compatibility -- slicing happens inside the CUDA kernel via the active-dims table)."""
        X = to_tensor(X)
        X2 = None if X2 is None else to_tensor(X2)
        return _ops.gram(self.program(presliced), X, X2)

    def Kdiag(self, X, presliced=False):
        return _ops.kdiag(self.program(presliced), to_tensor(X))

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
        nls = self.input_dim if self.ARD else 1
        return b.prim(self._ptype, self._dims(presliced), self.ARD,
                      [(lambda: self.variance, 1), (lambda: self.lengthscales, nls)])


class RBF(Stationary):
    """sigma^2 exp(-d^2/2) (kernels.py:432-439)."""
    _ptype = _lib.GPS_RBF


class Exponential(Stationary):
    """sigma^2 exp(-r/2) (kernels.py:555-566)."""
    _ptype = _lib.GPS_EXPONENTIAL


class Matern12(Stationary):
    """sigma^2 exp(-r) (kernels.py:569-577)."""
    _ptype = _lib.GPS_MATERN12


class Matern32(Stationary):
    """sigma^2 (1 + sqrt3 r) exp(-sqrt3 r) (kernels.py:585-594)."""
    _ptype = _lib.GPS_MATERN32


class Matern52(Stationary):
    """sigma^2 (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r) (kernels.py:601-610)."""
    _ptype = _lib.GPS_MATERN52


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

    def _emit(self, b, presliced=False):
        return b.prim(_lib.GPS_PERIODIC, self._dims(presliced), False,
                      [(lambda: self.variance, 1), (lambda: self.lengthscales, 1),
                       (lambda: self.period, 1)])


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
        for kern in kern_list:
            if isinstance(kern, Kernel):
                self._parameters = self._parameters + kern.parameters

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


class Sum(Combination):
    _op = _lib.GPS_OP_ADD


class Product(Combination):
    _op = _lib.GPS_OP_MUL
