"""Inducing features (reference features.py:24-87, :153-193)."""
from functools import singledispatch

import numpy as np
import torch

from . import conditionals
from ._backend import ops as _ops
from .params import Parameter, param_value


class InducingFeature(object):
    def __len__(self):
        raise NotImplementedError

    def Kuu(self, kern, jitter=0.0):
        raise NotImplementedError

    def Kuf(self, kern, Xnew):
        raise NotImplementedError

    def Kfu(self, kern, Xnew):
        """Kuf^T [N, M]: the orientation the row-major NT products want.  Features whose Kuf is
        not a plain kernel evaluation (Multiscale) go through their own Kuf."""
        return _ops.t(self.Kuf(kern, Xnew))

    @property
    def parameters(self):
        """Parameters of the feature itself (the reference trains tf.trainable_variables())."""
        return []


class InducingPoints(InducingFeature):
    """Real-space inducing points; Z is a trainable Parameter (features.py:55-81)."""

    def __init__(self, Z):
        super().__init__()
        self._Z = Parameter(Z, name='Z')

    Z = param_value('Z')

    def __len__(self):
        return self.Z.shape[0]

    def Kuu(self, kern, jitter=0.0):
        """K(Z) + jitter I -- the jitter lands in the Gram kernel's diagonal epilogue
        (features.py:74-77)."""
        return kern.K_jittered(self.Z, jitter)

    def Kuf(self, kern, Xnew):
        return kern.K(self.Z, Xnew)

    def Kfu(self, kern, Xnew):
        return kern.K(Xnew, self.Z)

    @property
    def parameters(self):
        return [self._Z]


class Multiscale(InducingPoints):
    """Multi-scale inducing features (Lazaro-Gredilla & Figueiras-Vidal 2009; features.py:89-150):
    each inducing point carries its own Gaussian widths `scales` [M, D]; defined for the RBF
    kernel only.  O(N M D) elementwise device work (the per-pair lengthscales rule out the
    inner-product form of the fused Gram kernel)."""

    def __init__(self, Z, scales):
        super().__init__(Z)
        from . import transforms
        self._scales = Parameter(scales, transform=transforms.positive)
        if tuple(self.Z.shape) != tuple(np.shape(scales)):
            raise ValueError('Input locations `Z` and `scales` must have the same shape.')

    scales = param_value('scales')

    @property
    def parameters(self):
        return [self._Z, self._scales]

    def Kfu(self, kern, Xnew):
        return _ops.t(self.Kuf(kern, Xnew))

    def _cust_square_dist(self, A, B, sc):
        """sum_d ((a_d - b_d) / sc_d)^2 with per-pair scales sc [N, M, D] (or broadcastable)."""
        return (((A.unsqueeze(1) - B.unsqueeze(0)) / sc) ** 2).sum(2)

    @staticmethod
    def _check(kern):
        from . import kernels
        if not isinstance(kern, kernels.RBF):
            raise NotImplementedError('Multiscale features not implemented for `%s`.' % str(type(kern)))

    def Kuf(self, kern, Xnew):
        self._check(kern)
        from .misc import to_tensor
        Xnew, _ = kern._slice(to_tensor(Xnew), None)
        Zmu, Zlen = kern._slice(self.Z, self.scales)
        idlengthscales = kern.lengthscales + Zlen
        d = self._cust_square_dist(Xnew, Zmu, idlengthscales)
        return (kern.variance * torch.exp(-d / 2) *
                (kern.lengthscales / idlengthscales).prod(1).reshape(1, -1)).t()

    def Kuu(self, kern, jitter=0.0):
        self._check(kern)
        Zmu, Zlen = kern._slice(self.Z, self.scales)
        idlengthscales2 = (kern.lengthscales + Zlen) ** 2
        sc = torch.sqrt(idlengthscales2.unsqueeze(0) + idlengthscales2.unsqueeze(1) - kern.lengthscales ** 2)
        d = self._cust_square_dist(Zmu, Zmu, sc)
        Kzz = kern.variance * torch.exp(-d / 2) * (kern.lengthscales / sc).prod(2)
        return Kzz + jitter * torch.eye(len(self), dtype=Kzz.dtype, device=Kzz.device)


@singledispatch
def conditional(feat, kern, Xnew, f, *, full_cov=False, q_sqrt=None, white=False):
    raise NotImplementedError('No implementation for {} found'.format(type(feat).__name__))


@conditional.register(InducingPoints)
def default_feature_conditional(feat, kern, Xnew, f, *, full_cov=False, q_sqrt=None, white=False):
    return conditionals.feature_conditional(Xnew, feat, kern, f, full_cov=full_cov, q_sqrt=q_sqrt,
                                            white=white)


def inducingpoint_wrapper(feat, Z):
    """features.py:177-193."""
    if feat is not None and Z is not None:
        raise ValueError('Cannot pass both an InducingFeature instance and Z values')
    elif feat is None and Z is None:
        raise ValueError('You must pass either an InducingFeature instance or Z values')
    elif Z is not None:
        feat = InducingPoints(Z)
    elif isinstance(feat, (np.ndarray, torch.Tensor)):
        feat = InducingPoints(feat)
    else:
        assert isinstance(feat, InducingFeature)
    return feat
