"""Inducing features (reference features.py:24-87, :153-193)."""
from functools import singledispatch

import numpy as np
import torch

from . import conditionals
from ._backend import ops as _ops
from .params import Parameter


class InducingFeature(object):
    def __len__(self):
        raise NotImplementedError

    def Kuu(self, kern, jitter=0.0):
        raise NotImplementedError

    def Kuf(self, kern, Xnew):
        raise NotImplementedError


class InducingPoints(InducingFeature):
    """Real-space inducing points; Z is a trainable Parameter (features.py:55-81)."""

    def __init__(self, Z):
        super().__init__()
        self._Z = Parameter(Z, name='Z')

    @property
    def Z(self):
        return self._Z.value

    def __len__(self):
        return self.Z.shape[0]

    def Kuu(self, kern, jitter=0.0):
        """K(Z) + jitter I -- the jitter lands in the Gram kernel's diagonal epilogue
        (features.py:74-77)."""
        return _ops.gram(kern.program(), self.Z, None, diag_add=float(jitter))

    def Kuf(self, kern, Xnew):
        return kern.K(self.Z, Xnew)


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
