"""Small host-side helpers (reference misc.py:35-109)."""
import numpy as np
import torch

from ._settings import SETTINGS as settings


def is_tensor(value):
    return isinstance(value, torch.Tensor)


def is_ndarray(value):
    return isinstance(value, np.ndarray)


def is_number(value):
    return not isinstance(value, str) and np.isscalar(value)


def to_tensor(value, device=None):
    """numpy / python / torch value -> float64 torch tensor on the settings device (the
    reference accepts numpy inputs everywhere, e.g. examples/gpr.py:49)."""
    device = device or settings.device
    if isinstance(value, torch.Tensor):
        return value.to(device=device, dtype=torch.float64)
    return torch.as_tensor(np.asarray(value, dtype=np.float64), device=device)


def vec_to_tri(vectors, N):
    """[K, N(N+1)/2] rows -> [K, N, N] lower-triangular matrices, numpy.tril_indices (row-major)
    order (misc.py:88-109)."""
    idx = torch.tril_indices(N, N, device=vectors.device)
    out = vectors.new_zeros((vectors.shape[0], N, N))
    out[:, idx[0], idx[1]] = vectors
    return out
