"""Parameter (reference params.py:131-194): an UNCONSTRAINED torch leaf `vf_val` =
transform.backward(value) (params.py:142-145) whose `.value` is
transform.forward_tensor(vf_val) (params.py:165-166).  Optimisers train `vf_val`
(== `unconstrained_tensor`)."""
import numpy as np
import torch

from ._settings import SETTINGS as settings
from .transforms import Identity


class Parameter(object):
    def __init__(self, value, transform=None, prior=None, trainable=True, dtype=None, name='Param'):
        self.instance_name = name
        self.prior = prior
        self.transform = transform if transform is not None else Identity()
        self.trainable = trainable
        if isinstance(value, torch.Tensor):
            value = value.detach().cpu().numpy()
        vf_value = self.transform.backward(np.asarray(value, dtype=np.float64))
        self.vf_val = torch.tensor(np.asarray(vf_value, dtype=np.float64), dtype=torch.float64,
                                   device=settings.device, requires_grad=bool(trainable))

    @property
    def name(self):
        return self.instance_name

    @property
    def shape(self):
        return tuple(self.vf_val.shape)

    @property
    def dtype(self):
        return self.vf_val.dtype

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def value(self):
        return self.transform.forward_tensor(self.vf_val)

    @property
    def unconstrained_tensor(self):
        return self.vf_val

    @property
    def constrained_tensor(self):
        return self.value

    def read_value(self):
        return self.value.detach().cpu().numpy()

    def assign(self, value):
        """Set the CONSTRAINED value in place."""
        if isinstance(value, torch.Tensor):
            value = value.detach().cpu().numpy()
        new = self.transform.backward(np.asarray(value, dtype=np.float64))
        with torch.no_grad():
            self.vf_val.copy_(torch.as_tensor(np.asarray(new), dtype=torch.float64).reshape(self.vf_val.shape))

    def to(self, device):
        req = self.vf_val.requires_grad
        self.vf_val = self.vf_val.detach().to(device).requires_grad_(req)
        return self

    def _build_prior(self, unconstrained_tensor, constrained_tensor):
        """log prior density incl. the log Jacobian (params.py:176-194); 0 without a prior."""
        if self.prior is None:
            return torch.zeros((), dtype=torch.float64, device=unconstrained_tensor.device)
        log_jacobian = self.transform.log_jacobian_tensor(unconstrained_tensor)
        logp_var = self.prior.logp(constrained_tensor)
        return (logp_var + log_jacobian).squeeze()

    def __repr__(self):
        return 'Parameter(%s, shape=%s, transform=%s)' % (self.name, self.shape, self.transform)


class param_value(object):
    """Class-level descriptor: `variance = param_value('variance')` exposes the CONSTRAINED value
    of the Parameter stored as `self._variance` (what the reference spells out as a four-line
    @property per parameter)."""

    def __init__(self, name):
        self.attr = '_' + name

    def __get__(self, obj, owner=None):
        return self if obj is None else getattr(obj, self.attr).value


class Parameterized(object):
    """Mixin: `_param(name, value, transform)` creates the Parameter, stores it as `_<name>` and
    appends it to `_parameters` (the list models collect, models/model.py:119)."""

    def _param(self, name, value, transform=None):
        prm = Parameter(value, transform=transform, name=name)
        setattr(self, '_' + name, prm)
        self._parameters = list(getattr(self, '_parameters', [])) + [prm]
        return prm

