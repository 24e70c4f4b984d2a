"""Settings object with the reference's attribute surface (_settings.py:24-89, gpflowrc:6-11):
`settings.float_type`, `settings.dtypes.float_type`, `settings.numerics.jitter_level`,
`settings.jitter`, `settings.set_jitter`, `settings.temp_settings`, `settings.get_settings`.

A `gpflowrc` is searched in cwd, then $HOME, then next to this file (same order as
_settings.py:159-181).  The packaged default differs from the reference in ONE value:
float_type is float64, because the B200 path is the FP64 path (the reference's packaged
default is float32, gpflowrc:7); float32 inputs are promoted.
`settings.device` (new) is the torch device models are placed on: cuda:$LOCAL_RANK, or cpu
when no GPU is visible (host-side logic only -- compute needs the GPU).
"""
import configparser
import copy
import os
from collections import OrderedDict

import numpy as np
import torch


class _Section(OrderedDict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _parse(s):
    if s in ('true', 'True'):
        return True
    if s in ('false', 'False'):
        return False
    if s in ('float64', 'float32', 'float16', 'int64', 'int32', 'int16'):
        return getattr(np, s)
    for cast in (int, float):
        try:
            return cast(s)
        except ValueError:
            pass
    return s


def _read():
    cfg = configparser.ConfigParser()
    here = os.path.dirname(os.path.realpath(__file__))
    for loc in (os.path.abspath(os.curdir), os.path.expanduser('~'), here):
        if cfg.read(os.path.join(loc, 'gpflowrc')) or cfg.read(os.path.join(loc, '.gpflowrc')):
            break
    out = _Section()
    for sec in cfg.sections():
        out[sec] = _Section((k, _parse(v)) for k, v in cfg.items(sec))
    out.setdefault('dtypes', _Section(float_type=np.float64, int_type=np.int32))
    out.setdefault('numerics', _Section(jitter_level=1e-6, ekern_quadrature='warn'))
    return out


class _Ctx(object):
    def __init__(self, mgr, tmp):
        self.mgr, self.tmp = mgr, tmp

    def __enter__(self):
        self.mgr._stack.append(self.mgr._cur)
        self.mgr._cur = self.tmp

    def __exit__(self, *a):
        self.mgr._cur = self.mgr._stack.pop()


class _SettingsManager(object):
    def __init__(self, cur):
        object.__setattr__(self, '_cur', cur)
        object.__setattr__(self, '_stack', [])
        object.__setattr__(self, '_jitter', cur['numerics']['jitter_level'])
        object.__setattr__(self, '_device', None)

    def __getattr__(self, name):
        try:
            return self._cur[name]
        except KeyError:
            raise AttributeError('Unknown setting.')

    def temp_settings(self, tmp):
        return _Ctx(self, tmp)

    def get_settings(self):
        return copy.deepcopy(self._cur)

    def set_jitter(self, jitter):
        object.__setattr__(self, '_jitter', jitter)

    @property
    def jitter(self):
        return self._jitter

    @property
    def float_type(self):
        return self.dtypes.float_type

    @property
    def int_type(self):
        return self.dtypes.int_type

    tf_float = np_float = float_type
    tf_int = np_int = int_type

    @property
    def device(self):
        if self._device is not None:
            return self._device
        if torch.cuda.is_available():
            return torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)) % torch.cuda.device_count())
        return torch.device('cpu')

    @device.setter
    def device(self, dev):
        object.__setattr__(self, '_device', None if dev is None else torch.device(dev))


SETTINGS = _SettingsManager(_read())
