"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz by running the UNMODIFIED reference.

    python oracle/gen_golden.py            # needs /root/reference (this container only)

The reference package (/root/reference/gpflowSlim) is imported as-is with `oracle/tf_shim`
standing in for the absent TensorFlow 1.x (see tf_shim/tensorflow/__init__.py for what that
does and does not pin) and float_type switched to float64 the way a cwd `gpflowrc` would
(gpflowrc:7, _settings.py:159-181).  Each case in oracle/cases.py is evaluated through the
reference's public API and its outputs + parameter gradients are stored.  The GPU box has no
/root/reference; it only ever sees the committed .npz files.
"""
import collections
import collections.abc
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('GPSLIM_REFERENCE', '/root/reference')

collections.Mapping = collections.abc.Mapping      # _settings.py:147 predates Python 3.10
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

import numpy as np                                  # noqa: E402
import torch                                        # noqa: E402
import tensorflow as tf                             # noqa: E402  (the shim)
import gpflowSlim as gpf                            # noqa: E402  (the reference)

assert os.path.realpath(gpf.__file__).startswith(os.path.realpath(REF)), gpf.__file__
gpf.settings.dtypes.float_type = np.float64

# TensorFlow converts numpy operands of a tensor op on the fly (`gh_x * tf.sqrt(...)`,
# likelihoods.py:147); torch does not multiply ndarray * Tensor.  The Gauss-Hermite nodes are
# therefore handed to the reference as tensors -- same values, no reference source touched.
import gpflowSlim.likelihoods as _ref_lik           # noqa: E402
import gpflowSlim.quadrature as _ref_quad           # noqa: E402

_np_hermgauss = _ref_quad.hermgauss


def _hermgauss_as_tensors(n):
    x, w = _np_hermgauss(n)
    return torch.as_tensor(x), torch.as_tensor(w)


_ref_quad.hermgauss = _hermgauss_as_tensors
_ref_lik.hermgauss = _hermgauss_as_tensors

# Same on-the-fly conversion in kernel_kitchen_sink.py:90 (`np.random.normal(...) / tf.reshape(ls)`):
# that module sees a numpy stand-in whose random draws (same global RNG stream, same values) come
# back as tensors; every other attribute is numpy's.
import gpflowSlim.kernel_kitchen_sink as _ref_ks    # noqa: E402


class _RandomAsTensors(object):
    def normal(self, *a, **kw):
        return torch.as_tensor(np.random.normal(*a, **kw))

    def uniform(self, *a, **kw):
        return torch.as_tensor(np.random.uniform(*a, **kw))


class _NumpyWithTensorDraws(object):
    random = _RandomAsTensors()

    def __getattr__(self, name):
        return getattr(np, name)


_ref_ks.np = _NumpyWithTensorDraws()

# ... and in priors.py:34-100, whose hyper-parameters are numpy arrays multiplied into tensors
# (`-shape * tf.log(scale)`, densities.py:46): they are stored as tensors of the same values.
import gpflowSlim.priors as _ref_priors             # noqa: E402


class _NumpyWithTensorHyper(object):
    @staticmethod
    def atleast_1d(a):
        return torch.as_tensor(np.atleast_1d(a))

    def __getattr__(self, name):
        return getattr(np, name)


_ref_priors.np = _NumpyWithTensorHyper()
torch.Tensor.astype = lambda self, dtype: self.to(tf._dt(dtype))

from oracle import cases                            # noqa: E402


def conv(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def main(names):
    outdir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(outdir, exist_ok=True)
    for name in names:
        tf.shim_reset()
        res = cases.run_case(gpf, name, conv)
        path = os.path.join(outdir, name + '.npz')
        np.savez_compressed(path, **res)
        print('%-22s %3d arrays  %8.1f KiB' % (name, len(res), os.path.getsize(path) / 1024.0))


if __name__ == '__main__':
    main(sys.argv[1:] or list(cases.CASES))
