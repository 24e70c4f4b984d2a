"""The package's HOST logic against the reference's golden vectors, without a GPU.

`tests/cpu_ops_double.py` replaces the CUDA-backed ops by torch-CPU stand-ins (a test double;
the product has no CPU path), so what is exercised here is everything ABOVE the C ABI: kernel
expression compilation into `gps_kernel_desc`, the composed kernels, parameter transforms,
models / conditionals / KL / likelihood glue and the optimisers.  The same case functions run
against the real CUDA kernels in tests/test_gpu_parity.py.  Tolerance: 1e-8 relative (north
star) -- observed ~1e-12."""
import numpy as np
import pytest
import torch

import cpu_ops_double
from oracle import cases

RTOL = 1e-8


def conv(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=torch.float64)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture
def gpf(monkeypatch):
    import gpflowSlim
    cpu_ops_double.install(monkeypatch)
    old = gpflowSlim.settings.device
    gpflowSlim.settings.device = 'cpu'
    yield gpflowSlim
    gpflowSlim.settings.device = None if old.type == 'cpu' else old


@pytest.mark.parametrize('name', list(cases.CASES))
def test_case_matches_reference_golden(gpf, golden, name):
    gold = golden(name)
    res = cases.run_case(gpf, name, conv)
    assert sorted(res) == sorted(gold), (sorted(set(res) ^ set(gold)))
    for key in sorted(gold):
        tol = 1e-14 if key.startswith('param/') else RTOL
        e = relerr(res[key], gold[key])
        assert e < tol, '%s/%s: relative error %.3e' % (name, key, e)
