"""The reference's own two hot-path tests (tests/reference_own_tests.py) on the CPU test double."""
import numpy as np
import pytest
import torch

import cpu_ops_double
import reference_own_tests as rot


def conv(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=torch.float64)


@pytest.fixture
def gpf(monkeypatch):
    import gpflowSlim
    cpu_ops_double.install(monkeypatch)
    old = gpflowSlim.settings.device
    gpflowSlim.settings.device = 'cpu'
    yield gpflowSlim
    gpflowSlim.settings.device = None if old.type == 'cpu' else old


def test_multivariate_normal_feature_logp(gpf):
    # the Woodbury form adds settings.jitter (1e-6) to diag(L) inside its log-determinant
    # (densities.py:111), so the two agree to ~1e-6 absolute, which is the reference test's
    # assertAllClose default (rtol = atol = 1e-6)
    a, b = rot.mvn_feature_vs_cholesky(gpf, conv)
    assert abs(a - b) < 1e-6 + 1e-6 * abs(b)


def test_gpr_feature_predict_equals_standard(gpf):
    for got, want in rot.predict_feature_vs_standard(gpf, conv):
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)
