"""GPU parity of the components added after the last session with GPU time (SURVEY.md section
8f rank 4: composed kernels, L-BFGS).  Same contract as tests/test_gpu_parity.py -- the case
functions of oracle/cases.py against the golden vectors of the unmodified reference, 1e-8
relative -- kept in a file that sorts last so these newer cases can never mask the established
ones under `pytest -x`."""
import numpy as np
import pytest
import torch

from oracle import cases
from util import assert_close, conv, relerr

pytestmark = pytest.mark.gpu

RTOL = 1e-8   # BASELINE.json north_star: "within 1e-8 relative"


@pytest.mark.parametrize('name', list(cases.LATE_CASES))
def test_late_case_matches_reference_golden(golden, name):
    import gpflowSlim as gpf
    gold = golden(name)
    res = cases.run_case(gpf, name, conv)
    assert set(res) == set(gold), sorted(set(res) ^ set(gold))
    worst = 0.0
    for key in sorted(gold):
        if key.startswith('param/'):
            assert_close(res[key], gold[key], 1e-12, name + ':' + key)
            continue
        e = relerr(res[key], gold[key])
        worst = max(worst, e)
        assert e < RTOL, '%s:%s relative error %.3e' % (name, key, e)
    print('%s: worst relative error %.2e over %d arrays' % (name, worst, len(gold)))


def test_composed_kernels_run_on_the_library():
    """The inner products of the composed kernels go through libgpslim_b200.so (DMMA GEMM /
    fused Linear Gram), not through torch.matmul: the handle's launch counter must move."""
    import gpflowSlim as gpf
    from gpflowSlim._backend.lib import handle_for
    rng = np.random.default_rng(0)
    X = conv(rng.standard_normal((300, 5)))
    h = handle_for(X)
    for kern in (gpf.kernels.RatQuad(5, ARD=True), gpf.kernels.ArcCosine(5, order=1),
                 gpf.kernels.Polynomial(5), gpf.kernels.TPS(5)):
        before = h.profile_read(reset=False)[2]
        K = kern.K(X)
        assert h.profile_read(reset=False)[2] > before, type(kern).__name__
        assert K.shape == (300, 300) and bool(torch.isfinite(K).all())
        assert_close(K, K.t(), 1e-12, type(kern).__name__ + ' symmetry')


def test_reference_own_tests_on_the_gpu():
    """densities.py:159-174 and models/gpr.py:135-203 of the reference (see
    tests/reference_own_tests.py)."""
    import gpflowSlim as gpf
    import reference_own_tests as rot
    a, b = rot.mvn_feature_vs_cholesky(gpf, conv)
    assert abs(a - b) < 1e-6 + 1e-6 * abs(b)
    for got, want in rot.predict_feature_vs_standard(gpf, conv):
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)


def _load_example(name):
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('example_' + name, os.path.join(root, 'examples', name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_example_scripts_run_on_the_gpu():
    """examples/gpr.py (Adam steps, then the L-BFGS trainer) and examples/svgp.py (network +
    multi-class SVGP trained jointly; the gradient reaches the network through the Gram
    kernels' d/dX path) on the real kernels."""
    ex = _load_example('gpr')
    first, _, _ = ex.main(iters=1, quiet=True)
    last, _, _ = ex.main(iters=25, quiet=True)
    assert last < first
    obj, rmse, ll = ex.main(lbfgs=True, quiet=True)
    assert obj < last and rmse < 6.0 and ll > -3.5
    sv = _load_example('svgp')
    loss, acc, ll = sv.main(iters=12, quiet=True, n_train=600, n_test=200, num_inducing=20,
                            minibatch_size=100, num_h=8)
    assert loss == loss and 0.0 <= acc <= 1.0 and ll == ll

