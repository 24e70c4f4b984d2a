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
