"""GPU parity tests proper: the B200 package is run through the SAME case functions
(oracle/cases.py) that produced tests/golden/*.npz from the unmodified reference, and every
output -- objectives, gradients w.r.t. every unconstrained parameter and the inducing inputs,
predictive means / variances, Gram matrices, KL terms -- must agree within the north-star
tolerance of 1e-8 relative (relative to the largest magnitude of the reference array)."""
import numpy as np
import pytest
import torch

from oracle import cases
from util import assert_close, conv, relerr

pytestmark = pytest.mark.gpu

RTOL = 1e-8   # BASELINE.json north_star: "within 1e-8 relative"


def _gpf():
    import gpflowSlim as gpf
    return gpf


@pytest.mark.parametrize('name', [c for c in cases.CASES if c not in cases.LATE_CASES + cases.HOST_ONLY_CASES])
def test_case_matches_reference_golden(golden, name):
    gold = golden(name)
    res = cases.run_case(_gpf(), name, conv)
    assert set(res) == set(gold), sorted(set(res) ^ set(gold))
    worst = 0.0
    for key in sorted(gold):
        if key.startswith('param/'):
            assert_close(res[key], gold[key], 1e-12, name + ':' + key)   # same initial state
            continue
        e = relerr(res[key], gold[key])
        worst = max(worst, e)
        assert e < RTOL, '%s:%s relative error %.3e' % (name, key, e)
    print('%s: worst relative error %.2e over %d arrays' % (name, worst, len(gold)))


def test_fused_gpr_equals_opwise_and_predict_nograd(golden):
    """The fused one-call GPR objective/predict agree with the op-by-op autograd path."""
    gpf = _gpf()
    gold = golden('gpr_c1_ls')
    X, Y = cases.synth_gpr(1000, 4, seed=0)
    Xs = np.random.default_rng(1).standard_normal((64, 4))
    outs = {}
    for fused in (True, False):
        kern = gpf.kernels.RBF(4, ARD=True, lengthscales=2.0)
        m = gpf.models.GPR(conv(X), conv(Y), kern=kern, fused=fused)
        obj = m.objective
        gr = torch.autograd.grad(obj, [p.unconstrained_tensor for p in m.parameters])
        with torch.no_grad():
            mu, var = m.predict_f(conv(Xs))
            mu2, cov = m.predict_f_full_cov(conv(Xs))
        outs[fused] = (obj, gr, mu, var, cov)
        assert_close(obj, gold['objective'], RTOL, 'objective fused=%s' % fused)
        for i, g in enumerate(gr):
            assert_close(g, gold['grad/objective/%d' % i], RTOL, 'grad %d fused=%s' % (i, fused))
        assert_close(mu, gold['pred_mu'], RTOL, 'pred_mu fused=%s' % fused)
        assert_close(var, gold['pred_var'], RTOL, 'pred_var fused=%s' % fused)
        assert_close(torch.diagonal(cov[:, :, 0]), gold['pred_var'][:, 0], RTOL, 'diag full cov')
    assert_close(outs[True][0], outs[False][0], 1e-11, 'fused vs opwise objective')


def test_tf_adam_step_parity(golden):
    """C1: NLML -> one TF-semantics Adam(1e-3) step -> NLML again, against the oracle."""
    from oracle import ref_torch as R
    gpf = _gpf()
    X, Y = cases.synth_gpr(1000, 4, seed=0)
    kern = gpf.kernels.RBF(4, ARD=True)
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern)
    opt = gpf.training.AdamOptimizer(1e-3)
    o0 = opt.minimize(m)
    o1 = m.objective.detach()
    # oracle
    raw = [torch.tensor(R.softplus_inv(v), dtype=torch.float64, requires_grad=True)
           for v in (1.0, np.ones(4), 0.1)]
    Xc, Yc = torch.tensor(X), torch.tensor(Y)

    def obj(raw):
        spec = dict(type='rbf', variance=R.softplus_fwd(raw[0]), lengthscales=R.softplus_fwd(raw[1]))
        return R.gpr_nlml(spec, Xc, Yc, R.softplus_fwd(raw[2]))
    r0 = obj(raw)
    gr = torch.autograd.grad(r0, raw)
    new = [p.detach().requires_grad_(True) for p in R.tf_adam_step([r.detach() for r in raw], gr, {})]
    r1 = obj(new)
    assert_close(o0, r0, RTOL, 'objective before step')
    assert_close(o1, r1, RTOL, 'objective after step')
    for p, q in zip(m.parameters, new):
        assert_close(p.unconstrained_tensor, q.reshape(p.unconstrained_tensor.shape), 1e-10, 'params after step')


def test_not_positive_definite_raises():
    gpf = _gpf()
    from gpflowSlim._backend import ops
    A = conv(np.array([[1.0, 2.0], [2.0, 1.0]]))
    with pytest.raises(gpf.CholeskyError):
        ops.potrf(A)
    big = np.eye(300)
    big[200, 200] = -1.0
    with pytest.raises(gpf.CholeskyError) as ei:
        ops.potrf(conv(big))
    assert '201' in str(ei.value)
