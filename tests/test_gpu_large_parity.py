"""Parity at the sizes BASELINE.json states ("parity is green but only at toy sizes").

  * C2  GPR ARD-RBF N=8192 D=8: objective + every gradient against the committed CPU-oracle scalars
        (tests/golden/gpr_large_scalars.json, oracle/gen_large_golden.py) and predict_f against
        oracle.ref_torch evaluated on this box's CPU;
  * C3  the NKN topology (6 primitives, Linear/Product x5) at N=4096: objective + all 120 gradients
        against torch autograd through the oracle on the CPU;
  * C4  the SVGP shape M=1024, B=8192, D=16 (whitened, full q_sqrt): ELBO + all gradients (incl. Z,
        q_mu, q_sqrt) against the oracle on the CPU;
  * C5  GPR ARD-RBF N=32768 D=8 (the headline size: 256 leaves, K=16384 trailing GEMMs, 8 GiB
        buffers): objective + gradients against the committed CPU-oracle scalars.
Tolerance: 1e-8 relative (north star)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import ref_torch as R
from util import assert_close, conv, dev

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden', 'gpr_large_scalars.json')
RTOL = 1e-8


def _constrained_grads(model, grads):
    # the package differentiates w.r.t. raw; theta = softplus(raw) + 1e-6 => d theta / d raw = sigmoid(raw)
    return [(g / torch.sigmoid(p.unconstrained_tensor.detach())).reshape(-1).cpu().numpy()
            for p, g in zip(model.parameters, grads)]


@pytest.mark.parametrize('n', [8192, 32768])
def test_gpr_rbf_at_baseline_size_matches_cpu_oracle_scalars(n):
    import gpflowSlim as gpf
    gold = json.load(open(GOLD))['cases'].get(str(n))
    if gold is None:
        pytest.skip('no committed oracle scalars for N=%d' % n)
    free, _ = torch.cuda.mem_get_info()
    if free < 3.2 * 8 * n * n:
        pytest.skip('not enough free device memory for N=%d' % n)
    d = 8
    X, Y = cases.synth_gpr(n, d)
    m = gpf.models.GPR(conv(X), conv(Y), kern=gpf.kernels.RBF(d, ARD=True, lengthscales=math.sqrt(d)))
    params = [p.unconstrained_tensor for p in m.parameters]
    obj = m.objective
    g = _constrained_grads(m, torch.autograd.grad(obj, params))
    assert_close(obj.detach(), np.array(gold['nlml']), RTOL, 'objective N=%d' % n)
    assert_close(g[0], np.array([gold['g_variance']]), RTOL, 'd/d variance')
    assert_close(g[1], np.array(gold['g_lengthscales']), RTOL, 'd/d lengthscales')
    assert_close(g[2], np.array([gold['g_noise']]), RTOL, 'd/d noise')


def test_c2_predict_f_matches_oracle_on_this_cpu():
    import gpflowSlim as gpf
    n, d = 8192, 8
    X, Y = cases.synth_gpr(n, d)
    Xs = np.random.default_rng(1).standard_normal((1024, d))
    m = gpf.models.GPR(conv(X), conv(Y), kern=gpf.kernels.RBF(d, ARD=True, lengthscales=math.sqrt(d)))
    with torch.no_grad():
        mu, var = m.predict_f(conv(Xs))
        spec = dict(type='rbf', variance=torch.tensor(1.0, dtype=torch.float64),
                    lengthscales=torch.full((d,), math.sqrt(d), dtype=torch.float64))
        mo, vo = R.gpr_predict(spec, torch.tensor(X), torch.tensor(Y), torch.tensor(0.1, dtype=torch.float64),
                               torch.tensor(Xs))
    assert_close(mu, mo.numpy(), RTOL, 'C2 predictive mean')
    assert_close(var, vo.numpy(), RTOL, 'C2 predictive variance')


def test_c3_nkn_topology_at_n4096_matches_oracle_autograd():
    import gpflowSlim as gpf
    n, d = 4096, 8
    X, Y = cases.synth_gpr(n, d)
    kern = cases.nkn_c3_kernel(gpf, d)
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern)
    params = [p.unconstrained_tensor for p in m.parameters]
    obj = m.objective
    grads = torch.autograd.grad(obj, params)
    # oracle: same raw values, parameter order wrapper (W0,b0,W2,b2,W4,b4) then primitives
    # (neural_kernel_network.py:31-33), then the noise
    raw = [p.detach().cpu().clone().requires_grad_(True) for p in params]
    c = [R.softplus_fwd(r) for r in raw]
    layers = [('linear', c[0], c[1]), ('product', 2), ('linear', c[2], c[3]), ('product', 2), ('linear', c[4], c[5])]
    prims = [dict(type='rbf', variance=c[6], lengthscales=c[7]), dict(type='rbf', variance=c[8], lengthscales=c[9]),
             dict(type='periodic', variance=c[10], lengthscales=c[11], period=c[12]),
             dict(type='periodic', variance=c[13], lengthscales=c[14], period=c[15]),
             dict(type='linear', variance=c[16]), dict(type='linear', variance=c[17])]
    assert len(raw) == 19
    o = R.gpr_nlml(dict(type='nkn', prims=prims, layers=layers), torch.tensor(X), torch.tensor(Y), c[18])
    go = torch.autograd.grad(o, raw)
    assert_close(obj.detach(), o.detach().numpy(), RTOL, 'C3 objective')
    for i, (a, b) in enumerate(zip(grads, go)):
        assert_close(a, b.numpy(), RTOL, 'C3 grad %d' % i)


def test_c4_svgp_shape_matches_oracle_autograd():
    import gpflowSlim as gpf
    n, d, M, B = 100000, 16, 1024, 8192         # the rows beyond the minibatch play no role in one step
    X, Y, Z = cases.synth_svgp(n, d, M, seed=0)
    kern = gpf.kernels.RBF(d, ARD=True, lengthscales=4.0)
    m = gpf.models.SVGP(conv(X[:B]), conv(Y[:B]), kern, gpf.likelihoods.Gaussian(var=0.1), Z=Z.copy(),
                        num_data=1000000)
    rng = np.random.default_rng(5)
    with torch.no_grad():
        qm = m._q_mu.unconstrained_tensor
        qm.copy_(torch.as_tensor(0.3 * rng.standard_normal(tuple(qm.shape))).to(qm))
        qs = m._q_sqrt.unconstrained_tensor
        qs.add_(torch.as_tensor(0.05 * rng.standard_normal(tuple(qs.shape))).to(qs))
    params = m.trainable_tensors                  # kern.variance, kern.ls, noise, q_mu, q_sqrt, Z
    assert len(params) == 6
    obj = m.objective
    grads = torch.autograd.grad(obj, params)
    raw = [p.detach().cpu().clone().requires_grad_(True) for p in params]
    spec = dict(type='rbf', variance=R.softplus_fwd(raw[0]), lengthscales=R.softplus_fwd(raw[1]))
    o = R.svgp_objective(spec, torch.tensor(X[:B]), torch.tensor(Y[:B]), raw[5], raw[3], R.vec_to_tri(raw[4], M),
                         R.softplus_fwd(raw[2]), 1000000, whiten=True)
    go = torch.autograd.grad(o, raw)
    assert_close(obj.detach(), o.detach().numpy(), RTOL, 'C4 objective')
    for i, (a, b) in enumerate(zip(grads, go)):
        assert_close(a, b.numpy(), RTOL, 'C4 grad %d' % i)
