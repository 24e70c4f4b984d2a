"""Pins oracle/ref_torch.py to the golden vectors the UNMODIFIED reference produced
(oracle/gen_golden.py).  CPU only.  Tolerance 1e-10 relative: both sides are float64 LAPACK
arithmetic of the same op sequence."""
import math

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import ref_torch as R

RTOL = 1e-10


def close(a, b, rtol=RTOL, what=''):
    a = np.asarray(a.detach().numpy() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err < rtol, '%s: rel err %.3e' % (what, err)


def sp(raw):
    return R.softplus_fwd(torch.tensor(raw, requires_grad=False))


def leaf(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=True)


def grads_wrt(obj, leaves):
    return torch.autograd.grad(obj, leaves, allow_unused=True)


def test_transforms_roundtrip():
    y = np.array([1e-5, 0.1, 1.0, 30.0, 800.0])
    raw = R.softplus_inv(y)
    close(R.softplus_fwd(raw), y, 1e-12, 'softplus roundtrip')
    v = np.arange(1.0, 13.0).reshape(2, 6)
    tri = R.vec_to_tri(v, 3)
    assert tri.shape == (3, 3, 2)
    assert tri[1, 0, 0] == 2.0 and tri[2, 2, 1] == 12.0 and tri[0, 1, 0] == 0.0


def _zoo_specs(d):
    ls = torch.tensor(0.7 + 0.15 * np.arange(d))
    t = lambda v: torch.tensor(v, dtype=torch.float64)
    st = lambda typ, v, l, ad=None, dim=d: dict(type=typ, variance=t(v), lengthscales=l,
                                                active_dims=ad, input_dim=dim)
    return {
        'rbf_iso': st('rbf', 1.3, t(0.9)),
        'rbf_ard': st('rbf', 0.8, ls),
        'm12_ard': st('matern12', 1.1, ls),
        'm32_ard': st('matern32', 0.6, ls),
        'm52_iso': st('matern52', 1.7, t(1.4)),
        'exp_ard': st('exponential', 0.9, ls),
        'lin_iso': dict(type='linear', variance=t(0.4), input_dim=d),
        'lin_ard': dict(type='linear', variance=t(0.3 + 0.1 * np.arange(d)), input_dim=d),
        'periodic': dict(type='periodic', variance=t(1.2), lengthscales=t(0.8), period=t(1.7),
                         input_dim=d),
        'rbf_active': st('rbf', 1.0, t([0.5, 1.5]), [2, 0], 2),
        'sum': dict(type='sum', children=[st('rbf', 1.0, ls),
                                          dict(type='linear', variance=t(0.2), input_dim=d), 0.37]),
        'product': dict(type='product', children=[
            st('matern32', 1.0, t(1.2)),
            dict(type='periodic', variance=t(1.0), lengthscales=t(1.0), period=t(2.1), input_dim=d),
            1.9]),
        'sum_of_product': dict(type='sum', children=[
            dict(type='product', children=[st('rbf', 1.0, t(1.0), [0, 1], 2),
                                           dict(type='linear', variance=t(1.0), active_dims=[2])]),
            st('matern52', 1.0, t(1.0))]),
    }


def test_kernels(golden):
    g = golden('kernels')
    d = 3
    rng = np.random.default_rng(10)
    X = torch.tensor(rng.standard_normal((37, d)) * 1.3)
    X2 = torch.tensor(rng.standard_normal((23, d)) * 1.3)
    for name, spec in _zoo_specs(d).items():
        # constrained values went through softplus_inv -> softplus in the reference: 1e-12 ok
        close(R.K(spec, X), g[name + '/K'], 1e-9, name + '/K')
        close(R.K(spec, X, X2), g[name + '/K2'], 1e-9, name + '/K2')
        close(R.Kdiag(spec, X), g[name + '/Kdiag'], 1e-9, name + '/Kdiag')


def _zoo_extra_specs(d):
    """Oracle specs of oracle/cases.py:_kernel_zoo_extra with a leaf per reference parameter, in
    the reference's parameter order; (spec, [(leaf, positive?)...])."""
    ls = 0.7 + 0.15 * np.arange(d)
    np.random.seed(5)
    cw = np.random.normal(size=[d, 1])
    out = {}

    def add(name, build, vals):
        leaves = [torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for v, _ in vals]
        out[name] = (build(*leaves), leaves, [pos for _, pos in vals])
    P, F = True, False
    add('white', lambda v: dict(type='white', variance=v), [(0.7, P)])
    add('constant', lambda v: dict(type='constant', variance=v), [(1.9, P)])
    add('bias', lambda v: dict(type='constant', variance=v), [(0.3, P)])
    rq = lambda v, l, a: dict(type='ratquad', variance=v, lengthscales=l, alpha=a, input_dim=d)
    add('ratquad_iso', rq, [(1.2, P), (0.9, P), (1.7, P)])
    add('ratquad_ard', rq, [(0.8, P), (ls, P), (0.6, P)])
    add('poly3', lambda v, o: dict(type='polynomial', variance=v, offset=o, degree=3.0, input_dim=d),
        [(0.4, P), (0.8, P)])
    add('poly2_ard', lambda v, o: dict(type='polynomial', variance=v, offset=o, degree=2.0, input_dim=d),
        [(0.3 + 0.1 * np.arange(d), P), (1.3, P)])
    cs = lambda v, l, w: dict(type='cosine', variance=v, lengthscales=l, weights=w, input_dim=d)
    add('cosine_iso', cs, [(1.1, P), (1.3, P), (cw, F)])
    add('cosine_ard', cs, [(0.9, P), (ls, P), (cw, F)])
    ac = lambda order: (lambda v, b, w: dict(type='arccosine', order=order, variance=v, bias_variance=b,
                                             weight_variances=w, input_dim=d))
    add('arccos0', ac(0), [(1.2, P), (0.4, P), (0.7, P)])
    add('arccos1_ard', ac(1), [(0.8, P), (1.1, P), (0.5 + 0.2 * np.arange(d), P)])
    add('arccos2', ac(2), [(0.6, P), (0.9, P), (1.3, P)])
    add('tps', lambda v, l: dict(type='tps', variance=v, input_dim=d), [(0.5, P), (1.0, P)])
    add('ratquad_active', lambda v, l, a: dict(type='ratquad', variance=v, lengthscales=l, alpha=a,
                                               active_dims=[2, 0]),
        [(1.0, P), ([0.5, 1.5], P), (2.0, P)])
    st = lambda typ, v, l, ad=None, dim=d: dict(type=typ, variance=v, lengthscales=l, active_dims=ad,
                                                input_dim=dim)
    add('mixed_sum', lambda v1, l1, v2, l2, a2, v3, v4: dict(type='sum', children=[
        st('rbf', v1, l1), rq(v2, l2, a2), dict(type='white', variance=v3),
        dict(type='linear', variance=v4, input_dim=d), 0.11]),
        [(1.0, P), (ls, P), (1.0, P), (1.0, P), (1.2, P), (0.05, P), (0.2, P)])
    add('mixed_product', lambda v1, l1, v2, o2, v3: dict(type='product', children=[
        st('matern32', v1, l1), dict(type='polynomial', variance=v2, offset=o2, degree=2.0, input_dim=d),
        dict(type='constant', variance=v3)]),
        [(1.0, P), (1.2, P), (1.0, P), (0.6, P), (1.4, P)])
    add('sum_of_mixed_product', lambda v1, b1, w1, v2, l2, v3, l3, v4: dict(type='sum', children=[
        dict(type='product', children=[
            dict(type='arccosine', order=1, variance=v1, bias_variance=b1, weight_variances=w1,
                 active_dims=[0, 1]),
            st('rbf', v2, l2, [2], 1)]),
        st('matern52', v3, l3), dict(type='constant', variance=v4)]),
        [(1.0, P), (1.0, P), (1.0, P), (1.0, P), (1.0, P), (1.0, P), (1.0, P), (0.2, P)])
    return out


def test_kernels_extra(golden):
    """SURVEY section 8(f) rank 4 covariances: Grams and parameter gradients of the oracle
    restatement against the unmodified reference."""
    g = golden('kernels_extra')
    d = 3
    rng = np.random.default_rng(40)
    X = torch.tensor(rng.standard_normal((33, d)) * 1.2)
    X2 = torch.tensor(rng.standard_normal((21, d)) * 1.2)
    W, W2 = rng.standard_normal((33, 33)), rng.standard_normal((33, 21))
    W = torch.tensor(W + W.T)
    W2 = torch.tensor(W2)
    specs = _zoo_extra_specs(d)
    in_golden = {k.split('/')[0] for k in g} - {'coregion', 'kdim', 'dimwise'}
    assert set(specs) == in_golden, sorted(set(specs) ^ in_golden)
    for name, (spec, leaves, positive) in specs.items():
        K, K2, Kd = R.K(spec, X), R.K(spec, X, X2), R.Kdiag(spec, X)
        close(K, g[name + '/K'], 1e-9, name + '/K')
        close(K2, g[name + '/K2'], 1e-9, name + '/K2')
        close(Kd, g[name + '/Kdiag'], 1e-9, name + '/Kdiag')
        val = (K2 * W2).sum() + (Kd * W[:, 0]).sum()
        if name not in ('tps', 'arccos0'):
            val = val + (K * W).sum()
        gs = torch.autograd.grad(val, leaves, allow_unused=True)
        for i, (leaf_, gr, pos) in enumerate(zip(leaves, gs, positive)):
            gr = torch.zeros_like(leaf_) if gr is None else gr
            if pos:   # d softplus(raw) / d raw = sigmoid(raw) = 1 - exp(-(y - 1e-6))
                gr = gr * (1.0 - torch.exp(-(leaf_.detach() - 1e-6)))
            close(gr.reshape(g['%s/grad%d' % (name, i)].shape), g['%s/grad%d' % (name, i)], 1e-8,
                  '%s/grad%d' % (name, i))


def _gpr_from_golden(g, d):
    raw = [leaf(g['param/objective/%d' % i]) for i in range(3)]
    spec = dict(type='rbf', variance=R.softplus_fwd(raw[0]), lengthscales=R.softplus_fwd(raw[1]))
    return raw, spec, R.softplus_fwd(raw[2])


@pytest.mark.parametrize('name,n,d,nstar', [('gpr_c1', 1000, 4, 1024), ('gpr_c1_ls', 1000, 4, 64),
                                            ('gpr_c2_small', 2048, 8, 64)])
def test_gpr(golden, name, n, d, nstar):
    g = golden(name)
    X, Y = cases.synth_gpr(n, d, seed=0)
    Xs = np.random.default_rng(1).standard_normal((nstar, d))
    X, Y, Xs = map(torch.tensor, (X, Y, Xs))
    raw, spec, noise = _gpr_from_golden(g, d)
    obj = R.gpr_nlml(spec, X, Y, noise)
    close(obj, g['objective'], RTOL, 'objective')
    for i, gr in enumerate(grads_wrt(obj, raw)):
        close(gr, g['grad/objective/%d' % i], 1e-9, 'grad %d' % i)
    mu, var = R.gpr_predict(spec, X, Y, noise, Xs)
    close(mu, g['pred_mu'], 1e-9, 'pred_mu')
    close(var, g['pred_var'], 1e-9, 'pred_var')
    # independent LAPACK + analytic-gradient route (the algebra the CUDA backward uses)
    nl, gv, gl, gn = R.gpr_nlml_grad_lapack(X.numpy(), Y.numpy(), spec['variance'].item(),
                                            spec['lengthscales'].detach().numpy(), noise.item())
    close(nl, g['objective'], RTOL, 'lapack nlml')
    sig = lambda r: torch.sigmoid(r).detach().numpy()       # d softplus / d raw
    close(gv * sig(raw[0]), g['grad/objective/0'], 1e-8, 'lapack g_var')
    close(gl * sig(raw[1]), g['grad/objective/1'], 1e-8, 'lapack g_ls')
    close(gn * sig(raw[2]), g['grad/objective/2'], 1e-8, 'lapack g_noise')


def test_nkn(golden):
    g = golden('nkn')
    d, n = 3, 150
    X, Y = cases.synth_gpr(n, d, seed=3)
    Xs = np.random.default_rng(4).standard_normal((20, d))
    X, Y, Xs = map(torch.tensor, (X, Y, Xs))
    raw = [leaf(g['param/objective/%d' % i]) for i in range(19)]
    c = [R.softplus_fwd(r) for r in raw]
    # parameter order: wrapper (W0,b0,W2,b2,W4,b4) then primitives (neural_kernel_network.py:31-33)
    layers = [('linear', c[0], c[1]), ('product', 2), ('linear', c[2], c[3]), ('product', 2),
              ('linear', c[4], c[5])]
    prims = [dict(type='rbf', variance=c[6], lengthscales=c[7]),
             dict(type='rbf', variance=c[8], lengthscales=c[9]),
             dict(type='periodic', variance=c[10], lengthscales=c[11], period=c[12]),
             dict(type='periodic', variance=c[13], lengthscales=c[14], period=c[15]),
             dict(type='linear', variance=c[16]), dict(type='linear', variance=c[17])]
    # NeuralKernelNetwork itself has no active_dims restriction on primitives
    spec = dict(type='nkn', prims=prims, layers=layers)
    noise = c[18]
    assert len([k for k in g if k.startswith('param/objective/')]) == 19
    close(R.K(spec, X[:30]), g['K'], RTOL, 'K')
    close(R.K(spec, X[:30], Xs), g['K2'], RTOL, 'K2')
    close(R.Kdiag(spec, X[:30]), g['Kdiag'], RTOL, 'Kdiag')
    obj = R.gpr_nlml(spec, X, Y, noise)
    close(obj, g['objective'], RTOL, 'objective')
    for i, gr in enumerate(grads_wrt(obj, raw[:19])):
        close(gr, g['grad/objective/%d' % i], 1e-8, 'grad %d' % i)
    mu, var = R.gpr_predict(spec, X, Y, noise, Xs)
    close(mu, g['pred_mu'], 1e-9, 'pred_mu')
    close(var, g['pred_var'], 1e-8, 'pred_var')


def _svgp_inputs(n, d, m, batch, latents):
    X, Y, Z = cases.synth_svgp(n, d, m, seed=0)
    if latents > 1:
        Y = np.concatenate([Y * (1 + 0.3 * j) + 0.1 * j for j in range(latents)], 1)
    return torch.tensor(X[:batch]), torch.tensor(Y[:batch])


@pytest.mark.parametrize('name,n,d,m,batch,whiten,q_diag,latents', [
    ('svgp_white_full', 4000, 16, 256, 1024, True, False, 1),
    ('svgp_nonwhite_full', 500, 4, 40, 200, False, False, 2),
    ('svgp_white_diag', 500, 4, 40, 200, True, True, 2),
    ('svgp_nonwhite_diag', 500, 4, 40, 200, False, True, 1)])
def test_svgp(golden, name, n, d, m, batch, whiten, q_diag, latents):
    g = golden(name)
    Xb, Yb = _svgp_inputs(n, d, m, batch, latents)
    # parameters: kern.variance, kern.ls, likelihood.variance, q_mu, q_sqrt, then Z
    raw = [leaf(g['param/objective/%d' % i]) for i in range(6)]
    spec = dict(type='rbf', variance=R.softplus_fwd(raw[0]), lengthscales=R.softplus_fwd(raw[1]))
    noise = R.softplus_fwd(raw[2])
    q_mu = raw[3]
    q_sqrt = R.softplus_fwd(raw[4]) if q_diag else R.vec_to_tri(raw[4], m)
    Z = raw[5]
    obj = R.svgp_objective(spec, Xb, Yb, Z, q_mu, q_sqrt, noise, n, whiten=whiten)
    close(obj, g['objective'], RTOL, 'objective')
    Kuu = None if whiten else R.K(spec, Z) + torch.eye(m, dtype=torch.float64) * 1e-6
    close(R.gauss_kl(q_mu, q_sqrt, Kuu), g['KL'], 1e-9, 'KL')
    for i, gr in enumerate(grads_wrt(obj, raw)):
        close(gr, g['grad/objective/%d' % i], 1e-8, 'grad %d' % i)
    Xs = None
    rng = np.random.default_rng(5)
    rng.standard_normal(tuple(g['param/objective/3'].shape))
    rng.standard_normal(tuple(g['param/objective/4'].shape))
    Xs = torch.tensor(rng.standard_normal((19, d)))
    mu, var = R.conditional(spec, Xs, Z, q_mu, q_sqrt=q_sqrt, white=whiten)
    close(mu, g['pred_mu'], 1e-9, 'pred_mu')
    close(var, g['pred_var'], 1e-8, 'pred_var')


def test_sgpr(golden):
    g = golden('sgpr')
    n, d, mi = 600, 4, 50
    X, Y, _ = cases.synth_svgp(n, d, mi, seed=8)
    Xs = torch.tensor(np.random.default_rng(9).standard_normal((21, d)))
    X, Y = torch.tensor(X), torch.tensor(Y)
    raw = [leaf(g['param/objective/%d' % i]) for i in range(4)]
    spec = dict(type='rbf', variance=R.softplus_fwd(raw[0]), lengthscales=R.softplus_fwd(raw[1]))
    noise = R.softplus_fwd(raw[2])
    Z = raw[3]
    obj = R.sgpr_objective(spec, X, Y, Z, noise)
    close(obj, g['objective'], RTOL, 'objective')
    for i, gr in enumerate(grads_wrt(obj, raw)):
        close(gr, g['grad/objective/%d' % i], 1e-8, 'grad %d' % i)
    mu, var = R.sgpr_predict(spec, X, Y, Z, noise, Xs)
    close(mu, g['pred_mu'], 1e-9, 'pred_mu')
    close(var, g['pred_var'], 1e-8, 'pred_var')
    mu, cov = R.sgpr_predict(spec, X, Y, Z, noise, Xs, full_cov=True)
    close(cov, g['full_cov'], 1e-8, 'full_cov')


def test_sparse_bounds(golden):
    """SGPRUpperMixin.compute_upper_bound on SGPR and GPRFITC, and the FITC
    likelihood, gradients and predictions (models/sgpr.py:30-82, 192-317)."""
    g = golden('sparse_bounds')
    n, d, mi = 500, 3, 40
    X, Y, Z0 = cases.synth_svgp(n, d, mi, seed=12)
    rng = np.random.default_rng(13)
    Y = np.concatenate([Y, rng.standard_normal((n, 1)) * 0.3], 1)
    Xs = torch.tensor(rng.standard_normal((17, d)))
    X, Y = torch.tensor(X), torch.tensor(Y)
    # SGPR (Matern32 ARD l=1.5 var=1.2, obs_var 0.2) at its initial state
    ls = torch.full((d,), 1.5, dtype=torch.float64)
    spec_s = dict(type='matern32', variance=torch.tensor(1.2, dtype=torch.float64), lengthscales=ls)
    noise0 = torch.tensor(0.2, dtype=torch.float64)
    Zt = torch.tensor(Z0)
    up = R.sgpr_upper_bound(spec_s, X, Y, Zt, noise0)
    lo = -R.sgpr_objective(spec_s, X, Y, Zt, noise0)
    close(up, g['sgpr_upper'], RTOL, 'sgpr upper')
    close(lo, g['sgpr_lower'], RTOL, 'sgpr lower')
    assert float(lo) <= float(up)                      # lower bound <= log p(Y) <= upper bound
    # FITC from the stored unconstrained parameters
    raw = [leaf(g['param/fitc_objective/%d' % i]) for i in range(4)]
    spec = dict(type='rbf', variance=R.softplus_fwd(raw[0]), lengthscales=R.softplus_fwd(raw[1]))
    noise = R.softplus_fwd(raw[2])
    Z = raw[3]
    obj = R.gprfitc_objective(spec, X, Y, Z, noise)
    close(obj, g['fitc_objective'], RTOL, 'fitc objective')
    for i, gr in enumerate(grads_wrt(obj, raw)):
        close(gr, g['grad/fitc_objective/%d' % i], 1e-8, 'fitc grad %d' % i)
    close(R.sgpr_upper_bound(spec, X, Y, Z, noise), g['fitc_upper'], RTOL, 'fitc upper')
    mu, var = R.gprfitc_predict(spec, X, Y, Z, noise, Xs)
    close(mu, g['fitc_mu'], 1e-9, 'fitc mu')
    close(var, g['fitc_var'], 1e-8, 'fitc var')
    mu, cov = R.gprfitc_predict(spec, X, Y, Z, noise, Xs, full_cov=True)
    close(mu, g['fitc_full_mu'], 1e-9, 'fitc full mu')
    close(cov, g['fitc_full_cov'], 1e-8, 'fitc full cov')


def test_functions(golden):
    g = golden('functions')
    rng = np.random.default_rng(11)
    M, N, K, d = 24, 31, 2, 3
    Xm = torch.tensor(rng.standard_normal((M, d)))
    Xn = torch.tensor(rng.standard_normal((N, d)))
    spec = dict(type='matern52', variance=torch.tensor(1.4, dtype=torch.float64),
                lengthscales=torch.tensor(1.3, dtype=torch.float64))
    f = torch.tensor(rng.standard_normal((M, K)))
    qd = torch.tensor(0.5 + rng.random((M, K)))
    qf = torch.tensor(np.stack([np.tril(rng.standard_normal((M, M))) * 0.3 + np.eye(M)
                                for _ in range(K)], 2))
    for white in (False, True):
        for full_cov in (False, True):
            for qname, q in (('none', None), ('diag', qd), ('full', qf)):
                mu, var = R.conditional(spec, Xn, Xm, f, full_cov=full_cov, q_sqrt=q, white=white)
                tag = 'cond/w%d_f%d_%s' % (white, full_cov, qname)
                close(mu, g[tag + '/mu'], 1e-9, tag + '/mu')
                close(var, g[tag + '/var'], 1e-9, tag + '/var')
    Kmm = R.K(spec, Xm) + torch.eye(M, dtype=torch.float64) * 1e-6
    for qname, q in (('diag', qd), ('full', qf)):
        close(R.gauss_kl(f, q), g['kl/white_' + qname], 1e-10, 'kl white ' + qname)
        close(R.gauss_kl(f, q, Kmm), g['kl/K_' + qname], 1e-9, 'kl K ' + qname)
    A = rng.standard_normal((M, M))
    L = np.linalg.cholesky(A @ A.T + M * np.eye(M))
    x = rng.standard_normal((M, 3))
    mu = rng.standard_normal((M, 3))
    close(R.multivariate_normal(torch.tensor(x), torch.tensor(mu), torch.tensor(L)), g['mvn'],
          1e-10, 'mvn')


def test_tf_adam_first_step():
    """First TF-Adam step moves every coordinate by lr*g/(|g| + eps*sqrt(1-b2)) ~= lr*sign(g)."""
    p = [torch.tensor([1.0, -2.0], dtype=torch.float64)]
    gr = [torch.tensor([0.5, -3.0], dtype=torch.float64)]
    new = R.tf_adam_step(p, gr, {}, lr=1e-3)
    close(new[0], np.array([1.0 - 1e-3, -2.0 + 1e-3]), 1e-6, 'adam')
