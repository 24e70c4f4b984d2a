"""TEST INFRASTRUCTURE ONLY -- parity cases shared by the golden-vector generator and the tests.

Every case is written ONCE against the reference's public API (`gpf.kernels.RBF(...)`,
`gpf.models.GPR(...)`, `m.objective`, `m.predict_f(...)` -- examples/gpr.py:48-56,
examples/svgp.py:142-160, README.md:17-30 of the reference).  `oracle/gen_golden.py` runs the
cases against the UNMODIFIED reference package (over `oracle/tf_shim`); `tests/` runs the very
same functions against the B200 package, which is only possible because that package is a
drop-in for the reference API.  `conv(a)` turns a numpy array into the tensor type of the
implementation under test (torch CPU for the shimmed reference, torch CUDA for the product).

Each case function returns a dict  name -> tensor  of scalar/array outputs, and a list of
(objective_name, model) for which gradients w.r.t. the unconstrained parameters are compared.
"""
import numpy as np


def synth_gpr(n, d, seed=0):
    """SURVEY.md section 8(d) synthetic GPR data: X~N(0,1), Y = sin(X.1/sqrt(D)) + 0.1 eps."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    Y = np.sin(X.sum(1, keepdims=True) / np.sqrt(d)) + 0.1 * rng.standard_normal((n, 1))
    return X, Y


def synth_svgp(n, d, m, seed=0):
    """SURVEY.md section 8(d) synthetic SVGP data: Y = sin(X.1/4) + 0.1 eps, Z = permuted rows."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    Y = np.sin(X.sum(1, keepdims=True) / 4.0) + 0.1 * rng.standard_normal((n, 1))
    Z = X[np.random.default_rng(2).permutation(n)[:m]].copy()
    return X, Y, Z


# --------------------------------------------------------------------------- kernels
def _kernel_zoo(gpf, d):
    k = gpf.kernels
    ls = 0.7 + 0.15 * np.arange(d)
    return [
        ('rbf_iso', lambda: k.RBF(d, variance=1.3, lengthscales=0.9, name='a')),
        ('rbf_ard', lambda: k.RBF(d, variance=0.8, lengthscales=ls, ARD=True, name='b')),
        ('m12_ard', lambda: k.Matern12(d, variance=1.1, lengthscales=ls, ARD=True, name='c')),
        ('m32_ard', lambda: k.Matern32(d, variance=0.6, lengthscales=ls, ARD=True, name='d')),
        ('m52_iso', lambda: k.Matern52(d, variance=1.7, lengthscales=1.4, name='e')),
        ('exp_ard', lambda: k.Exponential(d, variance=0.9, lengthscales=ls, ARD=True, name='f')),
        ('lin_iso', lambda: k.Linear(d, variance=0.4, name='g')),
        ('lin_ard', lambda: k.Linear(d, variance=0.3 + 0.1 * np.arange(d), ARD=True, name='h')),
        ('periodic', lambda: k.Periodic(d, period=1.7, variance=1.2, lengthscales=0.8, name='i')),
        ('rbf_active', lambda: k.RBF(2, variance=1.0, lengthscales=np.array([0.5, 1.5]), ARD=True,
                                     active_dims=[2, 0], name='j')),
        ('sum', lambda: k.RBF(d, lengthscales=ls, ARD=True, name='k1')
            + k.Linear(d, variance=0.2, name='k2') + 0.37),
        ('product', lambda: k.Matern32(d, lengthscales=1.2, name='k3')
            * k.Periodic(d, period=2.1, name='k4') * 1.9),
        ('sum_of_product', lambda: k.RBF(2, active_dims=[0, 1], name='k5')
            * k.Linear(1, active_dims=[2], name='k6') + k.Matern52(d, name='k7')),
    ]


def case_kernels(gpf, conv):
    """Gram matrices of every hot-path primitive and composition: K(X), K(X,X2), Kdiag(X)
    (kernels.py:408-439, 499-510, 562-610, 806-819, 1000-1084)."""
    d = 3
    rng = np.random.default_rng(10)
    X = rng.standard_normal((37, d)) * 1.3
    X2 = rng.standard_normal((23, d)) * 1.3
    out = {}
    for name, make in _kernel_zoo(gpf, d):
        kern = make()
        out[name + '/K'] = kern.K(conv(X))
        out[name + '/K2'] = kern.K(conv(X), conv(X2))
        out[name + '/Kdiag'] = kern.Kdiag(conv(X))
    return out, []


def _kernel_zoo_extra(gpf, d):
    """the remaining covariances of the reference
    (kernels.py:308-357 static, :447-471 RatQuad, :518-554 Polynomial, :617-646 Cosine,
    :649-766 ArcCosine, :822-881 Coregion, :943-970 TPS) alone and inside Sum / Product."""
    k = gpf.kernels
    ls = 0.7 + 0.15 * np.arange(d)

    def cosine(**kw):
        np.random.seed(5)                       # Cosine draws its weights from numpy's global RNG
        return k.Cosine(d, **kw)
    return [
        ('white', lambda: k.White(d, variance=0.7, name='xa')),
        ('constant', lambda: k.Constant(d, variance=1.9, name='xb')),
        ('bias', lambda: k.Bias(d, variance=0.3, name='xc')),
        ('ratquad_iso', lambda: k.RatQuad(d, alpha=1.7, variance=1.2, lengthscales=0.9, name='xd')),
        ('ratquad_ard', lambda: k.RatQuad(d, alpha=0.6, variance=0.8, lengthscales=ls, ARD=True,
                                          name='xe')),
        ('poly3', lambda: k.Polynomial(d, degree=3.0, variance=0.4, offset=0.8, name='xf')),
        ('poly2_ard', lambda: k.Polynomial(d, degree=2.0, variance=0.3 + 0.1 * np.arange(d),
                                           offset=1.3, ARD=True, name='xg')),
        ('cosine_iso', lambda: cosine(variance=1.1, lengthscales=1.3, name='xh')),
        ('cosine_ard', lambda: cosine(variance=0.9, lengthscales=ls, ARD=True, name='xi')),
        ('arccos0', lambda: k.ArcCosine(d, order=0, variance=1.2, weight_variances=0.7,
                                        bias_variance=0.4, name='xj')),
        ('arccos1_ard', lambda: k.ArcCosine(d, order=1, variance=0.8,
                                            weight_variances=0.5 + 0.2 * np.arange(d),
                                            bias_variance=1.1, ARD=True, name='xk')),
        ('arccos2', lambda: k.ArcCosine(d, order=2, variance=0.6, weight_variances=1.3,
                                        bias_variance=0.9, name='xl')),
        ('tps', lambda: k.TPS(d, variance=0.5, name='xm')),
        ('ratquad_active', lambda: k.RatQuad(2, alpha=2.0, lengthscales=np.array([0.5, 1.5]), ARD=True,
                                             active_dims=[2, 0], name='xn')),
        ('mixed_sum', lambda: k.RBF(d, lengthscales=ls, ARD=True, name='xo1')
            + k.RatQuad(d, alpha=1.2, name='xo2') + k.White(d, variance=0.05, name='xo3')
            + k.Linear(d, variance=0.2, name='xo4') + 0.11),
        ('mixed_product', lambda: k.Matern32(d, lengthscales=1.2, name='xp1')
            * k.Polynomial(d, degree=2.0, offset=0.6, name='xp2') * k.Bias(d, variance=1.4, name='xp3')),
        ('sum_of_mixed_product', lambda: k.ArcCosine(2, order=1, active_dims=[0, 1], name='xq1')
            * k.RBF(1, active_dims=[2], name='xq2') + k.Matern52(d, name='xq3')
            + k.Constant(d, variance=0.2, name='xq4')),
    ]


def case_kernels_extra(gpf, conv):
    """Gram matrices of the rank-4 covariances, and the gradient of a weighted sum of their
    entries w.r.t. every unconstrained parameter (the composed kernels' backward)."""
    import torch
    d = 3
    rng = np.random.default_rng(40)
    X = rng.standard_normal((33, d)) * 1.2
    X2 = rng.standard_normal((21, d)) * 1.2
    W, W2 = rng.standard_normal((33, 33)), rng.standard_normal((33, 21))
    W = W + W.T
    out = {}
    for name, make in _kernel_zoo_extra(gpf, d):
        kern = make()
        K, K2, Kd = kern.K(conv(X)), kern.K(conv(X), conv(X2)), kern.Kdiag(conv(X))
        out[name + '/K'], out[name + '/K2'], out[name + '/Kdiag'] = K, K2, Kd
        # TPS has no finite gradient on its diagonal (sqrt at exactly 0, kernels.py:965), and the
        # order-0 ArcCosine differentiates acos at 1 - 1e-15 there (:752, slope 2e7 times the
        # rounding noise of cos_theta, i.e. ill-conditioned in the reference itself): their
        # backward is exercised on the cross-covariance only
        val = (K2 * conv(W2)).sum() + (Kd * conv(W[:, 0])).sum()
        if name not in ('tps', 'arccos0'):
            val = val + (K * conv(W)).sum()
        params = []
        for p in kern.parameters:               # Polynomial lists its variance twice (:541)
            if not any(p is q for q in params):
                params.append(p)
        gs = torch.autograd.grad(val, [p.unconstrained_tensor for p in params], allow_unused=True)
        for i, (p, g) in enumerate(zip(params, gs)):
            out['%s/grad%d' % (name, i)] = torch.zeros_like(p.unconstrained_tensor) if g is None else g
    # Coregion: integer-coded inputs in its own column
    kc = gpf.kernels.Coregion(1, output_dim=4, rank=2, active_dims=[1], name='xr')
    with torch.no_grad():
        w = kc._W.unconstrained_tensor
        w.copy_(torch.as_tensor(rng.standard_normal(tuple(w.shape))).to(w))
    Xc = np.stack([rng.standard_normal(19), rng.integers(0, 4, 19).astype(np.float64)], 1)
    Xc2 = np.stack([rng.standard_normal(11), rng.integers(0, 4, 11).astype(np.float64)], 1)
    out['coregion/K'], out['coregion/K2'] = kc.K(conv(Xc)), kc.K(conv(Xc), conv(Xc2))
    out['coregion/Kdiag'] = kc.Kdiag(conv(Xc))
    gs = torch.autograd.grad((out['coregion/K2'] ** 2).sum(), [p.unconstrained_tensor for p in kc.parameters])
    out['coregion/grad0'], out['coregion/grad1'] = gs
    # Kdim / dimwise helpers (kernels.py:287-306, :441-444)
    kr = gpf.kernels.RBF(d, variance=1.3, lengthscales=0.7 + 0.15 * np.arange(d), ARD=True, name='xs')
    out['kdim/K'] = kr.Kdim(1, conv(X[:, 1:2]), conv(X2[:, 1:2]))
    with torch.no_grad():       # dimwise() feeds constrained TENSORS to Parameter (numpy conversion)
        kdw = kr.dimwise(2)
    out['dimwise/K'] = kdw.K(conv(X[:, 2:3]))
    return out, []


def case_gpr_composed(gpf, conv):
    """GPR whose covariance mixes fused primitives with composed kernels (RatQuad + Linear *
    Bias + White): objective, gradients and predictions through the op-by-op path
    (models/gpr.py:55-72, 118-131)."""
    n, d = 257, 4
    X, Y = synth_gpr(n, d, seed=14)
    Xs = np.random.default_rng(15).standard_normal((23, d))
    k = gpf.kernels
    kern = k.RatQuad(d, alpha=1.5, lengthscales=1.8, variance=0.9, name='gc_a') \
        + k.Linear(d, variance=0.3, name='gc_b') * k.Bias(d, variance=0.7, name='gc_c') \
        + k.White(d, variance=0.02, name='gc_d')
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern, obs_var=0.08, name='gc')
    out = {'objective': m.objective}
    out['pred_mu'], out['pred_var'] = m.predict_f(conv(Xs))
    out['full_mu'], out['full_cov'] = m.predict_f_full_cov(conv(Xs))
    return out, [('objective', m)]


def case_gpr_white(gpf, conv):
    """The common `RBF + White + Bias` covariance in a GPR (kernels.py:328-357, :1071-1077): the
    White term puts its variance on the diagonal of K(X) only, the Bias term is a constant --
    both are foldable into the fused one-call objective (White into the noise, Bias as a constant
    op).  Objective, all five gradients, predictions incl. full covariance."""
    n, d = 180, 3
    X, Y = synth_gpr(n, d, seed=25)
    Y = np.concatenate([Y, np.cos(X[:, :1]) + 0.05 * np.random.default_rng(26).standard_normal((n, 1))], 1)
    Xs = np.random.default_rng(27).standard_normal((13, d))
    k = gpf.kernels
    kern = k.RBF(d, ARD=True, lengthscales=1.6, variance=1.1, name='gw_a') + k.White(d, variance=0.05, name='gw_b') \
        + k.Bias(d, variance=0.3, name='gw_c')
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern, obs_var=0.07, name='gw')
    out = {'objective': m.objective, 'K': kern.K(conv(X[:9])), 'K2': kern.K(conv(X[:9]), conv(Xs)),
           'Kdiag': kern.Kdiag(conv(Xs))}
    out['pred_mu'], out['pred_var'] = m.predict_f(conv(Xs))
    out['full_mu'], out['full_cov'] = m.predict_f_full_cov(conv(Xs))
    return out, [('objective', m)]


def case_large_d(gpf, conv):
    """Inputs wider than the fused Gram kernel's per-primitive table (40 columns; the reference's
    examples/svgp.py feeds 100 network features): Grams, parameter and input gradients of the
    hot-path primitives and a composition, a GPR and an SVGP on them."""
    import torch
    d = 40
    rng = np.random.default_rng(60)
    X = rng.standard_normal((31, d)) * 0.6
    X2 = rng.standard_normal((19, d)) * 0.6
    W, W2 = rng.standard_normal((31, 31)), rng.standard_normal((31, 19))
    W = W + W.T
    k = gpf.kernels
    ls = 3.0 + 0.05 * np.arange(d)
    zoo = [('rbf_ard', lambda: k.RBF(d, variance=1.2, lengthscales=ls, ARD=True, name='la')),
           ('m32_iso', lambda: k.Matern32(d, variance=0.7, lengthscales=5.0, name='lb')),
           ('lin_ard', lambda: k.Linear(d, variance=0.02 + 0.001 * np.arange(d), ARD=True, name='lc')),
           ('periodic', lambda: k.Periodic(d, period=2.3, variance=1.1, lengthscales=3.0, name='ld')),
           ('composite', lambda: k.RBF(d, lengthscales=ls, ARD=True, name='le1')
            + k.Linear(d, variance=0.03, name='le2') * k.Matern52(d, lengthscales=6.0, name='le3') + 0.2)]
    out = {}
    for name, make in zoo:
        kern = make()
        Xg = conv(X).requires_grad_(True)
        K, K2, Kd = kern.K(Xg), kern.K(Xg, conv(X2)), kern.Kdiag(Xg)
        out[name + '/K'], out[name + '/K2'], out[name + '/Kdiag'] = K, K2, Kd
        val = (K * conv(W)).sum() + (K2 * conv(W2)).sum() + (Kd * conv(W[:, 0])).sum()
        params = [p.unconstrained_tensor for p in kern.parameters]
        gs = torch.autograd.grad(val, params + [Xg])
        for i, g in enumerate(gs[:-1]):
            out['%s/grad%d' % (name, i)] = g
        out[name + '/dX'] = gs[-1]
    n = 150
    Xr = rng.standard_normal((n, d)) * 0.6
    Yr = np.sin(Xr[:, :3].sum(1, keepdims=True)) + 0.1 * rng.standard_normal((n, 1))
    Xs = rng.standard_normal((11, d)) * 0.6
    m = gpf.models.GPR(conv(Xr), conv(Yr), kern=k.RBF(d, ARD=True, lengthscales=ls, name='lg_k'), name='lg')
    out['gpr/objective'] = m.objective
    out['gpr/pred_mu'], out['gpr/pred_var'] = m.predict_f(conv(Xs))
    Z = Xr[:20].copy()
    sv = gpf.models.SVGP(conv(Xr), conv(Yr), k.Matern52(d, ARD=True, lengthscales=ls, name='ls_k'),
                         gpf.likelihoods.Gaussian(var=0.2), Z=Z, name='ls')
    with torch.no_grad():
        qm = sv._q_mu.unconstrained_tensor
        qm.copy_(torch.as_tensor(0.3 * rng.standard_normal(tuple(qm.shape))).to(qm))
    out['svgp/objective'] = sv.objective
    out['svgp/pred_mu'], out['svgp/pred_var'] = sv.predict_f(conv(Xs))
    return out, [('gpr/objective', m), ('svgp/objective', sv)]


def nkn_c3_kernel(gpf, d, weights=None):
    """The section-8(d) NKN topology: k=6 primitives, Linear 6->8, Product 2, Linear 4->4,
    Product 2, Linear 2->1 (neural_kernel_network_wrapper.py:38-40 hparams schema).  The
    reference draws Linear weights from numpy's global RNG (wrapper.py:100-104), so it is
    seeded here; an implementation may instead be handed `weights` explicitly."""
    k = gpf.kernels
    prims = [
        k.RBF(d, ARD=True, name='p0'),
        k.RBF(d, lengthscales=2.0, ARD=True, name='p1'),
        k.Periodic(d, period=1.0, lengthscales=1.0, name='p2'),
        k.Periodic(d, period=2.0, name='p3'),
        k.Linear(d, ARD=True, name='p4'),
        k.Linear(d, ARD=True, name='p5'),
    ]
    hparams = [
        dict(name='Linear', params=dict(input_dim=6, output_dim=8, name='l0')),
        dict(name='Product', params=dict(input_dim=8, step=2, name='l1')),
        dict(name='Linear', params=dict(input_dim=4, output_dim=4, name='l2')),
        dict(name='Product', params=dict(input_dim=4, step=2, name='l3')),
        dict(name='Linear', params=dict(input_dim=2, output_dim=1, name='l4')),
    ]
    np.random.seed(0)
    wrapper = gpf.neural_kernel_network.NKNWrapper(hparams)
    return gpf.neural_kernel_network.NeuralKernelNetwork(d, prims, wrapper)


def case_nkn(gpf, conv):
    """NKN Gram + NKN-GPR objective, predict (neural_kernel_network.py:35-47)."""
    d, n = 3, 150
    X, Y = synth_gpr(n, d, seed=3)
    Xs = np.random.default_rng(4).standard_normal((20, d))
    kern = nkn_c3_kernel(gpf, d)
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern, name='nkn_gpr')
    out = {'K': kern.K(conv(X[:30])), 'K2': kern.K(conv(X[:30]), conv(Xs)),
           'Kdiag': kern.Kdiag(conv(X[:30])), 'objective': m.objective}
    mu, var = m.predict_f(conv(Xs))
    out['pred_mu'], out['pred_var'] = mu, var
    return out, [('objective', m)]


# --------------------------------------------------------------------------- GPR
def _gpr(gpf, conv, n, d, ls, nstar, name):
    X, Y = synth_gpr(n, d, seed=0)
    Xs = np.random.default_rng(1).standard_normal((nstar, d))
    kern = gpf.kernels.RBF(d, ARD=True, lengthscales=ls, name=name + '_k')
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern, name=name)
    out = {'objective': m.objective}
    mu, var = m.predict_f(conv(Xs))
    out['pred_mu'], out['pred_var'] = mu, var
    return out, [('objective', m)]


def case_gpr_c1(gpf, conv):
    """C1: GPR ARD-RBF N=1000 D=4, reference-default l=1 (examples/gpr.py:48-56)."""
    return _gpr(gpf, conv, 1000, 4, None, 1024, 'c1')


def case_gpr_c1_ls(gpf, conv):
    """C1 with l = sqrt(D): non-trivial off-diagonals."""
    return _gpr(gpf, conv, 1000, 4, 2.0, 64, 'c1b')


def case_gpr_c2_small(gpf, conv):
    """C2/C5 scaled down: GPR ARD-RBF N=2048 D=8, l = sqrt(8)."""
    return _gpr(gpf, conv, 2048, 8, np.sqrt(8.0), 64, 'c2s')


def case_gpr_misc(gpf, conv):
    """GPR with R=3 output columns, Matern32+Linear kernel, full_cov predict, predict_y and
    predict_density (models/gpr.py:118-131, models/model.py:121-166), N not a multiple of
    any tile size."""
    n, d = 301, 5
    rng = np.random.default_rng(7)
    X = rng.standard_normal((n, d))
    Y = np.stack([np.sin(X[:, 0]), np.cos(X[:, 1]) * X[:, 2], X.sum(1) * 0.3], 1) \
        + 0.05 * rng.standard_normal((n, 3))
    Xs = rng.standard_normal((17, d))
    Ys = rng.standard_normal((17, 3))
    kern = gpf.kernels.Matern32(d, lengthscales=1.5, name='gm_a') \
        + gpf.kernels.Linear(d, variance=0.1, name='gm_b')
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern, obs_var=0.05, name='gm')
    out = {'objective': m.objective}
    out['pred_mu'], out['pred_var'] = m.predict_f(conv(Xs))
    out['full_mu'], out['full_cov'] = m.predict_f_full_cov(conv(Xs))
    out['y_mu'], out['y_var'] = m.predict_y(conv(Xs))
    out['density'] = m.predict_density(conv(Xs), conv(Ys))
    return out, [('objective', m)]


def case_lbfgs(gpf, conv):
    """`GPModel.optimize` = the eager L-BFGS (LBFGS.py:44-338,
    models/model.py:172-195) on a small ARD-RBF GPR: objective history, final iterate, and the
    objective / predictions at the optimum."""
    import torch
    n, d = 150, 3
    X, Y = synth_gpr(n, d, seed=16)
    Xs = np.random.default_rng(17).standard_normal((15, d))
    kern = gpf.kernels.RBF(d, ARD=True, name='lb_k')
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern, name='lb')
    out = {'objective_before': m.objective}
    m.optimize()
    hist = m.LBFGS_opt.history
    out['f_hist'] = torch.stack([h[0].detach().reshape(()) for h in hist])
    out['x_final'] = hist[-1][2].detach()
    out['objective_after'] = m.objective
    out['pred_mu'], out['pred_var'] = m.predict_f(conv(Xs))
    return out, []


def case_lbfgs_rosenbrock(gpf, conv):
    """The LBFGS class on its own (LBFGS.py:44-338) minimising a 6-D Rosenbrock function with 5
    correction pairs: exercises memory eviction, both Zoom orientations and the iteration cap.
    Pure host logic -- no Gram / Cholesky -- so it is a CPU-only case (HOST_ONLY_CASES)."""
    import importlib
    import torch
    LBFGS = importlib.import_module(gpf.__name__ + '.LBFGS').LBFGS
    out = {}
    for tag, x0, kw in (('a', [-1.2, 1.0, 0.8, -0.5, 1.7, 0.3], dict(max_iter=60, nCorrection=5, tolFun=1e-10)),
                        ('b', [0.5, -0.4, 2.0], dict(max_iter=25, nCorrection=100, tolFun=1e-7))):
        w = conv(np.array(x0)).clone().requires_grad_(True)

        def opfunc(w=w):
            f = (100.0 * (w[1:] - w[:-1] ** 2) ** 2 + (1.0 - w[:-1]) ** 2).sum()
            (g,) = torch.autograd.grad(f, [w])
            return f, [(g, w)]
        opt = LBFGS(opfunc, **kw)
        ret = opt.run()
        out[tag + '/f_hist'] = torch.stack([h[0].detach().reshape(()) for h in opt.history])
        out[tag + '/x_hist'] = torch.stack([h[2].detach() for h in opt.history])
        out[tag + '/x_end'] = w.detach().clone()
        out[tag + '/n_eval'] = torch.tensor(float(ret[2]) if len(ret) > 2 else -1.0)
    return out, []


def case_gpr_features(gpf, conv):
    """GPR's feature path (models/gpr.py:62-66, :84-114): a random-feature kernel
    (kernel_kitchen_sink.SamplerKernel over RBFSampler / LinearSampler) makes GPR use
    densities.multivariate_normal_feature (densities.py:98-123) and the Woodbury predictor.
    Also the two functions the reference's own tests compare (densities.py:159-174,
    models/gpr.py:135-203): the feature path against the Cholesky path on the same K = C C^T."""
    import torch
    ks = gpf.kernel_kitchen_sink
    n, d, nc = 90, 3, 12
    X, Y = synth_gpr(n, d, seed=18)
    Y = np.concatenate([Y, 0.5 * Y ** 2], 1)
    Xs = np.random.default_rng(19).standard_normal((14, d))
    out = {}
    for tag, make in (('rbf', lambda: ks.RBFSampler(d, ls=1.3, var=0.9, n_components=nc)),
                      ('lin', lambda: ks.LinearSampler(d, var=0.7, n_components=6))):
        np.random.seed(7)
        sampler = make()
        kern = ks.SamplerKernel(sampler)
        m = gpf.models.GPR(conv(X), conv(Y), kern=kern, obs_var=0.4, name='gf_' + tag)
        obj = m.objective
        out[tag + '/objective'] = obj
        out[tag + '/K'] = kern.K(conv(X[:20]), conv(Xs))
        out[tag + '/Kdiag'] = kern.Kdiag(conv(Xs))
        out[tag + '/pred_mu'], out[tag + '/pred_var'] = m.predict_f(conv(Xs))
        out[tag + '/full_mu'], out[tag + '/full_cov'] = m.predict_f_full_cov(conv(Xs))
        params = [sampler._variance] + ([sampler._ls] if tag == 'rbf' else []) + list(m.likelihood.parameters)
        gs = torch.autograd.grad(obj, [p.unconstrained_tensor for p in params])
        for i, g in enumerate(gs):
            out['%s/grad%d' % (tag, i)] = g
        # the same density through the N x N Cholesky (densities.py:73-95)
        C = kern.features(conv(X))
        var = m.likelihood.variance
        Kfull = kern.K(conv(X)) + var * conv(np.eye(n))
        out[tag + '/mvn_feature'] = gpf.densities.multivariate_normal_feature(
            conv(Y[:, :1]), conv(np.zeros((n, 1))), C, var)
        out[tag + '/mvn_cholesky'] = gpf.densities.multivariate_normal(
            conv(Y[:, :1]), conv(np.zeros((n, 1))), torch.linalg.cholesky(Kfull))
    return out, []


def case_priors(gpf, conv):
    """Parameter priors (priors.py:31-124) entering the objective with the transform's
    log-Jacobian (params.py:176-194, models/model.py:57-73): a GPR whose kernel variance,
    lengthscales and noise carry Gamma / LogNormal / Gaussian priors, plus every prior's logp
    on fixed values."""
    n, d = 120, 3
    X, Y = synth_gpr(n, d, seed=23)
    kern = gpf.kernels.RBF(d, ARD=True, lengthscales=1.4, name='pr_k')
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern, name='pr')
    P = gpf.priors
    kern._variance.prior = P.Gamma(2.0, 1.5)
    kern._ls.prior = P.LogNormal(0.2, 0.8)
    m.likelihood.parameters[0].prior = P.Gaussian(0.1, 0.05)
    out = {'objective': m.objective, 'prior_tensor': m.prior_tensor,
           'likelihood_tensor': m.likelihood_tensor}
    rng = np.random.default_rng(24)
    xs = {'real': rng.standard_normal(7), 'positive': 0.1 + rng.gamma(2.0, 1.0, 7),
          'unit': np.clip(rng.random(7), 0.05, 0.95)}
    for name, prior, kind in (('gaussian', P.Gaussian(0.3, 1.7), 'real'),
                              ('lognormal', P.LogNormal(-0.2, 0.6), 'positive'),
                              ('gamma', P.Gamma(2.5, 0.7), 'positive'),
                              ('laplace', P.Laplace(0.1, 1.3), 'real'),
                              ('beta', P.Beta(2.0, 3.0), 'unit'),
                              ('uniform', P.Uniform(-1.0, 3.0), 'real')):
        out['logp/' + name] = prior.logp(conv(xs[kind]))
    return out, [('objective', m)]


# --------------------------------------------------------------------------- SVGP / SGPR
def _svgp(gpf, conv, n, d, minducing, batch, whiten, q_diag, latents, name, ls=None):
    X, Y, Z = synth_svgp(n, d, minducing, seed=0)
    if latents > 1:
        Y = np.concatenate([Y * (1 + 0.3 * j) + 0.1 * j for j in range(latents)], 1)
    rng = np.random.default_rng(5)
    kern = gpf.kernels.RBF(d, ARD=True, lengthscales=ls if ls is not None else np.sqrt(d),
                           name=name + '_k')
    lik = gpf.likelihoods.Gaussian(var=0.1)
    Xb, Yb = X[:batch], Y[:batch]
    m = gpf.models.SVGP(conv(Xb), conv(Yb), kern, lik, Z=Z.copy(), whiten=whiten,
                        q_diag=q_diag, num_data=n, name=name)
    # move q_mu / q_sqrt off their trivial initial values (svgp.py:81-89) so every term of the
    # bound and of base_conditional is exercised
    import torch
    with torch.no_grad():
        qm = m._q_mu.unconstrained_tensor
        qm.copy_(torch.as_tensor(0.3 * rng.standard_normal(tuple(qm.shape))).to(qm))
        qs = m._q_sqrt.unconstrained_tensor
        qs.add_(torch.as_tensor(0.05 * rng.standard_normal(tuple(qs.shape))).to(qs))
    out = {'objective': m.objective, 'KL': m.build_prior_KL()}
    Xs = rng.standard_normal((19, d))
    out['pred_mu'], out['pred_var'] = m.predict_f(conv(Xs))
    return out, [('objective', m)], m, Xs


def case_svgp_white_full(gpf, conv):
    """SVGP default config whiten=True, q_diag=False (svgp.py:45-130) at a scaled-down C4."""
    o, g, m, Xs = _svgp(gpf, conv, 4000, 16, 256, 1024, True, False, 1, 'sv1', ls=4.0)
    o['full_mu'], o['full_cov'] = m.predict_f_full_cov(conv(Xs))
    return o, g


def case_svgp_nonwhite_full(gpf, conv):
    """whiten=False, q_diag=False, 2 latent GPs (conditionals.py:99-100, KL :92-103)."""
    o, g, _, _ = _svgp(gpf, conv, 500, 4, 40, 200, False, False, 2, 'sv2')
    return o, g


def case_svgp_white_diag(gpf, conv):
    """whiten=True, q_diag=True, 2 latent GPs (conditionals.py:106-107)."""
    o, g, _, _ = _svgp(gpf, conv, 500, 4, 40, 200, True, True, 2, 'sv3')
    return o, g


def case_svgp_nonwhite_diag(gpf, conv):
    """whiten=False, q_diag=True (kullback_leiblers.py:84-91)."""
    o, g, _, _ = _svgp(gpf, conv, 500, 4, 40, 200, False, True, 1, 'sv4')
    return o, g


def case_sgpr(gpf, conv):
    """SGPR collapsed bound and predictions (models/sgpr.py:121-189)."""
    n, d, mi = 600, 4, 50
    X, Y, Z = synth_svgp(n, d, mi, seed=8)
    Xs = np.random.default_rng(9).standard_normal((21, d))
    kern = gpf.kernels.RBF(d, ARD=True, lengthscales=2.0, name='sg_k')
    m = gpf.models.SGPR(conv(X), conv(Y), kern, Z=Z.copy(), obs_var=0.1, name='sg')
    out = {'objective': m.objective}
    out['pred_mu'], out['pred_var'] = m.predict_f(conv(Xs))
    out['full_mu'], out['full_cov'] = m.predict_f_full_cov(conv(Xs))
    return out, [('objective', m)]


def _make_fitc(gpf, *args, **kwargs):
    """The reference's GPRFITC.__init__ reads self.name before GPModel.__init__ has set it
    (models/sgpr.py:221) and so cannot be constructed as shipped; giving the instance its name
    first lets the UNMODIFIED constructor run.  Harmless for an implementation without the bug."""
    cls = gpf.models.GPRFITC
    m = cls.__new__(cls)
    m._name = kwargs.get('name', 'GPModel')
    cls.__init__(m, *args, **kwargs)
    return m


def case_sparse_bounds(gpf, conv):
    """the upper bound of SGPRUpperMixin (models/sgpr.py:55-82) on
    SGPR and GPRFITC, and the FITC likelihood / predictions (models/sgpr.py:192-317)."""
    n, d, mi = 500, 3, 40
    X, Y, Z = synth_svgp(n, d, mi, seed=12)
    rng = np.random.default_rng(13)
    Y = np.concatenate([Y, rng.standard_normal((n, 1)) * 0.3], 1)
    Xs = rng.standard_normal((17, d))
    ks = gpf.kernels.Matern32(d, ARD=True, lengthscales=1.5, variance=1.2, name='ub_k')
    sg = gpf.models.SGPR(conv(X), conv(Y), ks, Z=Z.copy(), obs_var=0.2, name='ub_sg')
    kf = gpf.kernels.RBF(d, ARD=True, lengthscales=1.5, name='fitc_k')
    fitc = _make_fitc(gpf, conv(X), conv(Y), kf, Z=Z.copy(), obs_var=0.2, name='fitc')
    out = {'sgpr_upper': sg.compute_upper_bound(), 'sgpr_lower': sg.likelihood_tensor,
           'fitc_upper': fitc.compute_upper_bound(), 'fitc_objective': fitc.objective}
    out['fitc_mu'], out['fitc_var'] = fitc.predict_f(conv(Xs))
    out['fitc_full_mu'], out['fitc_full_cov'] = fitc.predict_f_full_cov(conv(Xs))
    return out, [('fitc_objective', fitc)]


def case_mc_models(gpf, conv):
    """The whitened-latent models around the same Cholesky / conditional kernels: GPMC
    (models/gpmc.py:28-95, Bernoulli likelihood), SGPMC (models/sgpmc.py:25-104, Poisson), each
    with its N(0, I) prior on V in the objective; an SVGP over Multiscale inducing features
    (features.py:89-150); and an SVGP whose covariance is a composed kernel (RatQuad + White),
    which takes the elementwise-jitter route for Kuu."""
    import torch
    rng = np.random.default_rng(50)
    out, grads = {}, []

    def shake(param, scale):
        with torch.no_grad():
            t = param.unconstrained_tensor
            t.add_(torch.as_tensor(scale * rng.standard_normal(tuple(t.shape))).to(t))

    n, d = 60, 2
    X = rng.standard_normal((n, d))
    Yb = (np.sin(X.sum(1, keepdims=True)) + 0.3 * rng.standard_normal((n, 1)) > 0).astype(np.float64)
    Xs = rng.standard_normal((9, d))
    m1 = gpf.models.GPMC(conv(X), conv(Yb), gpf.kernels.Matern32(d, lengthscales=1.3, name='mc1_k'),
                         gpf.likelihoods.Bernoulli(), name='mc1')
    shake(m1._V, 0.7)
    out['gpmc/objective'], out['gpmc/prior'] = m1.objective, m1.prior_tensor
    out['gpmc/pred_mu'], out['gpmc/pred_var'] = m1.predict_f(conv(Xs))
    out['gpmc/y_mu'], out['gpmc/y_var'] = m1.predict_y(conv(Xs))
    grads.append(('gpmc/objective', m1))

    n2, mi = 80, 15
    X2 = rng.standard_normal((n2, d))
    Yc = rng.poisson(np.exp(0.5 * np.sin(X2.sum(1, keepdims=True)) + 0.3)).astype(np.float64)
    Z = X2[:mi].copy()
    m2 = gpf.models.SGPMC(conv(X2), conv(Yc), gpf.kernels.RBF(d, ARD=True, lengthscales=1.1, name='mc2_k'),
                          gpf.likelihoods.Poisson(), Z=Z.copy(), name='mc2')
    shake(m2._V, 0.5)
    out['sgpmc/objective'] = m2.objective
    out['sgpmc/pred_mu'], out['sgpmc/pred_var'] = m2.predict_f(conv(Xs))
    out['sgpmc/full_mu'], out['sgpmc/full_cov'] = m2.predict_f_full_cov(conv(Xs))
    grads.append(('sgpmc/objective', m2))

    Yr = np.sin(X2.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((n2, 1))
    feat = gpf.features.Multiscale(Z.copy(), 0.3 + 0.4 * rng.random(Z.shape))
    m3 = gpf.models.SVGP(conv(X2), conv(Yr), gpf.kernels.RBF(d, ARD=True, lengthscales=1.2, name='mc3_k'),
                         gpf.likelihoods.Gaussian(var=0.2), feat=feat, whiten=False, name='mc3')
    shake(m3._q_mu, 0.3)
    shake(m3._q_sqrt, 0.05)
    obj3 = m3.objective
    out['multiscale/objective'] = obj3
    out['multiscale/Kuu'] = feat.Kuu(m3.kern, jitter=1e-6)
    out['multiscale/Kuf'] = feat.Kuf(m3.kern, conv(Xs))
    out['multiscale/pred_mu'], out['multiscale/pred_var'] = m3.predict_f(conv(Xs))
    out['multiscale/grad_scales'], = torch.autograd.grad(obj3, [feat._scales.unconstrained_tensor])
    grads.append(('multiscale/objective', m3))

    kc = gpf.kernels.RatQuad(d, alpha=1.4, lengthscales=1.2, name='mc4_a') \
        + gpf.kernels.White(d, variance=0.05, name='mc4_b')
    m4 = gpf.models.SVGP(conv(X2), conv(Yr), kc, gpf.likelihoods.Gaussian(var=0.2), Z=Z.copy(), name='mc4')
    shake(m4._q_mu, 0.3)
    shake(m4._q_sqrt, 0.05)
    out['svgp_composed/objective'] = m4.objective
    out['svgp_composed/pred_mu'], out['svgp_composed/pred_var'] = m4.predict_f(conv(Xs))
    grads.append(('svgp_composed/objective', m4))
    return out, grads


# --------------------------------------------------------------------------- likelihoods
def _lik_zoo(gpf):
    L = gpf.likelihoods
    return [('bernoulli', L.Bernoulli(), 'binary'), ('poisson', L.Poisson(), 'count'),
            ('poisson_bin', L.Poisson(binsize=0.5), 'count'), ('exponential', L.Exponential(), 'positive'),
            ('studentt', L.StudentT(deg_free=4.0), 'real'), ('gamma', L.Gamma(), 'positive'),
            ('beta', L.Beta(scale=2.0), 'unit'), ('gaussian', L.Gaussian(var=0.3), 'real')]


def case_likelihoods(gpf, conv):
    """every likelihood's logp / conditional moments / variational
    expectations / predictive mean, variance and density on fixed (Fmu, Fvar, Y)
    (likelihoods.py:47-151 Gauss-Hermite defaults and the closed forms of the subclasses), and
    MultiClass with the RobustMax link (:379-489).  Pure elementwise functions: no model."""
    rng = np.random.default_rng(31)
    n = 23
    Fmu, Fvar = rng.standard_normal((n, 1)) * 0.8, 0.05 + rng.random((n, 1))
    F = rng.standard_normal((n, 1))
    Ys = {'binary': (rng.random((n, 1)) < 0.5).astype(np.float64),
          'count': rng.poisson(2.0, (n, 1)).astype(np.float64),
          'positive': 0.1 + rng.gamma(2.0, 1.0, (n, 1)),
          'unit': np.clip(rng.random((n, 1)), 0.05, 0.95),
          'real': rng.standard_normal((n, 1))}
    out = {}
    for name, lik, kind in _lik_zoo(gpf):
        Y = conv(Ys[kind])
        out[name + '/logp'] = lik.logp(conv(F), Y)
        out[name + '/cmean'] = lik.conditional_mean(conv(F))
        out[name + '/cvar'] = lik.conditional_variance(conv(F))
        out[name + '/varexp'] = lik.variational_expectations(conv(Fmu), conv(Fvar), Y)
        out[name + '/pmean'], out[name + '/pvar'] = lik.predict_mean_and_var(conv(Fmu), conv(Fvar))
        out[name + '/pdens'] = lik.predict_density(conv(Fmu), conv(Fvar), Y)
    k = 4
    mc = gpf.likelihoods.MultiClass(k)
    Fk, Vk = rng.standard_normal((n, k)), 0.05 + rng.random((n, k))
    lab = conv(rng.integers(0, k, (n, 1)).astype(np.float64))
    out['multiclass/logp'] = mc.logp(conv(Fk), lab)
    out['multiclass/cmean'] = mc.conditional_mean(conv(Fk))
    out['multiclass/cvar'] = mc.conditional_variance(conv(Fk))
    out['multiclass/varexp'] = mc.variational_expectations(conv(Fk), conv(Vk), lab)
    out['multiclass/pmean'], out['multiclass/pvar'] = mc.predict_mean_and_var(conv(Fk), conv(Vk))
    out['multiclass/pdens'] = mc.predict_density(conv(Fk), conv(Vk), lab)
    return out, []


def case_likelihoods_extra(gpf, conv):
    """The Ordinal likelihood (likelihoods.py:554-631) through every base-class entry point, and
    the multivariate Gauss-Hermite helper mvnquad (quadrature.py:30-77)."""
    import torch
    rng = np.random.default_rng(33)
    n = 19
    Fmu, Fvar = rng.standard_normal((n, 1)) * 1.2, 0.05 + rng.random((n, 1))
    F = rng.standard_normal((n, 1)) * 1.5
    Y = conv(rng.integers(0, 4, (n, 1)).astype(np.float64))
    lik = gpf.likelihoods.Ordinal(np.array([-1.0, 0.2, 1.1]))
    # TensorFlow converts the numpy bin edges on the fly in `bin_edges / sigma` (:600); torch does
    # not divide ndarray by Tensor, so the attribute is handed over as a tensor of the same values
    lik.bin_edges = conv(lik.bin_edges)
    out = {'ordinal/logp': lik.logp(conv(F), Y),
           'ordinal/cmean': lik.conditional_mean(conv(F)),
           'ordinal/cvar': lik.conditional_variance(conv(F)),
           'ordinal/varexp': lik.variational_expectations(conv(Fmu), conv(Fvar), Y),
           'ordinal/pdens': lik.predict_density(conv(Fmu), conv(Fvar), Y)}
    out['ordinal/pmean'], out['ordinal/pvar'] = lik.predict_mean_and_var(conv(Fmu), conv(Fvar))
    D = 2
    means = rng.standard_normal((5, D))
    A = rng.standard_normal((5, D, D))
    covs = A @ A.transpose(0, 2, 1) + 0.3 * np.eye(D)
    out['mvnquad/scalar'] = gpf.quadrature.mvnquad(lambda x: torch.sin(x[:, 0]) * torch.exp(0.3 * x[:, 1]),
                                                   conv(means), conv(covs), 7, D)
    out['mvnquad/vector'] = gpf.quadrature.mvnquad(lambda x: torch.stack([x[:, 0] ** 2, x[:, 0] * x[:, 1], x[:, 1]], 1),
                                                   conv(means), conv(covs), 5, D, Dout=(3,))
    return out, []


def case_svgp_multiclass(gpf, conv):
    """The model of examples/svgp.py:142-146 at test size: SVGP with the MultiClass likelihood,
    one latent GP per class, whiten=False; bound, gradients, class probabilities and the
    predictive log density the example reports (:151-155)."""
    n, d, mi, k, batch = 400, 4, 30, 4, 160
    X, _, Z = synth_svgp(n, d, mi, seed=21)
    rng = np.random.default_rng(22)
    W = rng.standard_normal((d, k))
    lab = np.argmax(X @ W + 0.3 * rng.standard_normal((n, k)), 1).astype(np.float64)[:, None]
    kern = gpf.kernels.RBF(d, ARD=True, lengthscales=1.7, name='mc_k')
    lik = gpf.likelihoods.MultiClass(k)
    m = gpf.models.SVGP(conv(X[:batch]), conv(lab[:batch]), kern, lik, Z=Z.copy(), num_latent=k,
                        whiten=False, num_data=n, name='mc')
    import torch
    with torch.no_grad():
        qm = m._q_mu.unconstrained_tensor
        qm.copy_(torch.as_tensor(0.5 * rng.standard_normal(tuple(qm.shape))).to(qm))
        qs = m._q_sqrt.unconstrained_tensor
        qs.add_(torch.as_tensor(0.05 * rng.standard_normal(tuple(qs.shape))).to(qs))
    Xs, labs = conv(X[batch:batch + 25]), conv(lab[batch:batch + 25])
    out = {'objective': m.objective}
    fmu, fvar = m._build_predict(Xs)
    out['class_prob'], _ = m.likelihood.predict_mean_and_var(fmu, fvar)
    out['pred_density'] = m.likelihood.predict_density(fmu, fvar, labs)
    return out, [('objective', m)]


# --------------------------------------------------------------------------- free functions
def case_functions(gpf, conv):
    """base_conditional (conditionals.py:81-121), conditional (:25-66), gauss_kl
    (kullback_leiblers.py:26-105) and multivariate_normal (densities.py:73-95) called
    directly, all q_sqrt / white / full_cov variants."""
    rng = np.random.default_rng(11)
    M, N, K, d = 24, 31, 2, 3
    Xm = rng.standard_normal((M, d))
    Xn = rng.standard_normal((N, d))
    kern = gpf.kernels.Matern52(d, lengthscales=1.3, variance=1.4, name='fn_k')
    f = rng.standard_normal((M, K))
    qd = 0.5 + rng.random((M, K))
    qf = np.stack([np.tril(rng.standard_normal((M, M))) * 0.3 + np.eye(M) for _ in range(K)], 2)
    out = {}
    for white in (False, True):
        for full_cov in (False, True):
            for qname, q in (('none', None), ('diag', qd), ('full', qf)):
                mu, var = gpf.conditionals.conditional(
                    conv(Xn), conv(Xm), kern, conv(f), full_cov=full_cov,
                    q_sqrt=None if q is None else conv(q), white=white)
                tag = 'cond/w%d_f%d_%s' % (white, full_cov, qname)
                out[tag + '/mu'], out[tag + '/var'] = mu, var
    Kmm = kern.K(conv(Xm)) + conv(np.eye(M) * 1e-6)
    for qname, q in (('diag', qd), ('full', qf)):
        out['kl/white_' + qname] = gpf.kullback_leiblers.gauss_kl(conv(f), conv(q))
        out['kl/K_' + qname] = gpf.kullback_leiblers.gauss_kl(conv(f), conv(q), Kmm)
    A = rng.standard_normal((M, M))
    L = np.linalg.cholesky(A @ A.T + M * np.eye(M))
    x = rng.standard_normal((M, 3))
    mu = rng.standard_normal((M, 3))
    out['mvn'] = gpf.densities.multivariate_normal(conv(x), conv(mu), conv(L))
    return out, []


CASES = {
    'kernels': case_kernels,
    'kernels_extra': case_kernels_extra,
    'gpr_composed': case_gpr_composed,
    'large_d': case_large_d,
    'gpr_white': case_gpr_white,
    'gpr_features': case_gpr_features,
    'priors': case_priors,
    'mc_models': case_mc_models,
    'lbfgs': case_lbfgs,
    'lbfgs_rosenbrock': case_lbfgs_rosenbrock,
    'nkn': case_nkn,
    'gpr_c1': case_gpr_c1,
    'gpr_c1_ls': case_gpr_c1_ls,
    'gpr_c2_small': case_gpr_c2_small,
    'gpr_misc': case_gpr_misc,
    'svgp_white_full': case_svgp_white_full,
    'svgp_nonwhite_full': case_svgp_nonwhite_full,
    'svgp_white_diag': case_svgp_white_diag,
    'svgp_nonwhite_diag': case_svgp_nonwhite_diag,
    'sgpr': case_sgpr,
    'sparse_bounds': case_sparse_bounds,
    'likelihoods': case_likelihoods,
    'svgp_multiclass': case_svgp_multiclass,
    'likelihoods_extra': case_likelihoods_extra,
    'functions': case_functions,
}

# Cases added after the last session that had GPU time.  tests/test_gpu_parity.py runs the rest,
# tests/test_gpu_zz_widened.py (sorted last, so a surprise there cannot mask the established
# tests under `pytest -x`) runs these; once seen green on a B200 they simply leave this tuple.
LATE_CASES = ('kernels_extra', 'gpr_composed', 'lbfgs', 'gpr_features', 'priors', 'mc_models', 'likelihoods_extra', 'large_d', 'gpr_white')
# Pure host logic (no library call): checked on the CPU only.
HOST_ONLY_CASES = ('lbfgs_rosenbrock',)


def run_case(gpf, name, conv):
    """Run one case; returns {key: numpy array} including 'grad/<objective>/<i>' entries, the
    gradient of each listed objective w.r.t. the i-th entry of `model.parameters`
    (unconstrained tensors, reference order: models/model.py:119, svgp.py:91)."""
    import torch
    out, grads = CASES[name](gpf, conv)
    res = {}
    for key, val in out.items():
        res[key] = val.detach().cpu().numpy() if isinstance(val, torch.Tensor) else np.asarray(val)
    for oname, model in grads:
        params = [p.unconstrained_tensor for p in model.parameters]
        if hasattr(model, 'feature'):
            # the reference keeps the inducing inputs Z out of `model.parameters`
            # (svgp.py:91, sgpr.py:116) although Z is a trainable variable (features.py:65);
            # its gradient is part of the hot path, so it is appended as the last entry.
            params.append(model.feature._Z.unconstrained_tensor)
        obj = model.objective
        gs = torch.autograd.grad(obj, params, allow_unused=True)
        for i, (p, g) in enumerate(zip(params, gs)):
            res['param/%s/%d' % (oname, i)] = p.detach().cpu().numpy()
            res['grad/%s/%d' % (oname, i)] = (torch.zeros_like(p) if g is None else g) \
                .detach().cpu().numpy()
    return res
