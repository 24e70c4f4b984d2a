"""Host-side logic (no GPU): transforms / parameters against the oracle and the golden
initial states, kernel-expression compilation, settings."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle import ref_torch as R


def gpf():
    import gpflowSlim
    return gpflowSlim


def test_settings_surface():
    s = gpf().settings
    assert s.float_type is np.float64 and s.dtypes.float_type is np.float64
    assert s.numerics.jitter_level == 1e-6 and s.jitter == 1e-6
    s.set_jitter(1e-5)
    assert s.jitter == 1e-5
    s.set_jitter(1e-6)
    tmp = s.get_settings()
    tmp.numerics.jitter_level = 1e-3
    with s.temp_settings(tmp):
        assert s.numerics.jitter_level == 1e-3
    assert s.numerics.jitter_level == 1e-6


def test_log1pe_matches_oracle_and_is_exact_for_large_inputs():
    t = gpf().transforms.Log1pe()
    y = np.array([1e-5, 0.3, 1.0, 25.0, 900.0])
    raw = t.backward(y)
    np.testing.assert_allclose(raw, R.softplus_inv(y), rtol=0, atol=0)
    np.testing.assert_allclose(t.forward_tensor(torch.tensor(raw)).numpy(), y, rtol=1e-14)
    np.testing.assert_allclose(t.forward(raw), y, rtol=1e-14)
    # torch's softplus linearises above 20; ours (like tf.nn.softplus) must not
    x = torch.tensor([21.0], dtype=torch.float64)
    assert float(t.forward_tensor(x)) == pytest.approx(21.0 + np.log1p(np.exp(-21.0)) + 1e-6, rel=1e-15)


def test_lower_triangular_transform_matches_oracle():
    T = gpf().transforms.LowerTriangular(4, num_matrices=3)
    rng = np.random.default_rng(0)
    mats = np.stack([np.tril(rng.standard_normal((4, 4))) for _ in range(3)], 2)
    free = T.backward(mats)
    assert free.shape == (3, 10)
    np.testing.assert_array_equal(T.forward(free), mats)
    np.testing.assert_array_equal(T.forward_tensor(torch.tensor(free)).numpy(), mats)
    np.testing.assert_array_equal(R.vec_to_tri(free, 4).numpy(), mats)


def test_parameter_order_and_initial_state_match_reference(golden):
    """Same construction code as the reference => same parameter list, shapes and values."""
    g = gpf()
    X, Y = cases.synth_gpr(50, 4)
    m = g.models.GPR(X, Y, kern=g.kernels.RBF(4, ARD=True))
    gold = golden('gpr_c1')
    assert len(m.parameters) == 3
    for i, p in enumerate(m.parameters):
        np.testing.assert_allclose(p.unconstrained_tensor.detach().cpu().numpy(),
                                   gold['param/objective/%d' % i], rtol=1e-15)
    Xs, Ys, Z = cases.synth_svgp(100, 3, 7)
    sv = g.models.SVGP(Xs, Ys, g.kernels.RBF(3), g.likelihoods.Gaussian(), Z=Z, num_latent=2)
    assert [tuple(p.shape) for p in sv.parameters] == [(), (), (), (7, 2), (2, 28)]
    assert tuple(sv.q_sqrt.shape) == (7, 7, 2)
    np.testing.assert_array_equal(sv.q_sqrt[:, :, 1].detach().cpu().numpy(), np.eye(7))
    assert len(sv.trainable_tensors) == 6


def test_kernel_expression_compiles_to_descriptor():
    g = gpf()
    from gpflowSlim._backend import lib
    k = g.kernels
    kern = k.RBF(2, active_dims=[0, 1]) * k.Linear(1, active_dims=[2]) + k.Matern52(3) + 0.5
    prog = kern.program()
    d = prog.desc
    assert (d.n_prims, d.n_ops, d.n_theta) == (3, 4, 2 + 1 + 2 + 1)
    assert [d.prims[i].type for i in range(3)] == [lib.GPS_RBF, lib.GPS_LINEAR, lib.GPS_MATERN52]
    assert list(d.prims[1].dims[:1]) == [2] and list(d.prims[0].dims[:2]) == [0, 1]
    ops = [(d.ops[i].op, d.ops[i].dst, d.ops[i].a, d.ops[i].b) for i in range(4)]
    assert ops == [(lib.GPS_OP_MUL, 3, 0, 1), (lib.GPS_OP_CONST, 4, 5, 0), (lib.GPS_OP_ADD, 5, 3, 2),
                   (lib.GPS_OP_ADD, 6, 5, 4)]
    assert d.out_slot == 6
    np.testing.assert_allclose(prog.theta('cpu').detach().numpy(), [1, 1, 1, 1, 1, 0.5], rtol=1e-12)


def test_nkn_compiles_and_matches_reference_weights(golden):
    g = gpf()
    from gpflowSlim._backend import lib
    kern = cases.nkn_c3_kernel(g, 3)
    gold = golden('nkn')
    # identical numpy-RNG draw as the reference (wrapper.py:100-104)
    np.testing.assert_allclose(kern.parameters[0].unconstrained_tensor.detach().cpu().numpy(),
                               gold['param/objective/0'], rtol=1e-15)
    d = kern.program().desc
    assert d.n_prims == 6 and d.n_ops == 5
    assert [d.ops[i].op for i in range(5)] == [lib.GPS_OP_LINEAR, lib.GPS_OP_PRODUCT, lib.GPS_OP_LINEAR,
                                                lib.GPS_OP_PRODUCT, lib.GPS_OP_LINEAR]
    assert [d.ops[i].dst for i in range(5)] == [6, 14, 18, 22, 24] and d.out_slot == 24
    assert d.n_theta == (48 + 8 + 16 + 4 + 2 + 1) + (4 + 4 + 3 + 3 + 3 + 3)


def test_adam_matches_oracle_adam():
    g = gpf()
    p = torch.tensor([0.3, -1.2], dtype=torch.float64, requires_grad=True)
    q = p.detach().clone()
    opt = g.training.AdamOptimizer(1e-2)
    state = {}
    for step in range(3):
        gr = torch.tensor([0.5 * (step + 1), -2.0], dtype=torch.float64)
        opt.apply_gradients([(gr, p)])
        q = R.tf_adam_step([q], [gr], state, lr=1e-2)[0]
    np.testing.assert_allclose(p.detach().numpy(), q.numpy(), rtol=1e-15)


def test_likelihoods_match_reference_golden(golden):
    """the likelihood zoo is elementwise torch maths with no library call, so
    it can be held to the reference's golden vectors on CPU tensors as well (the GPU parity run
    repeats it on the device)."""
    import gpflowSlim as gpf
    from oracle import cases
    gold = golden('likelihoods')
    res = cases.run_case(gpf, 'likelihoods', lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64)))
    assert set(res) == set(gold)
    for key in sorted(gold):
        a, b = np.asarray(res[key], dtype=np.float64), np.asarray(gold[key], dtype=np.float64)
        assert a.shape == b.shape, key
        err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
        assert err < 1e-10, '%s: relative error %.2e' % (key, err)


def test_expressions_too_large_for_the_fused_kernel_fall_back_to_composed(monkeypatch):
    """More than 32 active dimensions in a primitive, more than 16 primitives or more than 160
    features per point do not fit the fused Gram kernel's tables: `fusable` turns False (no
    exception) and K / Kdiag take the composed route -- checked here against the oracle on the CPU
    test double, with the fused `gram` entry point booby-trapped for the oversized parts."""
    import cpu_ops_double
    g = gpf()
    k = g.kernels
    ops = cpu_ops_double.install(monkeypatch)
    g.settings.device = 'cpu'
    rng = np.random.default_rng(1)
    wide = k.RBF(40, ARD=True, lengthscales=4.0)
    assert not wide.fusable and k.RBF(32, ARD=True).fusable
    many = k.Sum([k.RBF(2, active_dims=[i % 5, (i + 1) % 5], lengthscales=1.0 + 0.1 * i) for i in range(17)])
    assert not many.fusable and all(c.fusable for c in many.kern_list)
    fat = k.Sum([k.Periodic(30, period=1.0 + 0.1 * i) for i in range(2)])          # 2 x 90 features
    assert not fat.fusable
    X40, X5, X30 = (torch.tensor(rng.standard_normal((23, d)) * 0.7) for d in (40, 5, 30))

    real_gram = ops.gram

    def trapped(prog, X, X2=None, diag_add=0.0):
        assert prog.desc.n_prims <= 16 and max(prog.desc.prims[i].ndims for i in range(prog.desc.n_prims)) <= 32
        return real_gram(prog, X, X2, diag_add)
    monkeypatch.setattr(ops, 'gram', trapped)
    spec = dict(type='rbf', variance=torch.tensor(1.0, dtype=torch.float64),
                lengthscales=torch.full((40,), 4.0, dtype=torch.float64))
    np.testing.assert_allclose(wide.K(X40).detach().numpy(), R.K(spec, X40).numpy(), rtol=1e-12)
    np.testing.assert_allclose(wide.K_jittered(X40, 1e-6).detach().numpy(),
                               (R.K(spec, X40) + 1e-6 * torch.eye(23)).numpy(), rtol=1e-12)
    want = sum(R.K(dict(type='rbf', variance=torch.tensor(1.0), lengthscales=torch.tensor(1.0 + 0.1 * i, dtype=torch.float64),
                        active_dims=[i % 5, (i + 1) % 5]), X5) for i in range(17))
    np.testing.assert_allclose(many.K(X5).detach().numpy(), want.numpy(), rtol=1e-12)
    np.testing.assert_allclose(many.Kdiag(X5).detach().numpy(), np.full(23, 17.0), rtol=1e-12)
    want = sum(R.K(dict(type='periodic', variance=torch.tensor(1.0), lengthscales=torch.tensor(1.0),
                        period=torch.tensor(1.0 + 0.1 * i, dtype=torch.float64)), X30) for i in range(2))
    np.testing.assert_allclose(fat.K(X30).detach().numpy(), want.numpy(), rtol=1e-10)
    # a gradient w.r.t. inputs wider than 32 columns also leaves the fused kernel, even if the
    # kernel itself only looks at two of them
    narrow = k.RBF(2, active_dims=[3, 38])
    assert narrow.fusable
    Xg = X40.clone().requires_grad_(True)
    assert not narrow._use_fused(Xg) and narrow._use_fused(X40)
    g.settings.device = None


def test_white_and_bias_terms_keep_the_gpr_on_the_fused_path(monkeypatch):
    """`RBF + White + Bias`: the Bias compiles to a constant op, the White variance is folded into
    the noise of the one-call objective -- ONE gpr_loglik call, no Cholesky op; a covariance made of
    White terms only, or one containing a composed kernel, takes the op-by-op path."""
    import cpu_ops_double
    g = gpf()
    k = g.kernels
    ops = cpu_ops_double.install(monkeypatch)
    g.settings.device = 'cpu'
    calls = {'fused': 0, 'chol': 0}
    real_ll, real_chol = ops.gpr_loglik, ops.cholesky

    def ll(prog, X, Yc, noise):
        calls['fused'] += 1
        assert prog.desc.n_prims == 1 and prog.desc.n_ops == 2          # RBF, CONST, ADD
        return real_ll(prog, X, Yc, noise)

    def chol(K):
        calls['chol'] += 1
        return real_chol(K)
    monkeypatch.setattr(ops, 'gpr_loglik', ll)
    monkeypatch.setattr(ops, 'cholesky', chol)
    X, Y = cases.synth_gpr(40, 2, seed=3)
    kern = k.RBF(2, ARD=True) + k.White(2, variance=0.05) + k.Bias(2, variance=0.3)
    m = g.models.GPR(X, Y, kern=kern)
    obj = m.objective
    grads = torch.autograd.grad(obj, [p.unconstrained_tensor for p in m.parameters])
    assert calls == {'fused': 1, 'chol': 0}
    white_grad, noise_grad = grads[2], grads[4]          # parameters: rbf var, ls, white var, bias var, noise
    # d/d(white variance) and d/d(noise variance) are the same trace term, times each softplus slope
    sig = lambda p: torch.sigmoid(p.unconstrained_tensor.detach())
    np.testing.assert_allclose(float(white_grad / sig(m.parameters[2])), float(noise_grad / sig(m.parameters[4])),
                               rtol=1e-12)
    # same numbers as the op-by-op evaluation
    m2 = g.models.GPR(X, Y, kern=k.RBF(2, ARD=True) + k.White(2, variance=0.05) + k.Bias(2, variance=0.3), fused=False)
    np.testing.assert_allclose(float(obj), float(m2.objective), rtol=1e-12)
    assert calls['chol'] == 1
    for kern2 in (k.White(2), k.RBF(2) + k.RatQuad(2) + k.White(2)):
        calls.update(fused=0, chol=0)
        g.models.GPR(X, Y, kern=kern2).objective
        assert calls == {'fused': 0, 'chol': 1}
    g.settings.device = None


def test_rescale_follows_reference_semantics():
    """Rescale is plain y = factor * x (reference transforms.py:215-251) and
    positiveRescale(s) = Chain(Rescale(s), positive): y = s * (softplus(x) + 1e-6), with
    log|dy/dx| = N log s + sum log sigmoid(x) (transforms.py:110-112, 244-248, 380-392)."""
    import gpflowSlim as gpf
    T = gpf.transforms
    s = 3.5
    t = T.positiveRescale(s)
    x = np.array([-2.0, 0.3, 4.0])
    y = t.forward(x)
    np.testing.assert_allclose(y, s * (np.log1p(np.exp(x)) + 1e-6), rtol=1e-15)
    np.testing.assert_allclose(t.backward(y), x, rtol=1e-12)
    xt = torch.tensor(x)
    np.testing.assert_allclose(t.forward_tensor(xt).numpy(), y, rtol=1e-15)
    lj = float(t.log_jacobian_tensor(xt))
    want = 3 * np.log(s) + np.log(1.0 / (1.0 + np.exp(-x))).sum()
    assert abs(lj - want) < 1e-12
    r = T.Rescale(s)
    np.testing.assert_allclose(r.forward(x), s * x)
    np.testing.assert_allclose(r.backward(s * x), x)
    assert str(r) == '%s*' % s and str(t).startswith('%s* ' % s)
    # a Parameter on the chained transform round-trips its value
    p = gpf.Param(2.0, transform=T.positiveRescale(s))
    assert abs(float(p.value) - 2.0) < 1e-12
