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


@pytest.fixture
def gpf_raw(monkeypatch):
    """The package with only the library-touching entry points replaced: its hand-written
    autograd adjoints (ops.py) are the code under test."""
    import gpflowSlim
    from gpflowSlim._backend import ops
    cpu_ops_double.install(monkeypatch, level='raw')
    # the composed adjoint formulas of ops.py are the code under test here; the one-call library
    # adjoints (the default, FUSED_ADJOINTS) are covered by tests/test_library_on_cpu.py and on the GPU
    monkeypatch.setattr(ops, 'FUSED_ADJOINTS', [False])
    old = gpflowSlim.settings.device
    gpflowSlim.settings.device = 'cpu'
    yield gpflowSlim
    gpflowSlim.settings.device = None if old.type == 'cpu' else old


_ADJOINT_CASES = ['gpr_misc', 'gpr_composed', 'gpr_features', 'svgp_white_full', 'svgp_nonwhite_full',
                  'svgp_white_diag', 'svgp_nonwhite_diag', 'sgpr', 'sparse_bounds', 'svgp_multiclass',
                  'mc_models', 'functions']


@pytest.mark.parametrize('tri_aware', [False, True])
@pytest.mark.parametrize('name', _ADJOINT_CASES)
def test_handwritten_adjoints_match_reference_golden(gpf_raw, golden, name, tri_aware, monkeypatch):
    """Same golden comparison with the REAL autograd Functions of _backend/ops.py (Cholesky
    adjoint via U = L^-T and triangular-aware GEMMs, TRSM adjoint, L^-T adjoint, NT-matmul
    adjoint) running over CPU stand-ins of the raw kernels."""
    from gpflowSlim._backend import ops
    monkeypatch.setattr(ops, 'TRI_AWARE_ADJOINTS', [tri_aware])     # experimental triangular-aware adjoints
    gold = golden(name)
    res = cases.run_case(gpf_raw, name, conv)
    for key in sorted(gold):
        tol = 1e-14 if key.startswith('param/') else RTOL
        e = relerr(res[key], gold[key])
        assert e < tol, '%s/%s: relative error %.3e' % (name, key, e)


def test_tri_aware_matmul_adjoint_against_torch(gpf_raw, monkeypatch):
    """Every (a_tri, b_tri) combination of the triangular-aware NT product, both adjoint
    implementations, against torch autograd on explicitly masked operands."""
    from gpflowSlim._backend import ops
    rng = np.random.default_rng(3)
    n, m = 37, 53
    for tri_aware in (False, True):
        monkeypatch.setattr(ops, 'TRI_AWARE_ADJOINTS', [tri_aware])
        for a_tri in (0, 1, 2):
            for b_tri in (0, 1, 2):
                # a triangular operand is square; its partner may be rectangular
                ra = n if a_tri else m
                A0 = conv(rng.standard_normal((ra, n)))
                B0 = conv(rng.standard_normal((n if b_tri else 41, n)))
                mask = lambda T, t: torch.tril(T) if t == 1 else (torch.triu(T) if t == 2 else T)
                A0, B0 = mask(A0, a_tri), mask(B0, b_tri)
                W = conv(rng.standard_normal((A0.shape[0], B0.shape[0])))
                A = A0.clone().requires_grad_(True)
                B = B0.clone().requires_grad_(True)
                val = (ops.matmul_nt(A, B, a_tri=a_tri, b_tri=b_tri) * W).sum()
                gA, gB = torch.autograd.grad(val, [A, B])
                At = A0.clone().requires_grad_(True)
                Bt = B0.clone().requires_grad_(True)
                valt = ((mask(At, a_tri) @ mask(Bt, b_tri).t()) * W).sum()
                wA, wB = torch.autograd.grad(valt, [At, Bt])
                assert relerr(val.detach(), valt.detach()) < 1e-13
                assert relerr(gA, wA) < 1e-12, (tri_aware, a_tri, b_tri, 'dA')
                assert relerr(gB, wB) < 1e-12, (tri_aware, a_tri, b_tri, 'dB')



def test_sgpr_with_multiscale_feature_uses_feature_kuf(gpf):
    """SGPR / GPRFITC / upper bound must build the cross-covariance through the feature
    (reference models/sgpr.py:132 `self.feature.Kuf(self.kern, self.X)`), not through kern.K --
    with a Multiscale feature (features.py:89-150) the two differ.  Checked against the reference
    formula (sgpr.py:121-156) written out with torch."""
    rng = np.random.default_rng(5)
    N, M, D = 60, 7, 3
    X, Y = conv(rng.standard_normal((N, D))), conv(rng.standard_normal((N, 1)))
    Z = rng.standard_normal((M, D))
    scales = 0.3 + rng.random((M, D))
    kern = gpf.kernels.RBF(D, ARD=True, lengthscales=1.3)
    feat = gpf.features.Multiscale(Z, scales)
    m = gpf.models.SGPR(X, Y, kern, feat=feat, obs_var=0.2)
    got = m.likelihood_tensor
    with torch.no_grad():
        Kuf = feat.Kuf(kern, X)
        Kuu = feat.Kuu(kern, jitter=gpf.settings.numerics.jitter_level)
        var = m.likelihood.variance
        L = torch.linalg.cholesky(Kuu)
        A = torch.linalg.solve_triangular(L, Kuf, upper=False) / torch.sqrt(var)
        B = A @ A.t() + torch.eye(M, dtype=torch.float64)
        LB = torch.linalg.cholesky(B)
        c = torch.linalg.solve_triangular(LB, A @ Y, upper=False) / torch.sqrt(var)
        want = (-0.5 * N * np.log(2 * np.pi) - torch.log(torch.diagonal(LB)).sum() - 0.5 * N * torch.log(var)
                - 0.5 * (Y ** 2).sum() / var + 0.5 * (c ** 2).sum() - 0.5 * kern.Kdiag(X).sum() / var
                + 0.5 * torch.diagonal(A @ A.t()).sum())
    assert relerr(got.detach(), want) < 1e-10
    # the feature's own parameters (Z and scales) are trained (reference: tf.trainable_variables())
    tt = m.trainable_tensors
    assert any(t is feat._scales.unconstrained_tensor for t in tt)
    assert any(t is feat._Z.unconstrained_tensor for t in tt)
    grads = torch.autograd.grad(m.objective, tt)
    assert all(torch.isfinite(g).all() for g in grads)
    # FITC and the upper bound run through the same feature path
    f = gpf.models.GPRFITC(X, Y, kern, feat=feat, obs_var=0.2)
    assert torch.isfinite(f.likelihood_tensor) and torch.isfinite(m.compute_upper_bound())


@pytest.mark.parametrize('comb', ['sum', 'product'])
def test_wide_input_combination_with_input_gradient_does_not_recurse(gpf, comb):
    """a fusable Sum / Product pushed onto the composed path because d/dX is wanted on
    inputs wider than GPS_MAX_DIMS used to rebuild itself for ever in Combination._split."""
    rng = np.random.default_rng(6)
    X0 = conv(rng.standard_normal((9, 40)))
    k1 = gpf.kernels.RBF(20, active_dims=list(range(20)), lengthscales=2.0)
    k2 = gpf.kernels.Linear(20, active_dims=list(range(20, 40)))
    kern = (k1 + k2) if comb == 'sum' else (k1 * k2)
    X = X0.clone().requires_grad_(True)
    K = kern.K(X)
    (g,) = torch.autograd.grad(K.sum(), [X])
    with torch.no_grad():
        want = kern.K(X0)               # no input gradient wanted: the fused path
    assert relerr(K.detach(), want) < 1e-12
    assert torch.isfinite(g).all()
    Kd = kern.Kdiag(X)
    assert relerr(Kd.detach(), torch.diagonal(want)) < 1e-12
