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
    cpu_ops_double.install(monkeypatch, level='raw')
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

