"""Kernel-level GPU tests through the C ABI: the DMMA GEMM against its plain-FMA twin and
torch fp64, the blocked Cholesky / TRSM / triangular inverse against LAPACK (torch), the
fused Gram forward/backward against the oracle's autograd, edge shapes (1, 127, 128, 129,
ragged, odd leading dimensions), and size-independent properties at larger N."""
import numpy as np
import pytest
import torch

from util import assert_close, conv, dev, relerr

pytestmark = pytest.mark.gpu


def _ops():
    from gpflowSlim._backend import ops
    return ops


def _spd(n, seed=0, cond_shift=None):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n + 3))
    S = A @ A.T / (n + 3) + (cond_shift if cond_shift is not None else 0.5) * np.eye(n)
    return S


@pytest.mark.parametrize('m,n,k', [(1, 1, 1), (5, 3, 2), (128, 128, 128), (129, 127, 65),
                                   (300, 200, 17), (257, 513, 384), (1000, 1, 1000),
                                   (130, 260, 16), (64, 64, 8), (200, 300, 50), (385, 129, 4098),
                                   (1000, 2, 1000), (127, 1, 2), (2500, 2100, 80)])
@pytest.mark.parametrize('impl', [0, 1, 2])
def test_gemm_nt_matches_torch(m, n, k, impl):
    ops = _ops()
    from gpflowSlim._backend.lib import handle_for
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    A, B, C = rng.standard_normal((m, k)), rng.standard_normal((n, k)), rng.standard_normal((m, n))
    h = handle_for(conv(A))
    h.set_option('gemm_impl', impl)
    try:
        out = ops.gemm_nt(conv(A), conv(B), alpha=0.7, beta=-0.3, out=conv(C))
    finally:
        h.set_option('gemm_impl', 0)
    assert_close(out, 0.7 * A @ B.T - 0.3 * C, 1e-13, 'gemm %dx%dx%d impl %d' % (m, n, k, impl))


def test_gemm_nt_unaligned_views():
    """Odd leading dimensions / 8-byte aligned bases take the 8-byte cp.async path (the TMA
    kernel needs 16-byte aligned operands with even leading dimensions)."""
    ops = _ops()
    rng = np.random.default_rng(3)
    Abig, Bbig = rng.standard_normal((150, 141)), rng.standard_normal((140, 141))
    A, B = conv(Abig)[3:140, 1:132], conv(Bbig)[5:133, 1:132]
    out = ops.gemm_nt(A, B)
    assert_close(out, Abig[3:140, 1:132] @ Bbig[5:133, 1:132].T, 1e-13, 'unaligned gemm')


@pytest.mark.parametrize('impl', [0, 2])
@pytest.mark.parametrize('n', [260, 700, 2200])
def test_gemm_nt_triangular_modes(n, impl):
    """impl 0: TMA + mbarrier kernel (even n: operands qualify); impl 2: cp.async kernel."""
    ops = _ops()
    from gpflowSlim._backend.lib import TRI_LOWER, TRI_UPPER, handle_for
    h = handle_for(dev())
    h.set_option('gemm_impl', impl)
    try:
        _triangular_modes(ops, n, TRI_LOWER, TRI_UPPER)
    finally:
        h.set_option('gemm_impl', 0)


def test_gemm_tma_submatrix_views_and_long_k():
    """TMA path on aligned sub-matrix views (tensor-map dims must clip at the VIEW's edge, not the
    parent's: out-of-range rows / k read as zero) and a K long enough to wrap the 6-stage ring
    many times with a ragged tail."""
    ops = _ops()
    rng = np.random.default_rng(11)
    Abig, Bbig = rng.standard_normal((400, 2100)), rng.standard_normal((300, 2100))
    for (r0, r1, c0, c1, s0, s1) in [(2, 259, 16, 2066, 4, 133), (0, 400, 0, 2100, 0, 300),
                                     (130, 131, 64, 66, 7, 8), (1, 129, 1024, 1040, 0, 128)]:
        A, B = conv(Abig)[r0:r1, c0:c1], conv(Bbig)[s0:s1, c0:c1]
        Cn = rng.standard_normal((r1 - r0, s1 - s0))
        out = ops.gemm_nt(A, B, alpha=-1.0, beta=1.0, out=conv(Cn))
        assert_close(out, Cn - Abig[r0:r1, c0:c1] @ Bbig[s0:s1, c0:c1].T, 1e-13, 'tma view gemm')


def _triangular_modes(ops, n, TRI_LOWER, TRI_UPPER):
    rng = np.random.default_rng(n)
    Lo, Up, G = np.tril(rng.standard_normal((n, n))), np.triu(rng.standard_normal((n, n))), \
        rng.standard_normal((n, n))
    assert_close(ops.gemm_nt(conv(Lo), conv(G), a_tri=TRI_LOWER), Lo @ G.T, 1e-13, 'a lower')
    assert_close(ops.gemm_nt(conv(Up), conv(G), a_tri=TRI_UPPER), Up @ G.T, 1e-13, 'a upper')
    assert_close(ops.gemm_nt(conv(G), conv(Lo), b_tri=TRI_LOWER), G @ Lo.T, 1e-13, 'b lower')
    assert_close(ops.gemm_nt(conv(Up), conv(Up), a_tri=TRI_UPPER, b_tri=TRI_UPPER), Up @ Up.T, 1e-13, 'UU^T')
    low = ops.gemm_nt(conv(Up), conv(Up), a_tri=TRI_UPPER, b_tri=TRI_UPPER, c_uplo=1)
    assert_close(low, np.tril(Up @ Up.T), 1e-13, 'UU^T lower only')


@pytest.mark.parametrize('leaf', [0, 1])
@pytest.mark.parametrize('n', [1, 2, 31, 127, 128, 129, 255, 256, 300, 641, 1500])
def test_potrf_trsm_inverse(n, leaf):
    """leaf=0: blocked DMMA leaf kernel (production); leaf=1: simple check kernel."""
    ops = _ops()
    from gpflowSlim._backend.lib import handle_for
    S = _spd(n, seed=n)
    h = handle_for(conv(S))
    h.set_option('leaf_impl', leaf)
    try:
        _check_potrf_trsm_inverse(ops, S, n)
    finally:
        h.set_option('leaf_impl', 0)


def _check_potrf_trsm_inverse(ops, S, n):
    L = ops.potrf(conv(S))
    Lref = np.linalg.cholesky(S)
    assert_close(L, Lref, 1e-12, 'potrf n=%d' % n)
    assert float(torch.triu(L, 1).abs().max()) == 0.0 if n > 1 else True
    rng = np.random.default_rng(n + 1)
    for m in (1, 70, 200):
        B = rng.standard_normal((m, n))
        X = ops.trsm_rlt_(L, conv(B).clone())
        assert_close(X, np.linalg.solve(Lref, B.T).T, 1e-11, 'trsm n=%d m=%d' % (n, m))
    U = ops.tri_inv_t(L)
    assert_close(U, np.linalg.inv(Lref).T, 1e-11, 'tri_inv_t n=%d' % n)


def test_potrf_large_property():
    """N = 4096 (config-2 scale / 2): L L^T = K and K^-1 K = I to LAPACK-level residuals."""
    ops = _ops()
    n = 4096
    g = torch.Generator(device='cpu').manual_seed(0)
    A = torch.randn(n, n + 8, generator=g, dtype=torch.float64).to(dev())
    S = A @ A.t() / n + 0.1 * torch.eye(n, dtype=torch.float64, device=dev())
    L = ops.potrf(S)
    res = (L @ L.t() - S).abs().max() / S.abs().max()
    assert float(res) < 1e-13, float(res)
    Lt = torch.linalg.cholesky(S)
    assert_close(L, Lt, 1e-11, 'potrf vs cuSOLVER')
    U = ops.tri_inv_t(L)
    Kinv = ops.gemm_nt(U, U, a_tri=2, b_tri=2)
    eye = Kinv @ S
    assert float((eye - torch.eye(n, dtype=torch.float64, device=dev())).abs().max()) < 1e-9


def test_autograd_ops_match_torch():
    """Cholesky / TRSM / triangular-aware matmul adjoints against torch's own autograd."""
    ops = _ops()
    n, m = 203, 77
    S0 = conv(_spd(n, 5))
    B0 = conv(np.random.default_rng(6).standard_normal((m, n)))
    W1 = conv(np.random.default_rng(7).standard_normal((m, n)))

    def run(mine):
        S = S0.clone().requires_grad_(True)
        B = B0.clone().requires_grad_(True)
        if mine:
            L = ops.cholesky(S)
            X = ops.trsm_rlt(B, L)
            Y = ops.matmul_nt(X, X)
            Z = ops.solve_upper_t(L, ops.t(X))
        else:
            L = torch.linalg.cholesky(S)
            X = torch.linalg.solve_triangular(L, B.t(), upper=False).t()
            Y = X @ X.t()
            Z = torch.linalg.solve_triangular(L.t(), X.t(), upper=True)
        val = (X * W1).sum() + torch.log(torch.diagonal(L)).sum() + (Y ** 2).sum() * 1e-3 + (Z * W1.t()).sum()
        gS, gB = torch.autograd.grad(val, [S, B])
        return val, 0.5 * (gS + gS.t()), gB
    a, b = run(True), run(False)
    assert_close(a[0], b[0], 1e-12, 'value')
    assert_close(a[1], b[1], 1e-10, 'dS')
    assert_close(a[2], b[2], 1e-10, 'dB')


def _zoo():
    import gpflowSlim as gpf
    from oracle import cases
    return cases._kernel_zoo(gpf, 3)


def test_gram_backward_against_oracle_autograd():
    """dtheta, dX and dX2 of every primitive / composition vs torch autograd through the oracle."""
    from oracle import ref_torch as R
    from test_oracle_golden import _zoo_specs
    import gpflowSlim as gpf
    rng = np.random.default_rng(21)
    Xn, X2n = rng.standard_normal((75, 3)) * 1.2, rng.standard_normal((41, 3)) * 1.2
    Wn, Wsn = rng.standard_normal((75, 41)), rng.standard_normal((75, 75))
    specs = _zoo_specs(3)
    for name, make in _zoo():
        kern = make()
        # ----- product under test
        X = conv(Xn).requires_grad_(True)
        X2 = conv(X2n).requires_grad_(True)
        val = (kern.K(X, X2) * conv(Wn)).sum() + (kern.K(X) * conv(Wsn)).sum() \
            + (kern.Kdiag(X) * conv(Wsn[:, 0])).sum()
        params = [p.unconstrained_tensor for p in kern.parameters]
        g = torch.autograd.grad(val, params + [X, X2])
        # ----- oracle with leaves at the constrained values
        spec, leaves = _leafify(specs[name])
        Xo = torch.tensor(Xn, requires_grad=True)
        X2o = torch.tensor(X2n, requires_grad=True)
        valo = (R.K(spec, Xo, X2o) * torch.tensor(Wn)).sum() + (R.K(spec, Xo) * torch.tensor(Wsn)).sum() \
            + (R.Kdiag(spec, Xo) * torch.tensor(Wsn[:, 0])).sum()
        go = torch.autograd.grad(valo, leaves + [Xo, X2o], allow_unused=True)
        # Matern diagonals carry sqrt(d2 + 1e-12) with d2 ~ +-1e-16 rounding noise: 1e-9 is their floor
        assert_close(val, valo, 2e-9, name + ' value')
        assert_close(g[-2], go[-2], 1e-8, name + ' dX')
        assert_close(g[-1], go[-1], 1e-8, name + ' dX2')
        # chain rule to the unconstrained parameters: d softplus = sigmoid
        assert len(params) == len(leaves), name
        for p, gp, leaf, gl in zip(params, g, leaves, go):
            gl = torch.zeros_like(leaf) if gl is None else gl
            want = gl.reshape(p.shape) * torch.sigmoid(p.detach().cpu())
            assert_close(gp, want, 1e-8, name + ' dtheta')


def _leafify(spec):
    """Replace every tensor-valued parameter of an oracle spec by a leaf; returns (spec, leaves)
    in the reference's parameter order (variance, lengthscales, period; children in order)."""
    leaves = []

    def walk(s):
        if not isinstance(s, dict):
            return s
        s = dict(s)
        if s['type'] in ('sum', 'product'):
            s['children'] = [walk(c) for c in s['children']]
            return s
        for key in ('variance', 'lengthscales', 'period'):
            if key in s:
                leaf = s[key].detach().clone().to(torch.float64).requires_grad_(True)
                leaves.append(leaf)
                s[key] = leaf
        return s
    return walk(spec), leaves


def test_kernel_symmetry_psd_and_tiling_edges():
    import gpflowSlim as gpf
    for n in (1, 63, 64, 65, 130):
        X = conv(np.random.default_rng(n).standard_normal((n, 2)))
        k = gpf.kernels.Matern52(2, lengthscales=0.7) + gpf.kernels.Periodic(2, period=1.3) * gpf.kernels.RBF(2)
        K = k.K(X)
        assert float((K - K.t()).abs().max()) < 1e-14
        ev = torch.linalg.eigvalsh(K + 1e-9 * torch.eye(n, dtype=torch.float64, device=dev()))
        assert float(ev.min()) > -1e-9
        assert_close(k.K(X, X), K, 1e-14, 'K(X,X) == K(X)')
    assert k.K(X[:0], X).shape == (0, 130)


def test_wrong_inputs_raise():
    import gpflowSlim as gpf
    from gpflowSlim._backend import ops
    with pytest.raises(ValueError):
        ops.gemm_nt(conv(np.zeros((3, 4))), conv(np.zeros((3, 5))))
    with pytest.raises(ValueError):
        gpf.kernels.RBF(4).K(conv(np.zeros((5, 2))))          # active dims beyond X's columns
    with pytest.raises(RuntimeError):
        ops.gemm_nt(torch.zeros(2, 2, dtype=torch.float64), torch.zeros(2, 2, dtype=torch.float64))


@pytest.mark.parametrize('cls', ['RBF', 'Matern12', 'Matern32', 'Matern52', 'Exponential'])
@pytest.mark.parametrize('d,ard', [(1, False), (3, True), (8, True), (11, True), (16, True), (5, False)])
def test_gram_fast_path_equals_interpreter(cls, d, ard):
    """The register-tiled kernels for a single stationary covariance (gram_impl 0) against the
    generic interpreter kernels (gram_impl 1): Gram K(X), K(X, X2), the dense backward and the
    fused GPR objective + gradient, on ragged sizes that leave partial 64-tiles."""
    import gpflowSlim as gpf
    from gpflowSlim._backend.lib import handle_for
    rng = np.random.default_rng(100 * d + len(cls))
    n, m = 203, 131
    Xn, X2n = rng.standard_normal((n, d)), rng.standard_normal((m, d))
    Wn, Wsn = rng.standard_normal((n, m)), rng.standard_normal((n, n))
    Yn = rng.standard_normal((n, 2))
    ls = 0.8 + 0.1 * np.arange(d) if ard else 1.3
    h = handle_for(dev())
    out = {}
    for impl in (0, 1):
        h.set_option('gram_impl', impl)
        try:
            kern = getattr(gpf.kernels, cls)(d, variance=1.4, lengthscales=ls, ARD=ard)
            params = [p.unconstrained_tensor for p in kern.parameters]
            K, K2 = kern.K(conv(Xn)), kern.K(conv(Xn), conv(X2n))
            val = (K2 * conv(Wn)).sum() + (K * conv(Wsn)).sum()
            g = torch.autograd.grad(val, params)
            mdl = gpf.models.GPR(conv(Xn), conv(Yn), kern=kern)
            obj = mdl.objective
            go = torch.autograd.grad(obj, [p.unconstrained_tensor for p in mdl.parameters])
            # input gradients: both arguments, only the second one (the sparse models' K(Xb, Z)),
            # and the one-argument (symmetric) form -- with active dims that skip a column
            Xg, X2g = conv(Xn).requires_grad_(True), conv(X2n).requires_grad_(True)
            v2 = (kern.K(Xg, X2g) * conv(Wn)).sum() + (kern.K(Xg) * conv(Wsn)).sum()
            gx = torch.autograd.grad(v2, [Xg, X2g] + params)
            X2h = conv(X2n).requires_grad_(True)
            gz = torch.autograd.grad((kern.K(conv(Xn), X2h) * conv(Wn)).sum(), [X2h] + params)
            out[impl] = [K.detach(), K2.detach(), val.detach(), obj.detach()] + \
                [t.detach() for t in g + go + gx + gz]
        finally:
            h.set_option('gram_impl', 0)
    for i, (a, b) in enumerate(zip(out[0], out[1])):
        # Gram entries: same arithmetic; sums (values, gradients): different summation order
        assert_close(a, b, 1e-13 if i < 2 else 1e-10, '%s d=%d fast vs interpreter [%d]' % (cls, d, i))
