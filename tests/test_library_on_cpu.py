"""libgpslim_b200 -- the C++ host orchestration AND the CUDA kernels' source -- built for the CPU
and driven through the package's own ctypes binding and autograd layer.

`tests/emu/cpu_build/build.py` transforms the five unmodified .cu files textually (kernel launches
-> EMU_LAUNCH, dynamic shared memory, one cache-hint asm, the six PTX helper functions of the
TMA/mbarrier GEMM kernel -> functional stand-ins (its main loop is the kernel's own), DLPack device type) and compiles them with g++
against a stand-in <cuda_runtime.h>: one host thread per warp with its lanes as cooperative fibers
(tests/emu/emu_fibers.h), barriers for __syncthreads / warp collectives, an emulation of
mma.sync.m8n8k4.f64.  The resulting shared
library exports the same C ABI; here it replaces the GPU library under `gpflowSlim._backend.lib`,
so everything above it is the shipped code: DLTensor marshaling, argument checks, workspace
management, the recursive blocked Cholesky / triangular inverse, the fused one-call GPR
objective + gradient, the Gram kernels, the autograd adjoints.

What this cannot see: that the hardware's 128-byte swizzle and mbarrier semantics ARE what the stand-ins
implement (the GPU tests do), stream
concurrency (launches run to completion in issue order), and anything about speed.

The whole selection runs in about a minute; GPSLIM_CPU_LIB_FULL=0 keeps only the short one."""
import os

import numpy as np
import pytest
import torch

from oracle import cases

FULL = os.environ.get('GPSLIM_CPU_LIB_FULL', '1') == '1'      # GPSLIM_CPU_LIB_FULL=0: the short selection


def conv(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=torch.float64)


@pytest.fixture(scope='module')
def cpu_lib(tmp_path_factory):
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location('cpu_build', os.path.join(here, 'emu', 'cpu_build', 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(str(tmp_path_factory.mktemp('cpu_lib')))


@pytest.fixture
def gpf(cpu_lib, monkeypatch):
    """The package bound to the CPU build of the library."""
    import gpflowSlim
    from gpflowSlim._backend import lib, ops
    monkeypatch.setattr(lib, 'LIB_PATH', cpu_lib)
    monkeypatch.setattr(lib, '_lib', None)

    class CpuHandle(lib.Handle):
        def sync_stream(self):          # no streams on the CPU: launches run at their call site
            pass
    holder = []

    def handle_for(_):
        if not holder:
            holder.append(CpuHandle(0))
        return holder[0]
    monkeypatch.setattr(lib, 'handle_for', handle_for)
    monkeypatch.setattr(ops, 'handle_for', handle_for)
    ops._U_CACHE.clear()
    old = gpflowSlim.settings.device
    gpflowSlim.settings.device = 'cpu'
    yield gpflowSlim
    gpflowSlim.settings.device = None if old.type == 'cpu' else old
    ops._U_CACHE.clear()


def test_abi_of_the_cpu_build_matches_the_header(cpu_lib):
    import ctypes
    from gpflowSlim._backend import lib
    cdll = ctypes.CDLL(cpu_lib)
    for name in lib.SIGNATURES:
        assert hasattr(cdll, name), name
    assert cdll.gps_version() == 100


@pytest.mark.parametrize('n', [1, 31, 129, 200, 256])    # 256: the level-batched triangular inverse
def test_cholesky_solve_inverse_through_the_host_recursion(gpf, n):
    """gps_potrf (recursive blocked, leaves + strip TRSMs + lower-masked GEMM updates),
    gps_trsm_rlt, gps_tri_inv_t, gps_sum_log_diag, gps_row_sumsq, gps_transpose."""
    from gpflowSlim._backend import ops
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n + 3))
    S = conv(A @ A.T / (n + 3) + 0.5 * np.eye(n))
    L = ops.potrf(S)
    ref = torch.linalg.cholesky(S)
    assert float((L - ref).abs().max()) < 1e-12 * float(ref.abs().max())
    assert float(torch.triu(L, 1).abs().max()) == 0.0
    B = conv(rng.standard_normal((17, n)))
    X = ops.trsm_rlt_(L, B.clone())
    want = torch.linalg.solve_triangular(ref, B.t(), upper=False).t()
    assert float((X - want).abs().max()) < 1e-10 * max(1.0, float(want.abs().max()))
    U = ops.tri_inv_t(L)
    Ti = torch.linalg.inv(ref).t()
    assert float((U - Ti).abs().max()) < 1e-10 * float(Ti.abs().max())
    assert float((ops.row_sumsq(B) - (B ** 2).sum(1)).abs().max()) < 1e-12
    assert torch.equal(ops.transpose(B), B.t().contiguous())


def test_fuzz_factorisation_solves_and_inverses_on_padded_buffers(gpf):
    """Randomised differential test of the host recursion + kernels behind gps_potrf, gps_trsm_rlt,
    gps_tri_inv_t, gps_potri, gps_chol_bwd and gps_trsm_bwd against LAPACK / torch autograd: random orders
    1..520 (every leaf / strip / GEMM-tile raggedness), operands that are column slices of wider NaN-filled
    buffers (leading dimension != width: nothing outside the view may be read or written), split-K on and
    off.  GPSLIM_FUZZ=<n> cases (default 6; 150 were run clean when this was written)."""
    import ctypes
    from gpflowSlim._backend import lib, ops
    ncases = int(os.environ.get('GPSLIM_FUZZ', '6'))
    rng = np.random.default_rng(77)
    h = lib.handle_for(None)

    def padded(a, pad):
        buf = torch.full((a.shape[0], a.shape[1] + pad), float('nan'), dtype=torch.float64)
        v = buf[:, :a.shape[1]]
        v.copy_(a)
        return buf, v
    for it in range(ncases):
        n = int(rng.choice([int(rng.integers(1, 40)), int(rng.integers(100, 300)), int(rng.integers(300, 521)),
                            int(rng.choice([128, 256, 384, 512]))]))      # powers of two: the level-batched inverse
        m = int(rng.integers(1, 200))
        pad = int(rng.choice([0, 1, 2, 5, 16]))
        h.set_option('gemm_splitk', int(rng.integers(0, 2)))
        try:
            A = rng.standard_normal((n, n + 2))
            S = conv(A @ A.T / (n + 2) + 0.3 * np.eye(n))
            ref = torch.linalg.cholesky(S)
            bufL, L = padded(S, pad)
            info = ctypes.c_int(0)
            h.check(h.lib.gps_potrf(h.ptr, lib.view(L).ref, 1, ctypes.byref(info)))
            assert info.value == 0
            assert float((L - ref).abs().max()) < 1e-12 * float(ref.abs().max()), (it, n, pad)
            assert float(torch.triu(L, 1).abs().max()) == 0.0 and bool(torch.isnan(bufL[:, n:]).all())
            B = conv(rng.standard_normal((m, n)))
            bufB, Bv = padded(B, pad)
            h.check(h.lib.gps_trsm_rlt(h.ptr, lib.view(L).ref, lib.view(Bv).ref))
            want = torch.linalg.solve_triangular(ref, B.t(), upper=False).t()
            assert float((Bv - want).abs().max()) < 1e-10 * max(1.0, float(want.abs().max())), (it, n, m, pad)
            assert bool(torch.isnan(bufB[:, n:]).all())
            bufU, U = padded(torch.zeros(n, n, dtype=torch.float64), pad)
            U.fill_(float('nan'))
            h.check(h.lib.gps_tri_inv_t(h.ptr, lib.view(L).ref, lib.view(U).ref))
            Ti = torch.linalg.inv(ref).t()
            assert float((U - Ti).abs().max()) < 1e-10 * float(Ti.abs().max()), (it, n, pad)
            Kinv = ops.potri(L)
            Ki = torch.linalg.inv(S)
            assert float((torch.tril(Kinv) - torch.tril(Ki)).abs().max()) < 1e-9 * float(Ki.abs().max()), (it, n, pad)
            # adjoints against torch autograd
            Lbar = conv(np.tril(rng.standard_normal((n, n))))
            Sg = S.clone().requires_grad_(True)
            (torch.linalg.cholesky(Sg) * Lbar).sum().backward()
            Sbar = ops.chol_bwd(L, Lbar)
            wantS = 0.5 * (Sg.grad + Sg.grad.t())
            gotS = torch.tril(Sbar) + torch.tril(Sbar, -1).t()
            assert float((gotS - wantS).abs().max()) < 1e-9 * max(1.0, float(wantS.abs().max())), (it, n, pad)
            Xbar = conv(rng.standard_normal((m, n)))
            Lg, Bg = ref.clone().requires_grad_(True), B.clone().requires_grad_(True)
            (torch.linalg.solve_triangular(Lg, Bg.t(), upper=False).t() * Xbar).sum().backward()
            Bbar, Lb = ops.trsm_bwd(L, Bv, Xbar)
            assert float((Bbar - Bg.grad).abs().max()) < 1e-9 * max(1.0, float(Bg.grad.abs().max())), (it, n, m, pad)
            assert float((torch.tril(Lb) - torch.tril(Lg.grad)).abs().max()) < 1e-9 * max(1.0, float(Lg.grad.abs().max())), (it, n, m)
        finally:
            h.set_option('gemm_splitk', 1)


def test_not_positive_definite_is_reported_through_the_abi(gpf):
    from gpflowSlim._backend import ops
    S = torch.eye(150, dtype=torch.float64)
    S[140, 140] = -1.0
    with pytest.raises(gpf.CholeskyError, match='141'):
        ops.potrf(S)


def test_argument_checks_and_error_reporting_of_the_abi(gpf):
    """Bad shapes / dtypes / layouts come back as -(argument index) with a message from
    gps_last_error ("every function returns int status"), empty inputs are
    no-ops, and the library never touches what it is not given."""
    import ctypes
    from gpflowSlim._backend import lib, ops
    with pytest.raises(ValueError, match='K mismatch'):
        ops.gemm_nt(conv(np.zeros((3, 4))), conv(np.zeros((3, 5))))
    with pytest.raises(ValueError, match='column'):
        gpf.kernels.RBF(4).K(conv(np.zeros((5, 2))))               # active dims beyond X's columns
    with pytest.raises(TypeError):
        lib.view(torch.zeros(2, 2, dtype=torch.float32))
    h = lib.handle_for(None)
    A = conv(np.eye(4))
    va, vi = lib.view(A), lib.view(torch.zeros(4, dtype=torch.int64))
    rc = h.lib.gps_potrf(h.ptr, vi.ref, 1, None)                    # int64 where float64 is required
    assert rc == -2 and b'float64' in h.lib.gps_last_error(h.ptr)
    rc = h.lib.gps_transpose(h.ptr, va.ref, lib.view(conv(np.zeros((3, 4)))).ref)
    assert rc == -3 and b'shape' in h.lib.gps_last_error(h.ptr)
    with pytest.raises(ValueError, match='unknown option'):
        h.set_option('no_such_option', 1)
    assert h.lib.gps_potrf(None, va.ref, 1, None) == -1             # null handle
    # empty inputs
    X = conv(np.random.default_rng(0).standard_normal((7, 2)))
    k = gpf.kernels.Matern52(2) + gpf.kernels.Linear(2)
    assert k.K(X[:0], X).shape == (0, 7) and k.K(X, X[:0]).shape == (7, 0) and k.Kdiag(X[:0]).shape == (0,)
    assert ops.potrf(conv(np.zeros((0, 0)))).shape == (0, 0)
    L = ops.potrf(conv(np.eye(5) * 4.0))
    assert ops.trsm_rlt_(L, conv(np.zeros((0, 5)))).shape == (0, 5)
    assert ops.gemm_nt(conv(np.zeros((0, 3))), conv(np.ones((4, 3)))).shape == (0, 4)
    assert float(ops.gemm_nt(conv(np.zeros((2, 0))), conv(np.zeros((3, 0)))).abs().max()) == 0.0
    # a model asked to predict at no points at all (models/gpr.py:118-131), fused and op-by-op
    Xd, Yd = cases.synth_gpr(40, 2, seed=1)
    for fused in (True, False):
        m = gpf.models.GPR(conv(Xd), conv(Yd), kern=gpf.kernels.RBF(2), fused=fused)
        with torch.no_grad():
            mu, var = m.predict_f(conv(np.zeros((0, 2))))
        assert mu.shape == (0, 1) and var.shape == (0, 1)


def test_gemm_flags_through_the_launch_code(gpf):
    from gpflowSlim._backend import ops
    rng = np.random.default_rng(5)
    n = 200
    Up, G = np.triu(rng.standard_normal((n, n))), rng.standard_normal((140, n))
    close = lambda a, b: np.testing.assert_allclose(a.numpy(), b, rtol=0, atol=3e-13 * max(1.0, np.abs(b).max()))
    close(ops.gemm_nt(conv(G), conv(Up), b_tri=2), G @ Up.T)
    close(ops.gemm_nt(conv(Up), conv(Up), a_tri=2, b_tri=2, c_uplo=1), np.tril(Up @ Up.T))
    C0 = rng.standard_normal((140, 140))
    close(ops.gemm_nt(conv(G), conv(G), alpha=-0.5, beta=1.0, out=conv(C0.copy())), C0 - 0.5 * G @ G.T)
    A3 = conv(rng.standard_normal((33, 7)))[:, :5]                    # odd leading dimension: 8-byte staging
    close(ops.gemm_nt(A3, A3), A3.numpy() @ A3.numpy().T)


def test_tma_kernel_main_loop_against_the_cp_async_kernel_and_numpy(gpf):
    """gemm_nt_tma_kernel -- the kernel that does 95 % of a GPR step -- with its OWN main loop on the CPU
    (build.py swaps only its six PTX helpers for the functional stand-ins of emu/cpu_build/tma_emulation.inc):
    the 6-stage ring with more k-tiles than stages (barrier phases flip several times), fewer k-tiles than
    stages, K not a multiple of the 16-wide box (zero-filled edge), ragged M / N, triangular operands (K ranges
    per tile), lower-only output, alpha / beta, and the K-sliced launch of few-tile products -- against numpy
    and against the cp.async tensor-core kernel (gemm_impl 2), which the same launch code selects for
    operands TMA cannot address."""
    from gpflowSlim._backend import lib, ops
    rng = np.random.default_rng(21)
    h = lib.handle_for(None)
    shapes = [(150, 140, 400), (129, 260, 97), (5, 3, 2), (130, 131, 16), (300, 70, 1100), (257, 257, 40)]
    for (m, n, k) in shapes:
        A, B = rng.standard_normal((m, k)), rng.standard_normal((n, k))
        C0 = rng.standard_normal((m, n))
        variants = [dict(), dict(alpha=-0.7, beta=0.3)]
        if m == n:
            Lo, Up = np.tril(rng.standard_normal((m, m))), np.triu(rng.standard_normal((m, m)))
            variants += [dict(sq=(Lo, Lo, 1, 1, 1)), dict(sq=(Up, Up, 2, 2, 0)), dict(sq=(Lo, Up, 1, 2, 0))]
        for v in variants:
            out = {}
            for impl in (0, 2):
                h.set_option('gemm_impl', impl)
                try:
                    if 'sq' in v:
                        a, b, ta, tb, cu = v['sq']
                        got = ops.gemm_nt(conv(a), conv(b), a_tri=ta, b_tri=tb, c_uplo=cu)
                        want = np.tril(a @ b.T) if cu else a @ b.T
                    elif 'alpha' in v:
                        got = ops.gemm_nt(conv(A), conv(B), alpha=v['alpha'], beta=v['beta'], out=conv(C0.copy()))
                        want = v['alpha'] * (A @ B.T) + v['beta'] * C0
                    else:
                        got = ops.gemm_nt(conv(A), conv(B))
                        want = A @ B.T
                finally:
                    h.set_option('gemm_impl', 0)
                out[impl] = got
                assert np.abs(got.numpy() - want).max() < 1e-12 * max(1.0, np.abs(want).max()), (m, n, k, list(v), impl)
            # same tile order, K ranges and per-lane accumulation order in both kernels
            assert torch.equal(out[0], out[2]), (m, n, k, list(v))


def test_fuzz_gemm_through_the_shipped_launch_code(gpf):
    """gps_gemm_nt with random shapes (K up to 1300: the stage ring wraps many times, few-tile products are
    K-sliced), leading-dimension paddings and 8-byte misalignments (the launch code picks the TMA kernel, the
    16-byte or the 8-byte cp.async staging), triangular operands, lower output, alpha / beta, output views
    inside wider buffers whose padding columns must stay untouched -- against numpy.
    GPSLIM_FUZZ=<n> cases (default 25; 300 were run clean when this was written)."""
    from gpflowSlim._backend import lib
    ncases = int(os.environ.get('GPSLIM_FUZZ', '25'))
    rng = np.random.default_rng(404)
    h = lib.handle_for(None)

    def tri(a, t):
        return a if t == 0 else (np.tril(a) if t == 1 else np.triu(a))
    for it in range(ncases):
        a_tri = b_tri = c_uplo = 0
        if rng.random() < 0.5:
            M = N = Kd = int(rng.integers(1, 300))
            a_tri, b_tri, c_uplo = int(rng.integers(0, 3)), int(rng.integers(0, 3)), int(rng.integers(0, 2))
            if rng.random() < 0.3:
                Kd, a_tri, b_tri = int(rng.integers(1, 1300)), 0, 0
        else:
            M, N = int(rng.integers(1, 300)), int(rng.integers(1, 300))
            Kd = int(rng.choice([int(rng.integers(1, 70)), int(rng.integers(70, 1300))]))
        pa, pb, pc = (int(rng.integers(0, 4)) for _ in range(3))
        oa, ob = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        An = rng.standard_normal(M * (Kd + pa) + 1)[oa:oa + M * (Kd + pa)].reshape(M, Kd + pa)[:, :Kd]
        Bn = rng.standard_normal(N * (Kd + pb) + 1)[ob:ob + N * (Kd + pb)].reshape(N, Kd + pb)[:, :Kd]
        An[...] = tri(An.copy(), a_tri)
        Bn[...] = tri(Bn.copy(), b_tri)
        C0 = rng.standard_normal((M, N))
        beta, alpha = float(rng.choice([0.0, 1.0, -0.4])), float(rng.choice([1.0, -1.0, 0.6]))
        Cfull = np.full((M, N + pc), 777.0)
        Cfull[:, :N] = C0 if beta else np.nan
        A, B, Cf = torch.from_numpy(An), torch.from_numpy(Bn), torch.from_numpy(Cfull)     # share memory: same strides / alignment
        h.set_option('gemm_splitk', int(rng.random() < 0.7))
        try:
            h.check(h.lib.gps_gemm_nt(h.ptr, alpha, lib.view(A).ref, lib.view(B).ref, beta, lib.view(Cf[:, :N]).ref,
                                      a_tri, b_tri, c_uplo))
        finally:
            h.set_option('gemm_splitk', 1)
        got = Cfull[:, :N]
        want = alpha * An @ Bn.T + (beta * C0 if beta else 0.0)
        tol = 5e-13 * max(1.0, np.abs(want).max())
        what = dict(it=it, M=M, N=N, K=Kd, a_tri=a_tri, b_tri=b_tri, c_uplo=c_uplo, pad=(pa, pb, pc), off=(oa, ob),
                    alpha=alpha, beta=beta)
        if c_uplo:
            assert np.abs(np.tril(got) - np.tril(want)).max() < tol, what
            iu = np.triu_indices(M, 1)
            assert (got[iu] == C0[iu]).all() if beta else np.isnan(got[iu]).all(), what
        else:
            assert np.abs(got - want).max() < tol, what
        assert (Cfull[:, N:] == 777.0).all(), what


def test_autograd_adjoints_on_the_real_kernels(gpf):
    """The check of tests/test_gpu_kernels.py::test_autograd_ops_match_torch, on the CPU build."""
    from gpflowSlim._backend import ops
    n, m = 70, 33
    rng = np.random.default_rng(6)
    A = rng.standard_normal((n, n + 3))
    S0 = conv(A @ A.T / (n + 3) + 0.5 * np.eye(n))
    B0, W1 = conv(rng.standard_normal((m, n))), conv(rng.standard_normal((m, n)))

    def run(mine):
        S, B = S0.clone().requires_grad_(True), B0.clone().requires_grad_(True)
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
        return val.detach(), 0.5 * (gS + gS.t()), gB
    for a, b in zip(run(True), run(False)):
        assert float((a - b).abs().max()) < 1e-10 * max(1.0, float(b.abs().max()))


def test_split_k_gemm_through_the_launch_code(gpf):
    """Option "gemm_splitk" (on by default): long-K products with few output tiles are cut into K
    slices by the host code (slice count from the SM count -- 4 in the CPU build), partial tiles
    go to the workspace, a reduction pass applies alpha / beta and the lower-output mask."""
    from gpflowSlim._backend import lib, ops
    rng = np.random.default_rng(17)
    h = lib.handle_for(None)
    close = lambda a, b: np.testing.assert_allclose(a.numpy(), b, rtol=0, atol=1e-12 * max(1.0, np.abs(b).max()))
    A, B = rng.standard_normal((100, 2050)), rng.standard_normal((90, 2050))
    C0 = rng.standard_normal((100, 90))
    Asq = rng.standard_normal((100, 1100))
    A3 = conv(rng.standard_normal((40, 1031)))[:, :1029]                 # odd leading dimension, ragged K
    h.set_option('gemm_splitk', 1)
    try:
        before = h.profile_read(reset=False)[2]
        close(ops.gemm_nt(conv(A), conv(B)), A @ B.T)
        assert h.profile_read(reset=False)[2] - before == 2              # sliced product + reduction
        close(ops.gemm_nt(conv(A), conv(B), alpha=-0.7, beta=0.4, out=conv(C0.copy())), 0.4 * C0 - 0.7 * A @ B.T)
        low = ops.gemm_nt(conv(Asq), conv(Asq), alpha=-1.0, c_uplo=1)
        close(low, -np.tril(Asq @ Asq.T))
        close(ops.gemm_nt(A3, A3), A3.numpy() @ A3.numpy().T)
        before = h.profile_read(reset=False)[2]
        close(ops.gemm_nt(conv(A[:, :200]), conv(B[:, :200])), A[:, :200] @ B[:, :200].T)   # K < 256: not sliced
        assert h.profile_read(reset=False)[2] - before == 1
    finally:
        h.set_option('gemm_splitk', 1)     # the default


def test_fast_path_gram_input_gradient(gpf):
    """gps_gram_bwd of a single stationary kernel with d/dX: the register-tiled kernel writes
    G = dObj/d(d2) and one skinny tensor-core product G [F | 1] gives the input gradient; against
    torch autograd through the oracle kernel, for K(X, X2) (both gradients, and X2 only: one
    swapped pass) and the symmetric K(X), with active dims that skip a column."""
    import gpflowSlim
    from oracle import ref_torch as R
    rng = np.random.default_rng(4)
    n, m, d = 150, 70, 5
    Xn, X2n = rng.standard_normal((n, d)), rng.standard_normal((m, d))
    W, Ws = conv(rng.standard_normal((n, m))), conv(rng.standard_normal((n, n)))
    dims = [0, 2, 3, 4]
    for cls, typ in (('RBF', 'rbf'), ('Matern52', 'matern52')):
        kern = getattr(gpflowSlim.kernels, cls)(4, variance=1.3, lengthscales=[0.7, 0.9, 1.1, 1.3], ARD=True,
                                                active_dims=dims)
        spec = dict(type=typ, variance=torch.tensor(1.3, dtype=torch.float64),
                    lengthscales=torch.tensor([0.7, 0.9, 1.1, 1.3], dtype=torch.float64), active_dims=dims)
        for mode in ('both', 'x2', 'sym'):
            X, X2 = conv(Xn).requires_grad_(mode != 'x2'), conv(X2n).requires_grad_(mode != 'sym')
            Xo, X2o = conv(Xn).requires_grad_(mode != 'x2'), conv(X2n).requires_grad_(mode != 'sym')
            if mode == 'sym':
                val, valo = (kern.K(X) * Ws).sum(), (R.K(spec, Xo) * Ws).sum()
                wrt, wrto = [X], [Xo]
            else:
                val, valo = (kern.K(X, X2) * W).sum(), (R.K(spec, Xo, X2o) * W).sum()
                wrt, wrto = ([X, X2], [Xo, X2o]) if mode == 'both' else ([X2], [X2o])
            g, go = torch.autograd.grad(val, wrt), torch.autograd.grad(valo, wrto)
            for a, b in zip(g, go):
                assert float((a - b).abs().max()) < 1e-11 * float(b.abs().max()), (cls, mode)


def _nkn(gpf, d, prims, widths):
    """Linear -> Product(2) -> Linear -> Product(2) -> Linear(->1) network over `prims`."""
    n1, n2 = widths
    hparams = [dict(name='Linear', params=dict(input_dim=len(prims), output_dim=n1, name='l0')),
               dict(name='Product', params=dict(input_dim=n1, step=2, name='l1')),
               dict(name='Linear', params=dict(input_dim=n1 // 2, output_dim=n2, name='l2')),
               dict(name='Product', params=dict(input_dim=n2, step=2, name='l3')),
               dict(name='Linear', params=dict(input_dim=n2 // 2, output_dim=1, name='l4'))]
    np.random.seed(1)
    return gpf.neural_kernel_network.NeuralKernelNetwork(d, prims, gpf.neural_kernel_network.NKNWrapper(hparams))


@pytest.mark.parametrize('topo', ['c3', 'seven', 'narrow'])
def test_nkn_tensor_core_gram_kernels_against_the_interpreter(gpf, topo):
    """gram_fwd_nkn_kernel / gram_bwd_nkn_kernel (Linear layers, adjoint mat-vecs and weight
    gradients as 8x8x4 FP64 tensor-core products over octets of matrix elements) against the generic
    interpreter kernels (gram_impl 1): K(X) with jitter, K(X, X2), the dense backward for both and the
    fused GPR objective gradient (K^-1 / beta weights, lower tiles only), ragged sizes."""
    from gpflowSlim._backend import lib
    k = gpf.kernels
    d = 3
    if topo == 'c3':
        prims = [k.RBF(d, ARD=True, name='a0'), k.RBF(d, lengthscales=2.0, ARD=True, name='a1'),
                 k.Periodic(d, period=1.0, lengthscales=1.0, name='a2'), k.Periodic(d, period=2.0, name='a3'),
                 k.Linear(d, ARD=True, name='a4'), k.Linear(d, ARD=True, name='a5')]
        widths = (8, 4)
    elif topo == 'seven':   # 7 primitives: the bias column is the 8th; non-ARD, Matern, Linear without ARD
        prims = [k.Matern32(d, lengthscales=1.5, name='b0'), k.RBF(d, name='b1'), k.Linear(d, name='b2'),
                 k.Matern52(2, active_dims=[0, 2], ARD=True, name='b3'), k.Periodic(d, period=1.5, name='b4'),
                 k.Matern12(d, name='b5'), k.RBF(1, active_dims=[1], name='b6')]
        widths = (6, 8)
    else:                   # fewer than 4 primitives and the narrowest layers
        prims = [k.RBF(d, ARD=True, name='c0'), k.Linear(d, name='c1')]
        widths = (2, 2)
    rng = np.random.default_rng(11)
    n, m = 77, 45
    X, X2 = conv(rng.standard_normal((n, d))), conv(rng.standard_normal((m, d)))
    Y = conv(rng.standard_normal((n, 2)))
    W, Ws = conv(rng.standard_normal((n, m))), conv(rng.standard_normal((n, n)))
    h = lib.handle_for(None)
    res = {}
    for impl in (0, 1):
        h.set_option('gram_impl', impl)
        try:
            kern = _nkn(gpf, d, prims, widths)
            params = [p.unconstrained_tensor for p in kern.parameters]
            K, K2 = kern.K(X), kern.K(X, X2)
            g1 = torch.autograd.grad((K * Ws).sum(), params, allow_unused=True)
            g2 = torch.autograd.grad((K2 * W).sum(), params, allow_unused=True)
            model = gpf.models.GPR(X, Y, kern=kern, name='nkn_fast_%s_%d' % (topo, impl))
            obj = model.objective
            g3 = torch.autograd.grad(obj, [p.unconstrained_tensor for p in model.parameters])
            res[impl] = [K.detach(), K2.detach(), obj.detach()] + [g for gs in (g1, g2, g3) for g in gs if g is not None]
        finally:
            h.set_option('gram_impl', 0)
    assert len(res[0]) == len(res[1]) > 10
    same = True
    for a, b in zip(res[0], res[1]):
        assert float((a - b).abs().max()) <= 1e-11 * max(float(b.abs().max()), 1e-30)
        same = same and torch.equal(a, b)
    assert not same          # another summation order: the two runs really took different kernels


def test_nkn_tensor_core_backward_with_several_column_tiles_per_block(gpf):
    """The column-tile loop of gram_bwd_nkn_kernel (a CTA walks tiles jt, jt + njc, ...; feature tiles and
    beta rows are reloaded between two barriers; the reduction scratch aliases the feature tiles): N = 520
    gives 9 row tiles and -- with the 4 SMs of the CPU build -- 8 column chunks, so the last row tiles have
    CTAs with two tiles.  Fused GPR gradient (lower tiles, K^-1 / beta weights, R = 2) against the interpreter."""
    from gpflowSlim._backend import lib
    k = gpf.kernels
    d, n = 3, 520
    rng = np.random.default_rng(12)
    X, Y = conv(rng.standard_normal((n, d))), conv(rng.standard_normal((n, 2)))
    h = lib.handle_for(None)
    res = {}
    for impl in (0, 1):
        h.set_option('gram_impl', impl)
        try:
            kern = _nkn(gpf, d, [k.RBF(d, ARD=True, name='m0'), k.Periodic(d, period=1.3, name='m1'),
                                 k.Linear(d, ARD=True, name='m2'), k.Matern32(d, name='m3'), k.RBF(d, name='m4')], (8, 4))
            model = gpf.models.GPR(X, Y, kern=kern, name='nkn_tiles_%d' % impl)
            obj = model.objective
            res[impl] = [obj.detach()] + list(torch.autograd.grad(obj, [p.unconstrained_tensor for p in model.parameters]))
        finally:
            h.set_option('gram_impl', 0)
    for a, b in zip(res[0], res[1]):
        assert float((a - b).abs().max()) <= 1e-10 * max(float(b.abs().max()), 1e-30)
    assert not all(torch.equal(a, b) for a, b in zip(res[0], res[1]))


def test_fuzz_nkn_tensor_core_kernels(gpf):
    """Randomised differential test of gram_fwd_nkn_kernel / gram_bwd_nkn_kernel against the interpreter:
    random numbers of primitives (1-7) of random types / ARD / active dimensions (1-8 of them), random
    layer widths, ragged sizes; K(X), K(X, X2), both dense backward passes and, every third case, the fused
    GPR gradient.  GPSLIM_FUZZ=<n> cases (default 16; 600 were run clean when this was written)."""
    from gpflowSlim._backend import lib
    k = gpf.kernels
    ncases = int(os.environ.get('GPSLIM_FUZZ', '16'))
    rng = np.random.default_rng(2024)
    h = lib.handle_for(None)
    for it in range(ncases):
        D = int(rng.integers(1, 10))
        P = int(rng.integers(1, 8))
        specs = []
        for p in range(P):
            nd = int(rng.integers(1, min(D, 8) + 1))
            dims = sorted(rng.choice(D, size=nd, replace=False).tolist())
            typ = str(rng.choice(['RBF', 'Matern12', 'Matern32', 'Matern52', 'Exponential', 'Linear', 'Periodic']))
            ard = bool(rng.integers(0, 2)) and typ != 'Periodic'
            specs.append((typ, nd, dims, ard, float(rng.uniform(0.5, 2.0)), rng.uniform(0.7, 2.5, size=nd if ard else 1)))
        widths = (2 * int(rng.integers(1, 5)), 2 * int(rng.integers(1, 5)))
        n, m = int(rng.integers(1, 150)), int(rng.integers(1, 100))

        def make(tag):
            prims = []
            for p, (typ, nd, dims, ard, var, ls) in enumerate(specs):
                kw = dict(active_dims=dims, name='f%d_%d_%s' % (it, p, tag))
                if typ == 'Linear':
                    prims.append(k.Linear(nd, variance=(ls if ard else float(ls[0])), ARD=ard, **kw))
                elif typ == 'Periodic':
                    prims.append(k.Periodic(nd, period=float(ls[0]) + 0.5, variance=var, lengthscales=1.2, **kw))
                else:
                    prims.append(getattr(k, typ)(nd, variance=var, lengthscales=(ls if ard else float(ls[0])), ARD=ard, **kw))
            return _nkn(gpf, D, prims, widths)
        X, X2 = conv(rng.standard_normal((n, D))), conv(rng.standard_normal((m, D)))
        W, Ws = conv(rng.standard_normal((n, m))), conv(rng.standard_normal((n, n)))
        Y = conv(rng.standard_normal((n, int(rng.integers(1, 4)))))
        res = {}
        for impl in (0, 1):
            h.set_option('gram_impl', impl)
            try:
                kern = make(str(impl))
                params = [p.unconstrained_tensor for p in kern.parameters]
                K, K2 = kern.K(X), kern.K(X, X2)
                out = [K.detach(), K2.detach()]
                out += [g for g in torch.autograd.grad((K * Ws).sum(), params, allow_unused=True) if g is not None]
                out += [g for g in torch.autograd.grad((K2 * W).sum(), params, allow_unused=True) if g is not None]
                if it % 3 == 0:
                    model = gpf.models.GPR(X, Y, kern=kern, name='fz%d_%d' % (it, impl))
                    obj = model.objective
                    out += [obj.detach()] + list(torch.autograd.grad(obj, [p.unconstrained_tensor for p in model.parameters]))
                res[impl] = out
            finally:
                h.set_option('gram_impl', 0)
        assert len(res[0]) == len(res[1])
        # the parameter gradients of one objective are sums of the same large terms with different signs: a
        # small one is a cancelled sum, so errors are measured against the largest gradient of the case too
        gmax = max(float(b.abs().max()) for b in res[1][2:])
        for j, (a, b) in enumerate(zip(res[0], res[1])):
            scale = max(float(b.abs().max()), 1e-4 * gmax if j >= 2 else 0.0, 1e-30)
            err = float((a - b).abs().max()) / scale
            assert err <= 1e-9, (it, j, err, specs, widths, n, m)


def test_networks_outside_the_tensor_core_shape_keep_the_interpreter(gpf):
    """nkn_match (gram.cu) must hand anything but Linear -> Product(2) -> Linear -> Product(2) -> Linear(->1)
    over <= 7 primitives of <= 8 active dimensions to the interpreter: for such programs gram_impl 0 and 1
    run the SAME kernels, so the results are bit-identical (the specialised kernels sum in another order,
    see the `not same` assertion of the test above)."""
    from gpflowSlim._backend import lib
    k = gpf.kernels
    nn = gpf.neural_kernel_network

    def net(prims, hparams, d):
        np.random.seed(2)
        return nn.NeuralKernelNetwork(d, prims, nn.NKNWrapper(hparams))

    def lin(i, o, name):
        return dict(name='Linear', params=dict(input_dim=i, output_dim=o, name=name))

    def prod(i, name, step=2):
        return dict(name='Product', params=dict(input_dim=i, step=step, name=name))
    cases_ = {
        # 8 primitives: no free column for the bias in the first 8 x 8 tile
        'eight': (3, lambda: net([k.RBF(3, name='e%d' % i) for i in range(8)],
                                 [lin(8, 4, 'l0'), prod(4, 'l1'), lin(2, 2, 'l2'), prod(2, 'l3'), lin(1, 1, 'l4')], 3)),
        # three layers only
        'short': (3, lambda: net([k.RBF(3, name='s0'), k.Linear(3, name='s1')],
                                 [lin(2, 4, 'l0'), prod(4, 'l1'), lin(2, 1, 'l2')], 3)),
        # a primitive with 9 active dimensions
        'wide': (9, lambda: net([k.RBF(9, ARD=True, name='w0'), k.Linear(9, name='w1')],
                                [lin(2, 4, 'l0'), prod(4, 'l1'), lin(2, 2, 'l2'), prod(2, 'l3'), lin(1, 1, 'l4')], 9)),
        # Product over 4 inputs at a time
        'step4': (3, lambda: net([k.RBF(3, name='p0'), k.Matern32(3, name='p1')],
                                 [lin(2, 8, 'l0'), prod(8, 'l1', step=4), lin(2, 2, 'l2'), prod(2, 'l3'), lin(1, 1, 'l4')], 3)),
    }
    h = lib.handle_for(None)
    rng = np.random.default_rng(5)
    for name, (d, make) in cases_.items():
        X, X2 = conv(rng.standard_normal((40, d))), conv(rng.standard_normal((21, d)))
        W = conv(rng.standard_normal((40, 21)))
        res = {}
        for impl in (0, 1):
            h.set_option('gram_impl', impl)
            try:
                kern = make()
                params = [p.unconstrained_tensor for p in kern.parameters]
                K, K2 = kern.K(X), kern.K(X, X2)
                g = torch.autograd.grad((K2 * W).sum(), params, allow_unused=True)
                res[impl] = [K.detach(), K2.detach()] + [x for x in g if x is not None]
            finally:
                h.set_option('gram_impl', 0)
        assert len(res[0]) == len(res[1]) > 4, name
        for a, b in zip(res[0], res[1]):
            assert torch.equal(a, b), name


def test_prefix_solves_with_big_leaves(gpf):
    """gps_trsm_rlt_prefix / gps_trsm_rln_prefix with option trsm_leaf = 256: the aligned 256-blocks
    are solved by one product with their explicit inverses (built for all blocks at once by
    strided-batch products), the ragged tail by 128-strips; rows of U = L^-T and of K^-1 against
    dense inverses, for rows that start inside and outside a leaf."""
    from gpflowSlim._backend import dist_gpr, lib
    rng = np.random.default_rng(11)
    n, bs = 600, 128                       # two aligned 256-leaves + a ragged 88-wide tail
    A = rng.standard_normal((n, n + 3))
    S = A @ A.T / (n + 3) + 0.5 * np.eye(n)
    L = np.linalg.cholesky(S)
    U, Kinv = np.linalg.inv(L).T, np.linalg.inv(S)
    rows = np.concatenate([np.arange(0, 40), np.arange(128, 200), np.arange(384, 420), np.arange(512, 560)])
    act = rows // bs * bs
    h = lib.handle_for(None)

    class Be(dist_gpr.CudaBackend):
        def __init__(self):
            self._L, self.device = lib, torch.device('cpu')
    be = Be()
    Ld = conv(L) + torch.triu(torch.full((n, n), 7.0, dtype=torch.float64), 1)       # junk above the diagonal
    Lt = conv(L.T.copy())
    results = {}
    for leaf in (128, 256):
        h.set_option('trsm_leaf', leaf)
        try:
            B = np.zeros((len(rows), n))
            B[np.arange(len(rows)), rows] = 1.0
            Bd = conv(B)
            be.trsm_rlt_prefix_(Ld, Bd, act)
            np.testing.assert_allclose(Bd.numpy(), U[rows], rtol=0, atol=1e-12 * np.abs(U).max())
            be.trsm_rln_prefix_(Ld, Lt, Bd, act)
            got = Bd.numpy()
            keep = np.arange(n)[None, :] >= act[:, None]
            assert np.abs(got - Kinv[rows])[keep].max() < 1e-11 * np.abs(Kinv).max()
            results[leaf] = got[keep]
        finally:
            h.set_option('trsm_leaf', 512)
    assert np.abs(results[128] - results[256]).max() < 1e-11 * np.abs(Kinv).max()


@pytest.mark.parametrize('n,m', [(70, 33), (200, 140), (1, 1)])
def test_library_side_adjoints(gpf, n, m):
    """gps_potri / gps_chol_bwd / gps_trsm_bwd (csrc/adjoint.cu), with U computed inside and with
    a caller-supplied U, against torch autograd and a dense inverse."""
    from gpflowSlim._backend import ops
    rng = np.random.default_rng(n + m)
    A = rng.standard_normal((n, n + 3))
    S = conv(A @ A.T / (n + 3) + 0.5 * np.eye(n)).requires_grad_(True)
    B = conv(rng.standard_normal((m, n))).requires_grad_(True)
    Lbar_in, Xbar_in = conv(rng.standard_normal((n, n))), conv(rng.standard_normal((m, n)))
    L = torch.linalg.cholesky(S)
    X = torch.linalg.solve_triangular(L, B.t(), upper=False).t()
    (want_A,) = torch.autograd.grad((L * torch.tril(Lbar_in)).sum(), [S], retain_graph=True)
    want_A = 0.5 * (want_A + want_A.t())
    Lleaf = L.detach().clone().requires_grad_(True)
    Xl = torch.linalg.solve_triangular(Lleaf, B.detach().t(), upper=False).t()
    want_L, = torch.autograd.grad((Xl * Xbar_in).sum(), [Lleaf])
    want_B, = torch.autograd.grad((X * Xbar_in).sum(), [B], retain_graph=True)
    Ld = L.detach()
    close = lambda a, b: np.testing.assert_allclose(a.numpy(), b.detach().numpy(), rtol=0,
                                                    atol=1e-10 * max(1.0, float(b.abs().max())))
    for U in (None, ops.tri_inv_t(Ld)):
        close(ops.chol_bwd(Ld, Lbar_in + torch.triu(torch.full((n, n), 9.0, dtype=torch.float64), 1), U), want_A)
        Bbar, Lb = ops.trsm_bwd(Ld, X.detach(), Xbar_in, U)
        close(Bbar, want_B)
        close(Lb, torch.tril(want_L))
        assert ops.trsm_bwd(Ld, X.detach(), Xbar_in, U, want_lbar=False)[1] is None
    close(ops.potri(Ld), torch.tril(torch.linalg.inv(S.detach())))


# case -> seconds on 8 host cores when this was written (all clean, worst relative error in brackets):
#   kernels 1 [8e-16], kernels_extra 8 [1e-13], svgp_white_diag 8 [4e-14], nkn 15 [2e-15],
#   svgp_nonwhite_diag 16 [3e-12], functions 19 [4e-13], gpr_features 19 [1e-10], mc_models 32 [1e-10],
#   sgpr 41 [1e-12], gpr_composed 65 [2e-14], gpr_misc 69 [2e-14], likelihoods_extra 0.2 [0], priors 5 [2e-15],
#   large_d 43 [4e-15], lbfgs 225 [2e-12], gpr_white 15 [5e-14]  (mc_models, likelihoods_extra, large_d had never run on a GPU then)
_DEFAULT_CASES = ['kernels', 'svgp_white_diag', 'nkn']
_FULL_CASES = ['gpr_white', 'kernels_extra', 'svgp_nonwhite_diag', 'functions', 'gpr_features', 'mc_models', 'sgpr',
               'gpr_composed', 'gpr_misc', 'likelihoods_extra', 'priors', 'large_d', 'lbfgs']


@pytest.mark.parametrize('name', _DEFAULT_CASES +
                         (_FULL_CASES if FULL else []))
def test_golden_cases_through_the_real_library(gpf, golden, name):
    """The parity contract of tests/test_gpu_parity.py -- reference golden vectors, 1e-8 relative --
    with the shipped library code running on the CPU (fused one-call GPR objective and gradient,
    Gram kernels, recursive Cholesky, conditionals, KL terms)."""
    gold = golden(name)
    res = cases.run_case(gpf, name, conv)
    assert set(res) == set(gold)
    for key in sorted(gold):
        a, b = np.asarray(res[key], dtype=np.float64), np.asarray(gold[key], dtype=np.float64)
        e = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.size else 0.0
        assert e < (1e-12 if key.startswith('param/') else 1e-8), '%s/%s: %.3e' % (name, key, e)


def test_fused_gpr_objective_gradient_and_prediction(gpf):
    """smoke() of __graft_entry__.py, on the CPU build: gps_gpr_nlml_fwd_bwd / gps_gpr_predict
    against the oracle."""
    from oracle import ref_torch as R
    X, Y = cases.synth_gpr(140, 4, seed=0)
    Xs = np.random.default_rng(1).standard_normal((20, 4))
    m = gpf.models.GPR(conv(X), conv(Y), kern=gpf.kernels.RBF(4, ARD=True, lengthscales=2.0))
    obj = m.objective
    grads = torch.autograd.grad(obj, [p.unconstrained_tensor for p in m.parameters])
    with torch.no_grad():
        mu, var = m.predict_f(conv(Xs))
    raw = [torch.tensor(R.softplus_inv(v), dtype=torch.float64, requires_grad=True)
           for v in (1.0, 2.0 * np.ones(4), 0.1)]
    spec = dict(type='rbf', variance=R.softplus_fwd(raw[0]), lengthscales=R.softplus_fwd(raw[1]))
    noise = R.softplus_fwd(raw[2])
    o = R.gpr_nlml(spec, torch.tensor(X), torch.tensor(Y), noise)
    go = torch.autograd.grad(o, raw)
    mo, vo = R.gpr_predict(spec, torch.tensor(X), torch.tensor(Y), noise, torch.tensor(Xs))
    rel = lambda a, b: float((a.detach().reshape(-1) - b.detach().reshape(-1)).abs().max() / b.detach().abs().max())
    errs = [rel(obj, o)] + [rel(a, b) for a, b in zip(grads, go)] + [rel(mu, mo), rel(var, vo)]
    assert max(errs) < 1e-9, errs


def test_fuzz_fused_gpr_against_the_op_by_op_path(gpf):
    """gps_gpr_nlml_fwd_bwd / gps_gpr_predict (one library call: Gram, factorisation, ride-along solves,
    fused weight formation + Gram backward) against the same model evaluated one autograd op at a time
    (`fused=False`: gram, cholesky, triangular solves with their hand-written adjoints), for every kernel of
    the zoo and the NKN network, random sizes 2..330, 1..4 output columns, random noise.  Both paths run on
    the CPU build; they share kernels but not orchestration, adjoint formulas or summation order.
    GPSLIM_FUZZ=<n> cases (default 6; 150 were run clean when this was written)."""
    ncases = int(os.environ.get('GPSLIM_FUZZ', '6'))
    rng = np.random.default_rng(31)
    d = 3
    zoo = cases._kernel_zoo(gpf, d) + [('nkn', lambda: cases.nkn_c3_kernel(gpf, d))]
    for it in range(ncases):
        name, make = zoo[int(rng.integers(0, len(zoo)))]
        n, r, ns = int(rng.integers(2, 331)), int(rng.choice([1, 1, 2, 3, 4, 9, 16, 17])), int(rng.integers(1, 40))
        # (17 output columns: beyond the fused call's 16, GPR falls back to the op-by-op path by itself)
        X, Y = conv(rng.standard_normal((n, d))), conv(rng.standard_normal((n, r)))
        Xs = conv(rng.standard_normal((ns, d)))
        noise = float(rng.uniform(0.05, 1.0))
        out = {}
        for fused in (True, False):
            m = gpf.models.GPR(X, Y, kern=make(), obs_var=noise, fused=fused, name='fz_gpr_%d_%d' % (it, fused))
            obj = m.objective
            gr = torch.autograd.grad(obj, [p.unconstrained_tensor for p in m.parameters])
            with torch.no_grad():
                mu, var = m.predict_f(Xs)
            out[fused] = [obj.detach()] + [g.detach() for g in gr] + [mu, var]
        gmax = max(float(b.abs().max()) for b in out[False][1:-2])
        for j, (a, b) in enumerate(zip(out[True], out[False])):
            scale = max(float(b.abs().max()), 1e-4 * gmax if 1 <= j < len(out[True]) - 2 else 0.0, 1e-30)
            err = float((a - b).abs().max()) / scale
            assert err < 1e-8, (it, name, n, r, j, err)


def test_fuzz_svgp_bound_against_the_torch_double(gpf, monkeypatch):
    """The SVGP bound (svgp.py:94-125: Kuu / Kuf Gram kernels incl. the inducing-input gradient, the
    conditional with its triangular-aware products, the KL term, the Gaussian expectations) with its
    gradients w.r.t. every parameter AND the inducing inputs: the CPU build of the library against the
    torch-CPU double of the ops layer (plain torch autograd), random batch 20..260, 4..130 inducing points,
    1..2 latents, whitened or not, diagonal or full q_sqrt, every kernel of the zoo.
    GPSLIM_FUZZ=<n> cases (default 5; 150 were run clean when this was written)."""
    import cpu_ops_double
    ncases = int(os.environ.get('GPSLIM_FUZZ', '5'))
    rng = np.random.default_rng(53)
    d = 3
    # (the bare Linear kernels are left out: Kuu of rank 3 + jitter has condition 1e8, and the two
    # implementations then differ by rounding x condition = 1e-5, as any two would)
    zoo = [z for z in cases._kernel_zoo(gpf, d) if not z[0].startswith('lin_')]
    specs = []
    for it in range(ncases):
        specs.append(dict(kern=int(rng.integers(0, len(zoo))), n=int(rng.integers(20, 261)), m=int(rng.integers(4, 131)),
                          lat=int(rng.integers(1, 3)), whiten=bool(rng.integers(0, 2)), q_diag=bool(rng.integers(0, 2)),
                          seed=int(rng.integers(0, 1 << 30))))

    def run(sp, tag):
        r = np.random.default_rng(sp['seed'])
        X = r.standard_normal((sp['n'], d))
        Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * r.standard_normal((sp['n'], sp['lat']))
        Z = r.standard_normal((sp['m'], d))
        m = gpf.models.SVGP(conv(X), conv(Y), zoo[sp['kern']][1](), gpf.likelihoods.Gaussian(var=0.2), Z=Z,
                            whiten=sp['whiten'], q_diag=sp['q_diag'], num_data=5 * sp['n'], name='fz_svgp_' + tag)
        with torch.no_grad():
            qm = m._q_mu.unconstrained_tensor
            qm.copy_(torch.as_tensor(0.3 * r.standard_normal(tuple(qm.shape))).to(qm))
            qs = m._q_sqrt.unconstrained_tensor
            qs.add_(torch.as_tensor(0.05 * r.standard_normal(tuple(qs.shape))).to(qs))
        params = [p.unconstrained_tensor for p in m.parameters] + [m.feature._Z.unconstrained_tensor]
        obj = m.objective
        with torch.no_grad():
            sp['cond'] = float(np.linalg.cond(m.feature.Kuu(m.kern, jitter=gpf.settings.numerics.jitter_level).numpy()))
        return [obj.detach()] + [g.detach() for g in torch.autograd.grad(obj, params)]
    got = [run(sp, 'lib%d' % i) for i, sp in enumerate(specs)]
    cpu_ops_double.install(monkeypatch)
    want = [run(sp, 'dbl%d' % i) for i, sp in enumerate(specs)]
    for i, (g, w) in enumerate(zip(got, want)):
        gmax = max(float(b.abs().max()) for b in w[1:])
        # the north-star tolerance (1e-8; Matern-type kernels reach a few 1e-9 through sqrt(d2 + 1e-12) at the
        # coincident points of Kuu's diagonal); beyond that two correct implementations differ by rounding x
        # condition of Kuu (inducing points are random here, so Kuu + 1e-6 I reaches 1e7)
        tol = max(1e-8, 1e-14 * specs[i]['cond'])
        for j, (a, b) in enumerate(zip(g, w)):
            scale = max(float(b.abs().max()), 1e-4 * gmax if j else 0.0, 1e-30)
            err = float((a - b).abs().max()) / scale
            assert err < tol, (i, zoo[specs[i]['kern']][0], specs[i], j, err, tol)
    # two implementations really ran: rounding differs somewhere
    assert not all(torch.equal(a, b) for g, w in zip(got, want) for a, b in zip(g, w))


@pytest.mark.skipif(not FULL, reason='GPSLIM_CPU_LIB_FULL=0')
def test_experimental_switches_through_the_real_dispatch_code(gpf, golden):
    """The switches with an alternative implementation (tests/test_gpu_switches.py), through the
    shipped host code on the CPU build: gram_impl = 2 (shared-memory interpreter kernels: their
    launch configuration and shared-memory sizing are host code) on the NKN case, and the
    triangular-aware matmul adjoints on a non-whitened SVGP."""
    from gpflowSlim._backend import lib, ops

    def check(name):
        gold = golden(name)
        res = cases.run_case(gpf, name, conv)
        for key in sorted(gold):
            a, b = np.asarray(res[key], dtype=np.float64), np.asarray(gold[key], dtype=np.float64)
            e = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.size else 0.0
            assert e < (1e-12 if key.startswith('param/') else 1e-8), '%s/%s: %.3e' % (name, key, e)
    h = lib.handle_for(None)
    h.set_option('gram_impl', 2)
    try:
        check('nkn')
    finally:
        h.set_option('gram_impl', 0)
    ops.TRI_AWARE_ADJOINTS[0] = False        # the defaults are True / True: run the other settings
    try:
        check('svgp_nonwhite_diag')
    finally:
        ops.TRI_AWARE_ADJOINTS[0] = True
    ops.FUSED_ADJOINTS[0] = False            # composed adjoints instead of one library call each
    try:
        check('svgp_white_diag')
        check('functions')
    finally:
        ops.FUSED_ADJOINTS[0] = True


# ----------------------------------------------------------------------------- distributed path
def _dist_worker(rank, world, port, so_path, n, r, block, out_q, schedule=True):
    """One rank of the REAL distributed GPR algorithm (_backend/dist_gpr.py: block-row layout,
    look-ahead schedule, panel broadcasts / all-gathers, inverse rows, gradient contraction) on
    the REAL kernels (gps_gemm_nt_rowmap, gps_trsm_rl{t,n}_prefix, gps_gpr_weight_rows, ...) of the
    CPU build, with gloo instead of NCCL and issue-order execution instead of CUDA streams."""
    import contextlib
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here), os.path.join(os.path.dirname(here), 'gpflow-slim_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import gpflowSlim
    from gpflowSlim._backend import dist_gpr, lib, ops
    from oracle import ref_torch as R
    torch.set_num_threads(1)
    lib.LIB_PATH = so_path

    class CpuHandle(lib.Handle):
        def sync_stream(self):
            pass
    holder = []

    def handle_for(_):
        if not holder:
            holder.append(CpuHandle(0))
        return holder[0]
    lib.handle_for = ops.handle_for = handle_for
    gpflowSlim.settings.device = 'cpu'

    class CpuLibBackend(dist_gpr.CudaBackend):
        poison = True                      # NaN-filled buffers: reads of never-written data show up

        def __init__(self):
            self._L, self.device = lib, torch.device('cpu')

        def streams(self):
            return 'main', 'chain', 'tb', 'gather', 'narrow'

        def on(self, stream):
            return contextlib.nullcontext()

        def record(self, stream):
            return None

        def wait(self, stream, event):
            pass
    if world > 1:
        dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    d = 3
    X, Y = cases.synth_gpr(n, d)
    rng = np.random.default_rng(5)
    Y = np.concatenate([Y] + [rng.standard_normal((n, 1)) for _ in range(r - 1)], 1)
    kern = gpflowSlim.kernels.RBF(d, ARD=True, lengthscales=1.7, variance=1.3)
    prog = kern.program()
    theta = prog.theta('cpu').detach()
    nlml, dth, dnz, dY = dist_gpr.nlml_and_grad(prog, theta, 0.1, torch.tensor(X), torch.tensor(Y), block=block,
                                                backend=CpuLibBackend(), lookahead=schedule)
    th = theta.clone().requires_grad_(True)
    nz = torch.tensor(0.1, dtype=torch.float64, requires_grad=True)
    Yt = torch.tensor(Y, requires_grad=True)
    obj = R.gpr_nlml(dict(type='rbf', variance=th[0], lengthscales=th[1:1 + d]), torch.tensor(X), Yt, nz)
    g = torch.autograd.grad(obj, [th, nz, Yt])
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    out_q.put((rank, [rel(nlml, obj.detach()), rel(dth, g[0]), rel(dnz, g[1]), rel(dY, g[2])]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_fuzz_distributed_path_at_world_1_against_the_fused_call(gpf):
    """dist_gpr.nlml_and_grad with one rank (every block row local, no collective) on the CPU build: the
    block-row factorisation in all three schedules, the row-map GEMM, the prefix solves for the big-leaf
    sizes, the inverse rows and gps_gpr_weight_rows -- against the fused single-call path
    (gps_gpr_nlml_fwd_bwd), random orders 2..700, 1..3 output columns, block 128 / 256 / 512, kernels of the
    zoo (any program takes this path).  The multi-rank layouts are the gloo test below.
    GPSLIM_FUZZ=<n> cases (default 3; 80 were run clean when this was written)."""
    import contextlib
    from gpflowSlim._backend import dist_gpr, lib, ops
    ncases = int(os.environ.get('GPSLIM_FUZZ', '3'))
    rng = np.random.default_rng(91)
    d = 3

    class Be(dist_gpr.CudaBackend):
        poison = True

        def __init__(self):
            self._L, self.device = lib, torch.device('cpu')

        def streams(self):
            return 'main', 'chain', 'tb', 'gather', 'narrow'

        def on(self, stream):
            return contextlib.nullcontext()

        def record(self, stream):
            return None

        def wait(self, stream, event):
            pass
    zoo = cases._kernel_zoo(gpf, d) + [('nkn', lambda: cases.nkn_c3_kernel(gpf, d))]
    h = lib.handle_for(None)
    for it in range(ncases):
        name, make = zoo[int(rng.integers(0, len(zoo)))]
        n, r = int(rng.choice([int(rng.integers(2, 130)), int(rng.integers(130, 701))])), int(rng.integers(1, 4))
        block = int(rng.choice([128, 256, 512]))
        schedule = [True, 'v2', False][int(rng.integers(0, 3))]
        leaf = int(rng.choice([128, 256, 512]))
        noise = float(rng.uniform(0.05, 0.8))
        X, Y = conv(rng.standard_normal((n, d))), conv(rng.standard_normal((n, r)))
        kern = make()
        prog = kern.program()
        theta = prog.theta('cpu').detach()
        h.set_option('trsm_leaf', leaf)
        try:
            nlml, dth, dnz, dY = dist_gpr.nlml_and_grad(prog, theta, noise, X, Y, block=block, backend=Be(),
                                                        lookahead=schedule)
        finally:
            h.set_option('trsm_leaf', 512)
        th = theta.clone().requires_grad_(True)
        nz = torch.tensor(noise, dtype=torch.float64, requires_grad=True)
        Yt = Y.clone().requires_grad_(True)
        # fused path through its autograd Function (theta is the program's own parameter vector there)
        obj = -ops._GprLogLik.apply(th, nz, Yt, X, prog)
        g = torch.autograd.grad(obj, [th, nz, Yt])
        gmax = max(float(x.abs().max()) for x in g[:2])
        for j, (a, b) in enumerate(zip([nlml, dth, dnz, dY], [obj.detach()] + list(g))):
            scale = max(float(b.abs().max()), 1e-4 * gmax if j in (1, 2) else 0.0, 1e-30)
            err = float((a - b).abs().max()) / scale
            assert err < 1e-8, (it, name, n, r, block, schedule, leaf, j, err)


class _ThreadComm(object):
    """The communicator interface of dist_gpr._Comm for P ranks that are THREADS of this process: a
    collective is a rendezvous at a barrier around a shared slot.  Every rank issues its collectives in the
    same order (what NCCL requires, and what the gloo tests already rely on), so one barrier serves all
    three channels.  Lets the randomised test below run dozens of multi-rank configurations without
    spawning processes."""

    def __init__(self, world, rank, shared):
        self.world, self.rank, self.sh = world, rank, shared

    def _sync(self):
        self.sh['bar'].wait(timeout=600)

    def broadcast(self, t, src, which='chain'):
        if self.rank == src:
            self.sh['slot'] = t
        self._sync()
        if self.rank != src:
            t.copy_(self.sh['slot'])
        self._sync()

    def all_gather(self, out, inp):
        self.sh['parts'][self.rank] = inp
        self._sync()
        out.copy_(torch.cat([p.reshape(-1) for p in self.sh['parts']]).reshape(out.shape))
        self._sync()

    def all_reduce_sum(self, t):
        self.sh['parts'][self.rank] = t.clone()
        self._sync()
        total = sum(self.sh['parts'][1:], self.sh['parts'][0].clone())      # fixed order on every rank
        self._sync()
        t.copy_(total)


def test_fuzz_distributed_path_with_ranks_as_threads(gpf, monkeypatch):
    """dist_gpr.nlml_and_grad and dist_gpr.predict at world sizes 2..5 with the ranks as threads of this
    process (_ThreadComm), every rank on the real kernels of the CPU build with NaN-poisoned buffers:
    snake block-row layouts with ragged last blocks and more ranks than block rows, all three schedules,
    1..3 output columns, fewer test points than ranks -- against the fused single-call path.  The emulated
    library is not re-entrant, so its calls are serialised by a lock; each rank has its own handle.
    GPSLIM_FUZZ=<n> cases (default 3; 60 were run clean when this was written)."""
    import contextlib
    import ctypes
    import threading
    from gpflowSlim._backend import dist_gpr, lib, ops
    ncases = int(os.environ.get('GPSLIM_FUZZ', '3'))
    rng = np.random.default_rng(123)
    d = 3
    lock = threading.RLock()

    class LockedLib(object):
        def __init__(self, cdll):
            self._cdll = cdll

        def __getattr__(self, name):
            fn = getattr(self._cdll, name)

            def call(*a):
                with lock:
                    return fn(*a)
            return call

    class RankHandle(lib.Handle):
        def __init__(self):
            lib.Handle.__init__(self, 0)
            self.lib = LockedLib(self.lib)

        def sync_stream(self):
            pass
    tls = threading.local()
    main_handle = lib.handle_for(None)

    def handle_for(_):
        if threading.current_thread() is threading.main_thread():
            return main_handle
        if not hasattr(tls, 'h'):
            with lock:
                tls.h = RankHandle()
        return tls.h
    monkeypatch.setattr(lib, 'handle_for', handle_for)
    monkeypatch.setattr(ops, 'handle_for', handle_for)
    monkeypatch.setattr(dist_gpr, '_Comm', lambda g: g)

    class Be(dist_gpr.CudaBackend):
        poison = True

        def __init__(self):
            self._L, self.device = lib, torch.device('cpu')

        def streams(self):
            return 'main', 'chain', 'tb', 'gather', 'narrow'

        def on(self, stream):
            return contextlib.nullcontext()

        def record(self, stream):
            return None

        def wait(self, stream, event):
            pass
    zoo = [z for z in cases._kernel_zoo(gpf, d) if z[0] in ('rbf_ard', 'm32_ard', 'sum', 'product', 'periodic')]
    for it in range(ncases):
        name, make = zoo[int(rng.integers(0, len(zoo)))]
        world = int(rng.integers(2, 6))
        n, r = int(rng.integers(2, int(os.environ.get('GPSLIM_FUZZ_MAXN', '560')))), int(rng.choice([1, 1, 2, 3, 16]))
        ns = int(rng.integers(1, 30))
        block = int(rng.choice([128, 256]))
        schedule = [True, 'v2', False][int(rng.integers(0, 3))]
        noise = float(rng.uniform(0.05, 0.8))
        X, Y = conv(rng.standard_normal((n, d))), conv(rng.standard_normal((n, r)))
        Xs = conv(rng.standard_normal((ns, d)))
        kern = make()
        prog = kern.program()
        theta = prog.theta('cpu').detach()
        kd = kern.Kdiag(Xs).detach()
        shared = dict(bar=threading.Barrier(world), slot=None, parts=[None] * world)
        results, errors = [None] * world, []

        def rank_main(rank):
            try:
                comm = _ThreadComm(world, rank, shared)
                be = Be()
                out = dist_gpr.nlml_and_grad(prog, theta, noise, X, Y, block=block, group=comm, backend=be,
                                             lookahead=schedule)
                mu, var = dist_gpr.predict(prog, theta, noise, X, Y, Xs, kd, block=block, group=comm, backend=be,
                                           lookahead=schedule)
                results[rank] = list(out) + [mu, var]
            except BaseException as e:      # noqa: a failing rank must not leave the others at the barrier
                errors.append((rank, repr(e)))
                shared['bar'].abort()
        threads = [threading.Thread(target=rank_main, args=(q,)) for q in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, (it, name, world, n, r, block, schedule, errors[:2])
        th = theta.clone().requires_grad_(True)
        nz = torch.tensor(noise, dtype=torch.float64, requires_grad=True)
        Yt = Y.clone().requires_grad_(True)
        obj = -ops._GprLogLik.apply(th, nz, Yt, X, prog)
        g = torch.autograd.grad(obj, [th, nz, Yt])
        mu0, var0 = ops.gpr_predict(prog, X, Y, torch.tensor(noise, dtype=torch.float64), Xs)
        want = [obj.detach()] + list(g) + [mu0, var0]
        gmax = max(float(x.abs().max()) for x in g[:2])
        for rank in range(world):
            for j, (a, b) in enumerate(zip(results[rank], want)):
                scale = max(float(b.abs().max()), 1e-4 * gmax if j in (1, 2) else 0.0, 1e-30)
                err = float((a.reshape(b.shape) - b).abs().max()) / scale
                assert err < 1e-8, (it, name, world, rank, n, r, ns, block, schedule, j, err)
            for a, b in zip(results[rank], results[0]):
                assert torch.equal(a, b)          # every rank returns the same bits


@pytest.mark.parametrize('world,n,r,block,schedule', [(2, 300, 1, 128, True)] + (
    [(1, 330, 2, 128, 'v2'), (2, 300, 1, 128, 'v2'), (3, 420, 1, 128, True)] if FULL else []))
def test_distributed_gpr_on_the_real_kernels_under_gloo(cpu_lib, world, n, r, block, schedule):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_worker, args=(i, world, port, cpu_lib, n, r, block, q, schedule)) for i in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, errs in res:
        assert max(errs) < 1e-9, (rank, errs)

