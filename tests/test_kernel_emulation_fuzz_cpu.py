"""Randomised differential tests on the CPU emulation of the CUDA kernels' source (tests/emu/):
random GEMM shapes / leading dimensions / alignments / triangular flags against numpy, and random
kernel EXPRESSIONS (primitives on random active dimensions, nested Sum / Product with constants,
small NKNs) through the interpreter Gram kernels -- validated and experimental variants --
against torch autograd through the oracle.  Seeded and bounded (100 cases per family, about ten
seconds on the fiber emulation); set GPSLIM_FUZZ=<n> for n cases per family (500 Gram and 280 GEMM
cases were run clean when this was written, 400 + 400 again on the final code of round 2; the only deviations seen were the known 1e-8-level rounding noise of Matern-type kernels
at nearly coincident 1-D points, which the reference's own distance formula has as well)."""
import os

import numpy as np
import pytest

import test_gemm_kernel_emulation_cpu as G
import test_gram_kernel_emulation_cpu as K

NCASES = int(os.environ.get('GPSLIM_FUZZ', '100'))

gemm_lib = G.emu          # module-scoped fixtures of the two emulation tests, re-exported
gram_lib = K.emu


def test_fuzz_gemm(gemm_lib):
    rng = np.random.default_rng(101)
    for it in range(NCASES):
        a_tri = b_tri = c_uplo = 0
        if rng.random() < 0.5:
            M = N = Kd = int(rng.integers(1, 280))
            a_tri, b_tri, c_uplo = int(rng.integers(0, 3)), int(rng.integers(0, 3)), int(rng.integers(0, 2))
            if rng.random() < 0.3:
                Kd, a_tri, b_tri = int(rng.integers(1, 60)), 0, 0
        else:
            M, N, Kd = int(rng.integers(1, 300)), int(rng.integers(1, 300)), int(rng.integers(1, 70))
        pa, pb, pc = (int(rng.integers(0, 4)) for _ in range(3))
        oa, ob = int(rng.integers(0, 2)), int(rng.integers(0, 2))            # 8-byte misalignment
        A = rng.standard_normal(M * (Kd + pa) + 1)[oa:oa + M * (Kd + pa)].reshape(M, Kd + pa)[:, :Kd]
        B = rng.standard_normal(N * (Kd + pb) + 1)[ob:ob + N * (Kd + pb)].reshape(N, Kd + pb)[:, :Kd]
        A[...] = G.tri(A.copy(), a_tri)
        B[...] = G.tri(B.copy(), b_tri)
        C0 = rng.standard_normal((M, N))
        beta, alpha = float(rng.choice([0.0, 1.0, -0.4])), float(rng.choice([1.0, -1.0, 0.6]))
        Cfull = np.full((M, N + pc), 777.0)
        Cfull[:, :N] = C0 if beta else np.nan
        got = G.gemm(gemm_lib, A, B, Cfull[:, :N], alpha=alpha, beta=beta, a_tri=a_tri, b_tri=b_tri, c_uplo=c_uplo)
        want = alpha * A @ B.T + (beta * C0 if beta else 0.0)
        tol = 5e-13 * max(1.0, np.abs(want).max())
        what = dict(it=it, M=M, N=N, K=Kd, a_tri=a_tri, b_tri=b_tri, c_uplo=c_uplo, pad=(pa, pb, pc), off=(oa, ob),
                    alpha=alpha, beta=beta)
        if c_uplo:
            assert np.abs(np.tril(got) - np.tril(want)).max() < tol, what
            iu = np.triu_indices(M, 1)
            assert (got[iu] == C0[iu]).all() if beta else np.isnan(got[iu]).all(), what
        else:
            assert np.abs(got - want).max() < tol, what
        assert (Cfull[:, N:] == 777.0).all(), what                          # padding columns untouched


def _random_kernel(gpf, rng, D):
    k = gpf.kernels

    def prim():
        nd = int(rng.integers(1, D + 1))
        dims = rng.permutation(D)[:nd].tolist()
        kind, ard = int(rng.integers(0, 7)), bool(rng.random() < 0.5)
        ls = (0.6 + rng.random(nd)) if ard else float(0.6 + rng.random())
        var, name = float(0.3 + rng.random()), 'f%d' % rng.integers(1e9)
        if kind <= 4:
            cls = [k.RBF, k.Matern12, k.Matern32, k.Matern52, k.Exponential][kind]
            return cls(nd, variance=var, lengthscales=ls, active_dims=dims, ARD=ard, name=name)
        if kind == 5:
            return k.Linear(nd, variance=(0.2 + rng.random(nd)) if ard else var, active_dims=dims, ARD=ard, name=name)
        return k.Periodic(nd, period=float(0.8 + 2 * rng.random()), variance=var,
                          lengthscales=float(0.5 + rng.random()), active_dims=dims, name=name)

    def expr(depth=0):
        if depth >= 2 or rng.random() < 0.35:
            return prim()
        kids = [expr(depth + 1) for _ in range(int(rng.integers(2, 4)))]
        add = rng.random() < 0.5
        out = kids[0]
        for c in kids[1:]:
            out = out + c if add else out * c
        if rng.random() < 0.4:
            out = out + float(rng.random()) if add else out * float(0.5 + rng.random())
        return out
    if rng.random() < 0.25:
        np.random.seed(int(rng.integers(1e6)))                               # NKN Linear weights: numpy's global RNG
        P, h1 = int(rng.integers(2, 5)), int(rng.choice([2, 4, 6]))
        hp = [dict(name='Linear', params=dict(input_dim=P, output_dim=h1, name='a')),
              dict(name='Product', params=dict(input_dim=h1, step=2, name='b')),
              dict(name='Linear', params=dict(input_dim=h1 // 2, output_dim=1, name='c'))]
        return gpf.neural_kernel_network.NeuralKernelNetwork(
            D, [prim() for _ in range(P)], gpf.neural_kernel_network.NKNWrapper(hp))
    return expr()


def test_fuzz_gram_interpreter(gram_lib):
    gpf = K._gpf()
    rng = np.random.default_rng(202)
    for it in range(NCASES):
        D = int(rng.integers(1, 7))
        prog = _random_kernel(gpf, rng, D).program()
        theta = prog.theta('cpu').detach().numpy().copy()
        N, M = int(rng.integers(1, 100)), int(rng.integers(1, 100))
        X, X2 = rng.standard_normal((N, D)) * 1.1, rng.standard_normal((M, D)) * 1.1
        W = rng.standard_normal((N, M))
        Kref, (gth, gx, _) = K._torch_reference(prog, theta, X, X2, W)
        R = int(rng.integers(1, 4))
        A = rng.standard_normal((N, N))
        Kinv, beta = A @ A.T / N + np.eye(N), rng.standard_normal((R, N))
        Wf = 0.5 * (R * Kinv - beta.T @ beta)
        _, (gsym, _) = K._torch_reference(prog, theta, X, None, Wf)
        Klow = np.tril(Kinv) + np.triu(np.full((N, N), np.nan), 1)
        for impl in (1, 2):
            what = (it, impl, D, N, M, prog.n_theta)
            assert K.rel(K.emu_fwd(gram_lib, prog, theta, X, X2, impl=impl), Kref) < 1e-11, what
            dth, dX = K.emu_bwd(gram_lib, prog, theta, X, X2, W, impl, want_dx=True, njc=int(rng.integers(1, 4)))
            assert np.abs(dth[:-1] - gth).max() < 1e-8 * max(np.abs(gth).max(), 1e-300), what
            assert np.abs(dX - gx).max() < 1e-7 * max(np.abs(gx).max(), 1e-12), what
            dth, _ = K.emu_bwd(gram_lib, prog, theta, X, None, Klow, impl, mode=1, beta=beta, sym_lower=1,
                               njc=int(rng.integers(1, 3)))
            assert np.abs(dth[:-1] - gsym).max() < 5e-8 * max(np.abs(gsym).max(), 1e-300), what
            assert abs(dth[-1] - np.trace(Wf)) < 1e-10 * max(1.0, abs(np.trace(Wf))), what
