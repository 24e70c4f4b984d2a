"""TEST INFRASTRUCTURE (oracle): GPR ARD-RBF objective + analytic gradient at the FULL BASELINE sizes
(C2: N=8192, C5: N=32768, D=8) on the CPU, written to tests/golden/gpr_large_scalars.json.

Same algebra as `oracle.ref_torch.gpr_nlml_grad_lapack` (reference call sites: models/gpr.py:55-72,
densities.py:73-95, kernels.py:408-439; gradient = 1/2 tr((R K^-1 - beta beta^T) dK/dtheta), the
identity TensorFlow's autodiff of tf.cholesky evaluates), restated so that at most THREE N x N
buffers are alive (LAPACK potrf / potri through torch-CPU, Gram and gradient contraction by row
blocks): at N=32768 one matrix is 8.6 GB.  Checked against `gpr_nlml_grad_lapack` and against the torch-autograd
oracle at N=2048 before the large sizes run.  `bench.py` and tests/test_gpu_large_parity.py hold the
CUDA path to these scalars at 1e-8 relative.

    python oracle/gen_large_golden.py [--sizes 8192,32768]

Inputs: bench.synth_gpr(N, 8, seed 0); variance 1, lengthscales sqrt(8), noise 0.1 (the bench
configuration).  Gradients are w.r.t. the CONSTRAINED (variance, lengthscales[8], noise)."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LOG2PI = math.log(2.0 * math.pi)


def synth_gpr(n, d, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    Y = np.sin(X.sum(1, keepdims=True) / np.sqrt(d)) + 0.1 * rng.standard_normal((n, 1))
    return X, Y


def gram_block(Xs, s, r0, r1, variance):
    d2 = -2.0 * Xs[r0:r1] @ Xs.T
    d2 += s[r0:r1, None]
    d2 += s[None, :]
    np.maximum(d2, 0.0, out=d2)            # kernels.py:415 clip
    d2 *= -0.5
    np.exp(d2, out=d2)
    d2 *= variance
    return d2


def gpr_large(X, Y, variance, ls, noise, block=2048):
    from scipy.linalg import lapack
    N, R = Y.shape
    D = X.shape[1]
    ls = np.asarray(ls, dtype=np.float64) * np.ones(D)
    Xs = X / ls
    s = (Xs ** 2).sum(1)
    t0 = time.perf_counter()
    A = np.empty((N, N))
    for r0 in range(0, N, block):
        r1 = min(N, r0 + block)
        A[r0:r1] = gram_block(Xs, s, r0, r1, variance)
    A[np.arange(N), np.arange(N)] += noise
    t_gram = time.perf_counter() - t0
    # LAPACK through torch (MKL): scipy's bundled OpenBLAS fails beyond 2^32 bytes per matrix
    # (dpotrf reports a non-positive pivot at row 16545 of the N=32768 matrix, i.e. right after
    # the 4 GiB mark; the same matrix factors fine with MKL and on the GPU)
    import torch
    t0 = time.perf_counter()
    At = torch.from_numpy(A)
    Lt, info = torch.linalg.cholesky_ex(At)
    assert int(info) == 0, 'potrf info %d' % int(info)
    t_potrf = time.perf_counter() - t0
    del At, A
    Yt = torch.from_numpy(np.ascontiguousarray(Y))
    alpha = torch.linalg.solve_triangular(Lt, Yt, upper=False)
    nlml = float(0.5 * N * R * LOG2PI + R * torch.log(torch.diagonal(Lt)).sum() + 0.5 * (alpha ** 2).sum())
    beta = torch.linalg.solve_triangular(Lt.T, alpha, upper=True).numpy()
    t0 = time.perf_counter()
    A = torch.cholesky_inverse(Lt).numpy()            # full symmetric K^-1
    del Lt
    t_potri = time.perf_counter() - t0
    t0 = time.perf_counter()
    g_var = 0.0
    g_noise = 0.0
    g_ls = np.zeros(D)
    for r0 in range(0, N, block):
        r1 = min(N, r0 + block)
        Kinv = A[r0:r1]
        W = 0.5 * (R * Kinv - beta[r0:r1] @ beta.T)
        g_noise += np.trace(W[:, r0:r1])
        Kf = gram_block(Xs, s, r0, r1, variance)
        W *= Kf                                     # W o K
        g_var += W.sum() / variance
        for d in range(D):
            diff2 = (X[r0:r1, d:d + 1] - X[None, :, d]) ** 2
            g_ls[d] += (W * diff2).sum() / ls[d] ** 3
    t_grad = time.perf_counter() - t0
    return dict(nlml=float(nlml), g_variance=float(g_var), g_lengthscales=[float(v) for v in g_ls],
                g_noise=float(g_noise),
                seconds=dict(gram=t_gram, potrf=t_potrf, potri=t_potri, contraction=t_grad,
                             total=t_gram + t_potrf + t_potri + t_grad))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default='8192,32768')
    ap.add_argument('--out', default=os.path.join(ROOT, 'tests', 'golden', 'gpr_large_scalars.json'))
    args = ap.parse_args()
    d = 8
    variance, ls, noise = 1.0, math.sqrt(d), 0.1
    # pin the lean restatement to the oracle first
    import torch
    from oracle import ref_torch as R
    X, Y = synth_gpr(2048, d)
    got = gpr_large(X, Y, variance, ls, noise, block=512)
    ref = R.gpr_nlml_grad_lapack(X, Y, variance, ls, noise)
    th = [torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in (variance, ls * np.ones(d), noise)]
    o = R.gpr_nlml(dict(type='rbf', variance=th[0], lengthscales=th[1]), torch.tensor(X), torch.tensor(Y), th[2])
    ga = torch.autograd.grad(o, th)
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(np.asarray(b))))
    checks = dict(nlml_vs_lapack=rel(got['nlml'], ref[0]), nlml_vs_autograd=rel(got['nlml'], float(o)),
                  gvar_vs_autograd=rel(got['g_variance'], float(ga[0])),
                  gls_vs_autograd=rel(got['g_lengthscales'], ga[1].numpy()),
                  gnoise_vs_autograd=rel(got['g_noise'], float(ga[2])))
    print('self-check at N=2048 (relative):', checks, flush=True)
    assert max(checks.values()) < 1e-10, checks
    out = {'config': dict(d=d, variance=variance, lengthscales=ls, noise=noise, data='bench.synth_gpr(N, 8, seed=0)'),
           'self_check_n2048': checks, 'threads': os.cpu_count(), 'cases': {}}
    if os.path.exists(args.out):
        out['cases'] = json.load(open(args.out)).get('cases', {})
    for n in [int(v) for v in args.sizes.split(',')]:
        X, Y = synth_gpr(n, d)
        res = gpr_large(X, Y, variance, ls, noise)
        print(n, res, flush=True)
        out['cases'][str(n)] = res
        json.dump(out, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
