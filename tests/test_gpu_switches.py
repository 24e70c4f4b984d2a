"""Switches of the library that have an alternative implementation behind gps_set_option / the ops
module flags, each held to the default path (or to torch) on the GPU.  Written blind at the end of
round 1, first run (all green) on a B200 at the start of round 2 (profiles/r02_experimental_switches_gpu.txt);
since then part of the regular `pytest -m gpu` run.

gram_impl = 0 on Linear/Product(2) neural-kernel networks: the tensor-core kernels gram_fwd_nkn_kernel /
gram_bwd_nkn_kernel (csrc/gram.cu) against the interpreter (gram_impl = 1), Gram, dense backward and the
fused GPR gradient; the timing test prints all three implementations.

gram_impl = 2: interpreter Gram forward / backward with the slot values, slot adjoints and
theta-gradient accumulators in shared memory ([index][thread] layout) instead of local memory
(csrc/gram.cu: gram_fwd_smem_kernel, gram_bwd_smem_kernel).  Must reproduce the default interpreter
(gram_impl = 1) to rounding: same per-element arithmetic, different summation order."""
import os

import numpy as np
import pytest
import torch

from oracle import cases
from util import assert_close, conv

pytestmark = [pytest.mark.gpu]


def _grads(kern, X, X2, W, Ws, impl, want_dx):
    from gpflowSlim._backend.lib import handle_for
    h = handle_for(X)
    h.set_option('gram_impl', impl)
    try:
        Xg = X.clone().requires_grad_(want_dx)
        X2g = X2.clone().requires_grad_(want_dx)
        val = (kern.K(Xg, X2g) * W).sum() + (kern.K(Xg) * Ws).sum()
        params = [p.unconstrained_tensor for p in kern.parameters]
        g = torch.autograd.grad(val, params + ([Xg, X2g] if want_dx else []), allow_unused=True)
        return [val.detach()] + [torch.zeros_like(p) if gi is None else gi
                                 for gi, p in zip(g, params + [Xg, X2g])]
    finally:
        h.set_option('gram_impl', 0)


@pytest.mark.parametrize('want_dx', [False, True])
@pytest.mark.parametrize('n,m', [(75, 41), (33, 97), (257, 130)])
def test_smem_accumulator_backward_equals_interpreter_on_the_zoo(n, m, want_dx):
    import gpflowSlim as gpf
    rng = np.random.default_rng(n + m)
    X, X2 = conv(rng.standard_normal((n, 3)) * 1.2), conv(rng.standard_normal((m, 3)) * 1.2)
    W, Ws = conv(rng.standard_normal((n, m))), conv(rng.standard_normal((n, n)))
    for name, make in cases._kernel_zoo(gpf, 3):
        kern = make()
        ref = _grads(kern, X, X2, W, Ws, 1, want_dx)
        got = _grads(kern, X, X2, W, Ws, 2, want_dx)
        for i, (a, b) in enumerate(zip(got, ref)):
            assert_close(a, b, 1e-11, '%s grad %d (n=%d m=%d dx=%s)' % (name, i, n, m, want_dx))


def test_smem_accumulator_backward_nkn_gpr(golden):
    """The NKN config through the fused GPR objective (W_GPR weights formed on the fly)."""
    import gpflowSlim as gpf
    from gpflowSlim._backend.lib import handle_for
    gold = golden('nkn')
    d, n = 3, 150
    X, Y = cases.synth_gpr(n, d, seed=3)
    kern = cases.nkn_c3_kernel(gpf, d)
    m = gpf.models.GPR(conv(X), conv(Y), kern=kern, name='nkn_gpr')
    h = handle_for(m.X)
    h.set_option('gram_impl', 2)
    try:
        obj = m.objective
        gs = torch.autograd.grad(obj, [p.unconstrained_tensor for p in m.parameters])
    finally:
        h.set_option('gram_impl', 0)
    assert_close(obj, gold['objective'], 1e-8, 'objective')
    for i, g in enumerate(gs):
        assert_close(g, gold['grad/objective/%d' % i], 1e-8, 'grad %d' % i)


@pytest.mark.parametrize('n,m', [(75, 41), (300, 517), (1024, 64)])
def test_nkn_tensor_core_kernels_equal_the_interpreter(n, m):
    """gram_impl 0 on a Linear/Product(2) network (gram_fwd_nkn_kernel / gram_bwd_nkn_kernel: layers,
    adjoint mat-vecs and weight gradients as FP64 tensor-core products over octets of elements) against
    the interpreter (gram_impl 1): K(X), K(X, X2), dense backward of both, the fused GPR gradient."""
    import gpflowSlim as gpf
    from gpflowSlim._backend.lib import handle_for
    d = 5
    rng = np.random.default_rng(n * 7 + m)
    X, X2 = conv(rng.standard_normal((n, d))), conv(rng.standard_normal((m, d)))
    Y = conv(rng.standard_normal((n, 2)))
    W, Ws = conv(rng.standard_normal((n, m))), conv(rng.standard_normal((n, n)))
    h = handle_for(X)
    res = {}
    for impl in (0, 1):
        h.set_option('gram_impl', impl)
        try:
            kern = cases.nkn_c3_kernel(gpf, d)
            params = [p.unconstrained_tensor for p in kern.parameters]
            K, K2 = kern.K(X), kern.K(X, X2)
            g1 = torch.autograd.grad((K * Ws).sum(), params)
            g2 = torch.autograd.grad((K2 * W).sum(), params)
            model = gpf.models.GPR(X, Y, kern=kern, name='nkn_tc_%d_%d_%d' % (n, m, impl))
            obj = model.objective
            g3 = torch.autograd.grad(obj, [p.unconstrained_tensor for p in model.parameters])
            res[impl] = [K.detach(), K2.detach(), obj.detach()] + list(g1) + list(g2) + list(g3)
        finally:
            h.set_option('gram_impl', 0)
    same = True
    for i, (a, b) in enumerate(zip(res[0], res[1])):
        assert_close(a, b, 1e-9, 'nkn item %d (n=%d m=%d)' % (i, n, m))
        same = same and torch.equal(a, b)
    assert not same            # the two runs took different kernels


def test_smem_accumulator_backward_speed_nkn():
    """Not an assertion about speed -- prints the two timings for the NKN Gram backward."""
    import gpflowSlim as gpf
    from gpflowSlim._backend.lib import handle_for
    n, d = 4096, 8
    X, _ = cases.synth_gpr(n, d, seed=0)
    kern = cases.nkn_c3_kernel(gpf, d)
    Xc = conv(X)
    W = conv(np.random.default_rng(0).standard_normal((n, n)))
    h = handle_for(Xc)
    for impl in (0, 1, 2):     # 0: the tensor-core NKN kernels, 1: interpreter, 2: interpreter with shared-memory arrays
        h.set_option('gram_impl', impl)
        try:
            params = [p.unconstrained_tensor for p in kern.parameters]
            for it in range(3):
                if it == 1:
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                val = (kern.K(Xc) * W).sum()
                torch.autograd.grad(val, params)
            e1.record()
            torch.cuda.synchronize()
            print('gram_impl=%d: NKN K + backward at N=%d: %.2f ms' % (impl, n, e0.elapsed_time(e1) / 2))
        finally:
            h.set_option('gram_impl', 0)


def test_graphed_svgp_step_equals_eager_steps():
    """training.GraphedStep (CUDA-graph replay of objective + gradients + Adam) against the same
    steps taken eagerly with AdamOptimizer, five minibatches; then the two step times."""
    import gpflowSlim as gpf
    n, d, m, batch = 20000, 8, 256, 2048
    X, Y, Z = cases.synth_svgp(n, d, m, seed=0)
    Xd, Yd = conv(X), conv(Y)

    def make():
        kern = gpf.kernels.RBF(d, ARD=True, lengthscales=2.0)
        return gpf.models.SVGP(Xd[:batch], Yd[:batch], kern, gpf.likelihoods.Gaussian(var=0.1), Z=Z.copy(),
                               num_data=n)
    batches = [(Xd[i * batch:(i + 1) * batch], Yd[i * batch:(i + 1) * batch]) for i in range(5)]
    eager = make()
    opt = gpf.training.AdamOptimizer(1e-3)
    objs_e = []
    for Xb, Yb in batches:
        eager.X, eager.Y = Xb, Yb
        objs_e.append(float(opt.minimize(eager)))
    graphed = make()
    step = gpf.training.GraphedStep(graphed, batches[0][0], batches[0][1], learning_rate=1e-3)
    objs_g = [float(step(Xb, Yb)) for Xb, Yb in batches]
    np.testing.assert_allclose(objs_g, objs_e, rtol=1e-9)
    for a, b in zip(graphed.trainable_tensors, eager.trainable_tensors):
        assert_close(a, b, 1e-9, 'parameters after 5 steps')

    def timed(fn, reps=20):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    Xb, Yb = batches[0]

    def eager_step():
        eager.X, eager.Y = Xb, Yb
        opt.minimize(eager)
    print('SVGP step (M=%d, B=%d): eager %.3f ms, graphed %.3f ms' % (m, batch, timed(eager_step),
                                                                       timed(lambda: step(Xb, Yb))))


@pytest.mark.parametrize('fused,tri_aware', [(False, False), (False, True), (True, False)])
@pytest.mark.parametrize('name', ['svgp_white_full', 'svgp_nonwhite_full', 'sgpr', 'functions'])
def test_adjoint_switches_match_reference_golden(golden, name, fused, tri_aware, monkeypatch):
    """ops.FUSED_ADJOINTS / ops.TRI_AWARE_ADJOINTS are both on by default (that combination is what
    tests/test_gpu_parity.py runs); the other three combinations -- the composed adjoint formulas,
    and products whose adjoints do not skip the zero tiles -- must give the same gradients."""
    import gpflowSlim as gpf
    from gpflowSlim._backend import ops
    from util import relerr
    monkeypatch.setattr(ops, 'TRI_AWARE_ADJOINTS', [tri_aware])
    monkeypatch.setattr(ops, 'FUSED_ADJOINTS', [fused])
    gold = golden(name)
    res = cases.run_case(gpf, name, conv)
    for key in sorted(gold):
        e = relerr(res[key], gold[key])
        assert e < (1e-12 if key.startswith('param/') else 1e-8), '%s:%s %.3e' % (name, key, e)


def test_library_side_adjoints_on_the_gpu(golden, monkeypatch):
    """gps_potri / gps_chol_bwd / gps_trsm_bwd (csrc/adjoint.cu) against torch autograd, then two
    golden cases with ops.FUSED_ADJOINTS on."""
    import gpflowSlim as gpf
    from gpflowSlim._backend import ops
    from util import relerr
    rng = np.random.default_rng(3)
    n, m = 700, 300
    A = rng.standard_normal((n, n + 3))
    S = conv(A @ A.T / (n + 3) + 0.5 * np.eye(n)).requires_grad_(True)
    B = conv(rng.standard_normal((m, n))).requires_grad_(True)
    Lbar_in, Xbar_in = conv(rng.standard_normal((n, n))), conv(rng.standard_normal((m, n)))
    L = torch.linalg.cholesky(S)
    X = torch.linalg.solve_triangular(L, B.t(), upper=False).t()
    (want_A,) = torch.autograd.grad((L * torch.tril(Lbar_in)).sum(), [S], retain_graph=True)
    want_B, = torch.autograd.grad((X * Xbar_in).sum(), [B], retain_graph=True)
    Ld = L.detach()
    for U in (None, ops.tri_inv_t(Ld)):
        assert_close(ops.chol_bwd(Ld, Lbar_in, U), 0.5 * (want_A + want_A.t()), 1e-10, 'chol_bwd')
        assert_close(ops.trsm_bwd(Ld, X.detach(), Xbar_in, U)[0], want_B, 1e-10, 'trsm_bwd Bbar')
    assert_close(ops.potri(Ld), torch.tril(torch.linalg.inv(S.detach())), 1e-9, 'potri')
    monkeypatch.setattr(ops, 'FUSED_ADJOINTS', [True])
    for name in ('svgp_white_full', 'sgpr'):
        gold = golden(name)
        res = cases.run_case(gpf, name, conv)
        for key in sorted(gold):
            e = relerr(res[key], gold[key])
            assert e < (1e-12 if key.startswith('param/') else 1e-8), '%s:%s %.3e' % (name, key, e)


def test_split_k_gemm_on_the_gpu():
    """Split-K (on by default): products with a long K and few output tiles -- the 1024 x 1024 x 8192
    products of the SVGP backward (64 / 36 output tiles on 148 SMs), the M x 1 products with K = 8192
    -- sliced along K in BOTH tensor-core kernels (TMA: gemm_impl 0, cp.async: 2); correctness
    against torch incl. alpha / beta, lower output and ragged shapes, and the timings."""
    from gpflowSlim._backend import ops
    from gpflowSlim._backend.lib import handle_for
    rng = np.random.default_rng(0)
    A, B = conv(rng.standard_normal((1024, 8192))), conv(rng.standard_normal((1024, 8192)))
    C0 = conv(rng.standard_normal((1024, 1024)))
    v = conv(rng.standard_normal((1, 8192)))
    Ar, Br = conv(rng.standard_normal((300, 5000))), conv(rng.standard_normal((170, 5000)))
    h = handle_for(A)
    want = A @ B.t()

    def timed(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    try:
        for impl in (0, 2):
            h.set_option('gemm_impl', impl)
            h.set_option('gemm_splitk', 0)
            t_plain = timed(lambda: ops.gemm_nt(A, B))
            t_low = timed(lambda: ops.gemm_nt(A, A, c_uplo=1))
            t_vec = timed(lambda: ops.gemm_nt(A, v))
            h.set_option('gemm_splitk', 1)
            assert_close(ops.gemm_nt(A, B), want, 1e-12, 'split-K')
            assert_close(ops.gemm_nt(A, B, alpha=-0.5, beta=2.0, out=C0.clone()), 2.0 * C0 - 0.5 * want, 1e-12,
                         'split-K rmw')
            assert_close(ops.gemm_nt(A, A, c_uplo=1), torch.tril(A @ A.t()), 1e-12, 'split-K lower')
            assert_close(ops.gemm_nt(A, v), A @ v.t(), 1e-12, 'split-K M x 1')
            assert_close(ops.gemm_nt(v, A), v @ A.t(), 1e-12, 'split-K 1 x M')
            assert_close(ops.gemm_nt(Ar, Br), Ar @ Br.t(), 1e-12, 'split-K ragged')
            assert_close(ops.gemm_nt(Ar, Ar, c_uplo=1), torch.tril(Ar @ Ar.t()), 1e-12, 'split-K ragged lower')
            t_split = timed(lambda: ops.gemm_nt(A, B))
            t_split_low = timed(lambda: ops.gemm_nt(A, A, c_uplo=1))
            t_split_vec = timed(lambda: ops.gemm_nt(A, v))
            print('gemm_impl %d  1024 x 1024 x 8192: %.3f ms -> split-K %.3f ms;  lower output: %.3f -> %.3f ms;  '
                  '1024 x 1 x 8192: %.3f -> %.3f ms' % (impl, t_plain, t_split, t_low, t_split_low, t_vec, t_split_vec))
    finally:
        h.set_option('gemm_impl', 0)
        h.set_option('gemm_splitk', 1)
