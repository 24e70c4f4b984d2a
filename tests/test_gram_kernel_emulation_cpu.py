"""The SOURCE TEXT of the interpreter Gram kernels (csrc/gram.cu: feature_kernel,
gram_fwd_kernel, gram_bwd_kernel and the experimental gram_bwd_smem_kernel) executed on the CPU
and held to the oracle -- a check of the kernels' indexing, masking, program interpreter and
reductions that needs no GPU.

How: the region of gram.cu between `#include "internal.cuh"` and the first host launch function
(all kernels: interpreter, experimental shared-memory variants, register-tiled fast path, Kdiag) is copied
verbatim into a host translation unit between tests/emu/harness_prelude.h (CUDA keywords defined
away, one host thread per warp with its 32 lanes as cooperative fibers (emu/emu_fibers.h), __syncthreads and
warp shuffles as barriers) and tests/emu/harness_driver.inc (launch loops + the fixed-order
second-pass reductions of the host code), with exactly three textual substitutions:
    extern __shared__ double sm[];   ->  double* sm = emu_smem;
    __syncthreads()                  ->  emu_barrier()
    __shfl_xor_sync(0xffffffffu, v, m) -> emu_shfl_xor(v, m)
The kernel descriptors come from the package's own host code (`kern.program()`), so the chain
  Python kernel expression -> gps_kernel_desc -> CUDA interpreter source -> numbers
is tested end to end against torch autograd through the oracle's primitives.  What this cannot
see: anything hardware-specific (occupancy, shared-memory limits are checked by size only)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import cpu_ops_double
from oracle import cases

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GRAM_CU = os.path.join(ROOT, 'gpflow-slim_b200', 'csrc', 'gram.cu')


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    src = open(GRAM_CU).read()
    start = src.index('#include "internal.cuh"') + len('#include "internal.cuh"')
    end = src.index('int features(gps_handle* h')          # everything up to the host launch code
    region = src[start:end]
    # the tensor-core NKN kernels need the warp-level mma emulation of the whole-library CPU build
    # (tests/test_library_on_cpu.py::test_nkn_tensor_core_gram_kernels_against_the_interpreter covers them)
    n0, n1_ = region.index('// --------------------------------------------------------------------------- NKN fast path'), \
        region.index('// --------------------------------------------------------------------------- Kdiag')
    region = region[:n0] + region[n1_:]
    region, n1 = re.subn(r'extern __shared__ double sm\[\];', 'double* sm = emu_smem;', region)
    region, n2 = re.subn(r'__syncthreads\(\)', 'emu_barrier()', region)
    region, n3 = re.subn(r'__shfl_xor_sync\(0xffffffffu, (\w+), (\d+)\)', r'emu_shfl_xor(\1, \2)', region)
    assert n1 >= 3 and n2 >= 6 and n3 >= 8, (n1, n2, n3)
    assert '__shfl' not in region and '<<<' not in region
    d = tmp_path_factory.mktemp('gram_emu')
    tu = d / 'gram_emu.cpp'
    tu.write_text('#include "harness_prelude.h"\n' + region + '\n' +
                  open(os.path.join(HERE, 'emu', 'harness_driver.inc')).read())
    so = d / 'libgram_emu.so'
    cmd = ['g++', '-std=c++17', '-O1', '-fPIC', '-shared', '-pthread', '-Wno-attributes',
           '-I', os.path.join(HERE, 'emu'), '-I', os.path.join(ROOT, 'include'), str(tu), '-o', str(so)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
    lib = ctypes.CDLL(str(so))
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _gpf():
    import gpflowSlim as gpf
    gpf.settings.device = 'cpu'
    return gpf


def emu_fwd(lib, prog, theta, X, X2=None, diag_add=0.0, uplo=0, impl=1):
    N, M = X.shape[0], (X.shape[0] if X2 is None else X2.shape[0])
    K = np.full((N, M), np.nan)
    rc = lib.emu_gram_fwd(ctypes.byref(prog.desc), _ptr(theta), _ptr(X), ctypes.c_int64(N),
                          ctypes.c_int64(X.shape[1]), _ptr(X2), ctypes.c_int64(M), ctypes.c_double(diag_add),
                          ctypes.c_int(uplo), ctypes.c_int(impl), _ptr(K))
    assert rc == 0
    return K


def emu_bwd(lib, prog, theta, X, X2, W, impl, mode=0, beta=None, sym_lower=0, want_dx=False, njc=2):
    N, M = X.shape[0], (X.shape[0] if X2 is None else X2.shape[0])
    dth = np.full(prog.n_theta + 1, np.nan)
    dX = np.full(X.shape, np.nan) if want_dx else None
    R = 0 if beta is None else beta.shape[0]
    rc = lib.emu_gram_bwd(ctypes.byref(prog.desc), _ptr(theta), _ptr(X), ctypes.c_int64(N),
                          ctypes.c_int64(X.shape[1]), _ptr(X2), ctypes.c_int64(M), _ptr(W),
                          ctypes.c_int64(W.shape[1]), ctypes.c_int(mode), _ptr(beta), ctypes.c_int(R),
                          ctypes.c_int(sym_lower), ctypes.c_int(impl), ctypes.c_int(njc), _ptr(dth), _ptr(dX))
    assert rc == 0
    return dth, dX


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _torch_reference(prog, theta, X, X2, W):
    """sum(W * K) differentiated w.r.t. theta and the inputs through the oracle's primitives."""
    th = torch.tensor(theta, requires_grad=True)
    Xt = torch.tensor(X, requires_grad=True)
    X2t = None if X2 is None else torch.tensor(X2, requires_grad=True)
    K = cpu_ops_double._gram_desc(prog, th, Xt, X2t)
    val = (K * torch.tensor(W)).sum()
    leaves = [th, Xt] + ([X2t] if X2t is not None else [])
    g = torch.autograd.grad(val, leaves, allow_unused=True)
    g = [torch.zeros_like(l) if gi is None else gi for gi, l in zip(g, leaves)]
    return K.detach().numpy(), [gi.numpy() for gi in g]


@pytest.mark.parametrize('impl', [1, 2])
def test_interpreter_kernels_on_the_kernel_zoo(emu, impl):
    gpf = _gpf()
    rng = np.random.default_rng(7)
    X = rng.standard_normal((75, 3)) * 1.2            # 2 row tiles of 64 / 3 of 32, ragged
    X2 = rng.standard_normal((41, 3)) * 1.2
    W = rng.standard_normal((75, 41))
    Ws = rng.standard_normal((75, 75))
    for name, make in cases._kernel_zoo(gpf, 3):
        prog = make().program()
        theta = prog.theta('cpu').detach().numpy().copy()
        # cross-covariance: forward, dtheta, dX
        Kref, (gth, gx, gx2) = _torch_reference(prog, theta, X, X2, W)
        assert rel(emu_fwd(emu, prog, theta, X, X2, impl=impl), Kref) < 1e-12, name
        dth, dX = emu_bwd(emu, prog, theta, X, X2, W, impl, want_dx=True)
        assert rel(dth[:-1], gth) < 1e-10, (name, 'dtheta')
        assert rel(dX, gx) < 1e-10, (name, 'dX')
        # symmetric problem with a dense (symmetrised) weight: dX gets the factor 2
        Wsym = 0.5 * (Ws + Ws.T)
        Kref, (gth, gx) = _torch_reference(prog, theta, X, None, Wsym)
        Klow = emu_fwd(emu, prog, theta, X, None, diag_add=0.3, uplo=1, impl=impl)
        il = np.tril_indices(75)
        assert rel(Klow[il], (Kref + 0.3 * np.eye(75))[il]) < 1e-12, name
        assert np.isnan(Klow[np.triu_indices(75, 1)]).all(), 'upper triangle must stay untouched'
        dth, dX = emu_bwd(emu, prog, theta, X, None, Wsym, impl, want_dx=True, njc=1)
        # Matern-type diagonals carry sqrt(d2 + 1e-12) with d2 = +-1e-16 of rounding noise in BOTH
        # implementations (kernels.py:424-426): 1e-8 is the floor there, as in the GPU tests
        assert rel(dth[:-1], gth) < 1e-8, (name, 'sym dtheta')
        assert rel(dX, gx) < 1e-8, (name, 'sym dX')


@pytest.mark.parametrize('impl', [1, 2])
def test_interpreter_backward_nkn_fused_gpr_weights(emu, impl):
    """The C3 NKN topology with the fused-GPR weight mode: W = 1/2 (R K^-1 - beta beta^T) formed
    on the fly from the LOWER triangle of K^-1, off-diagonal elements counted twice, trace W as
    the extra accumulator (gpr.cu uses exactly this call)."""
    gpf = _gpf()
    d, n, R = 3, 83, 2
    rng = np.random.default_rng(11)
    X = rng.standard_normal((n, d))
    prog = cases.nkn_c3_kernel(gpf, d).program()
    theta = prog.theta('cpu').detach().numpy().copy()
    A = rng.standard_normal((n, n))
    Kinv = A @ A.T / n + np.eye(n)
    beta = rng.standard_normal((R, n))
    Wfull = 0.5 * (R * Kinv - beta.T @ beta)
    _, (gth, _) = _torch_reference(prog, theta, X, None, Wfull)
    Kref, _ = _torch_reference(prog, theta, X, None, Wfull)
    assert rel(emu_fwd(emu, prog, theta, X, None, impl=impl), Kref) < 1e-12
    Klow = np.tril(Kinv) + np.triu(np.full((n, n), np.nan), 1)       # upper triangle must not be read
    dth, _ = emu_bwd(emu, prog, theta, X, None, Klow, impl, mode=1, beta=beta, sym_lower=1, njc=2)
    assert rel(dth[:-1], gth) < 1e-10
    assert abs(dth[-1] - np.trace(Wfull)) < 1e-10 * abs(np.trace(Wfull))


def test_smem_variant_fits_the_c3_program():
    """Shared-memory budget of gram_bwd_smem_kernel for the BASELINE C3 topology at D=8
    (the formula of bwd_smem_doubles in gram.cu): must stay under the 227 KB per-CTA limit."""
    gpf = _gpf()
    prog = cases.nkn_c3_kernel(gpf, 8).program()
    d = prog.desc
    ft = 0
    for i in range(d.n_prims):
        pr = d.prims[i]
        ft += pr.ndims + 1 if pr.type <= 4 else (pr.ndims if pr.type == 5 else 3 * pr.ndims)
    S = ft | 1
    last = d.ops[d.n_ops - 1]
    nslots = last.dst + (last.n if last.op in (4, 5) else 1)
    doubles = d.n_theta + 2 * 32 * S + 2 * 1 * 32 + (d.n_theta + 1) * 128 + 2 * nslots * 128
    assert doubles * 8 <= 227 * 1024, doubles * 8


def emu_kdiag(lib, prog, theta, X, w=None, want_dx=False):
    N = X.shape[0]
    out = np.full(N, np.nan)
    dth = np.full(prog.n_theta, np.nan)
    dX = np.full(X.shape, np.nan) if want_dx else None
    rc = lib.emu_kdiag(ctypes.byref(prog.desc), _ptr(theta), _ptr(X), ctypes.c_int64(N),
                       ctypes.c_int64(X.shape[1]), _ptr(w), _ptr(out), _ptr(dth), _ptr(dX))
    assert rc == 0
    return out, dth, dX


def test_kdiag_kernel_on_the_kernel_zoo(emu):
    gpf = _gpf()
    rng = np.random.default_rng(9)
    X = rng.standard_normal((300, 3)) * 1.2                # two blocks of 256 threads, ragged
    w = rng.standard_normal(300)
    zoo = cases._kernel_zoo(gpf, 3) + [('nkn', lambda: cases.nkn_c3_kernel(gpf, 3))]
    for name, make in zoo:
        prog = make().program()
        theta = prog.theta('cpu').detach().numpy().copy()
        th = torch.tensor(theta, requires_grad=True)
        Xt = torch.tensor(X, requires_grad=True)
        kd = cpu_ops_double._kdiag_desc(prog, th, Xt)
        g = torch.autograd.grad((kd * torch.tensor(w)).sum(), [th, Xt], allow_unused=True)
        gth = np.zeros_like(theta) if g[0] is None else g[0].numpy()
        gx = np.zeros_like(X) if g[1] is None else g[1].numpy()
        out, _, _ = emu_kdiag(emu, prog, theta, X)
        assert rel(out, kd.detach().numpy()) < 1e-13, name
        _, dth, dX = emu_kdiag(emu, prog, theta, X, w=w, want_dx=True)
        assert rel(dth, gth) < 1e-11, (name, 'dtheta')
        assert np.abs(dX - gx).max() <= 1e-11 * max(np.abs(gx).max(), 1.0), (name, 'dX')


@pytest.mark.parametrize('cls,d,ard', [('RBF', 1, False), ('RBF', 8, True), ('Matern32', 3, True),
                                       ('Matern52', 11, True), ('Exponential', 5, False),
                                       ('Matern12', 16, True)])
def test_fast_path_kernels_equal_interpreter_and_oracle(emu, cls, d, ard):
    """gram_fwd_stat_kernel<D> / gram_bwd_stat_kernel<D> (the kernels of every BASELINE GPR / SVGP
    config except the NKN one), all three template widths, dense and fused-GPR weight modes."""
    gpf = _gpf()
    rng = np.random.default_rng(d * 7 + ard)
    n, m, R = 150, 70, 2
    X, X2 = rng.standard_normal((n, d)), rng.standard_normal((m, d))
    ls = (0.8 + 0.1 * np.arange(d)) * np.sqrt(d) if ard else 1.1 * np.sqrt(d)
    prog = getattr(gpf.kernels, cls)(d, variance=1.3, lengthscales=ls, ARD=ard).program()
    theta = prog.theta('cpu').detach().numpy().copy()
    W = rng.standard_normal((n, m))
    Kref, (gth, _, _) = _torch_reference(prog, theta, X, X2, W)
    assert rel(emu_fwd(emu, prog, theta, X, X2, impl=0), Kref) < 1e-12
    dth, _ = emu_bwd(emu, prog, theta, X, X2, W, 0)
    assert rel(dth[:-1], gth) < 1e-10
    # symmetric, lower triangle only, + diag_add
    Ksym, _ = _torch_reference(prog, theta, X, None, np.zeros((n, n)))
    Klow = emu_fwd(emu, prog, theta, X, None, diag_add=0.2, uplo=1, impl=0)
    il = np.tril_indices(n)
    # sqrt(d2 + 1e-12) with d2 = +-1e-16 of rounding noise on the diagonal of the Matern family
    assert rel(Klow[il], (Ksym + 0.2 * np.eye(n))[il]) < (1e-12 if cls == 'RBF' else 2e-9)
    assert np.isnan(Klow[np.triu_indices(n, 1)]).all()
    # fused-GPR weights from the lower triangle of K^-1
    A = rng.standard_normal((n, n))
    Kinv = A @ A.T / n + np.eye(n)
    beta = rng.standard_normal((R, n))
    Wfull = 0.5 * (R * Kinv - beta.T @ beta)
    _, (gth, _) = _torch_reference(prog, theta, X, None, Wfull)
    Kl = np.tril(Kinv) + np.triu(np.full((n, n), np.nan), 1)
    dth, _ = emu_bwd(emu, prog, theta, X, None, Kl, 0, mode=1, beta=beta, sym_lower=1, njc=2)
    tol = 1e-10 if cls == 'RBF' else 1e-8          # sqrt(d2 + 1e-12) noise on Matern-type diagonals
    assert rel(dth[:-1], gth) < tol
    assert abs(dth[-1] - np.trace(Wfull)) < 1e-10 * abs(np.trace(Wfull))

