"""The SOURCE TEXT of the Cholesky leaf kernels (csrc/potrf.cu: potrf_leaf_kernel -- the blocked
FP64 tensor-core 128 x 128 factor-and-invert kernel with in-CTA look-ahead --, potrf_base_kernel,
and the DMMA strip TRSM) executed on the CPU and held to LAPACK.

Same mechanism as tests/test_gram_kernel_emulation_cpu.py, with the warp-level layer of
tests/emu/harness_prelude_warp.h on top: per-warp barriers for __syncwarp / __shfl_sync, and an
emulation of mma.sync.m8n8k4.f64 (`dmma884`) that follows the PTX fragment layout, so the
kernels' fragment indexing is what gets tested.  Substitutions on the kernel text:
    extern __shared__ __align__(16) double sm[];  ->  double* sm = emu_smem;
    __syncthreads()                               ->  emu_barrier()
    __shfl_xor_sync(0xffffffffu, v, m)            ->  emu_shfl_xor_w(v, m)
The region copied is everything of potrf.cu's first anonymous namespace up to the host launch
helpers (constants, both leaf kernels, factor8 / inv8, the strip TRSM)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
POTRF_CU = os.path.join(ROOT, 'gpflow-slim_b200', 'csrc', 'potrf.cu')
NB = 128


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    src = open(POTRF_CU).read()
    start = src.index('#include "internal.cuh"') + len('#include "internal.cuh"')
    end = src.index('void set_attrs(gps_handle* h) {')
    region = src[start:end]
    region = region.replace('constexpr int NB = GPS_NB;', 'constexpr int NB = 128;')
    region, n1 = re.subn(r'extern __shared__ __align__\(16\) double sm\[\];', 'double* sm = emu_smem;', region)
    region, n2 = re.subn(r'__syncthreads\(\)', 'emu_barrier()', region)
    region, n3 = re.subn(r'__shfl_xor_sync\(0xffffffffu, (\w+), (\d+)\)', r'emu_shfl_xor_w(\1, \2)', region)
    assert n1 == 3 and n2 >= 10 and n3 >= 2, (n1, n2, n3)
    assert '<<<' not in region
    d = tmp_path_factory.mktemp('potrf_emu')
    tu = d / 'potrf_emu.cpp'
    tu.write_text('#include "harness_prelude_warp.h"\n#define warp_sum warp_sum_blockwide_unused\n'
                  '#undef warp_sum\n'
                  'static inline double warp_sum_w(double v) { for (int o = 16; o > 0; o >>= 1) v += emu_shfl_xor_w(v, o); return v; }\n'
                  '#define warp_sum warp_sum_w\n' + region + '\n' +
                  open(os.path.join(HERE, 'emu', 'potrf_driver.inc')).read())
    so = d / 'libpotrf_emu.so'
    cmd = ['g++', '-std=c++17', '-O1', '-fPIC', '-shared', '-pthread', '-Wno-attributes', '-Wno-unused-function',
           '-I', os.path.join(HERE, 'emu'), '-I', os.path.join(ROOT, 'include'), str(tu), '-o', str(so)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-6000:]
    return ctypes.CDLL(str(so))


def _p(a, t=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n + 5))
    return A @ A.T / (n + 5) + 0.5 * np.eye(n)


def run_blocks(lib, A, do_chol, impl, want_u=True):
    n = A.shape[0]
    nblk = (n + NB - 1) // NB
    A = np.ascontiguousarray(A)
    tinv = np.full((nblk, NB, NB), np.nan)
    logdet = np.full(nblk, np.nan)
    info = np.zeros(1, dtype=np.int32)
    U = np.full((n, n), np.nan) if want_u else None
    rc = lib.emu_potrf_blocks(_p(A), ctypes.c_int64(n), ctypes.c_int(n), ctypes.c_int(do_chol), ctypes.c_int(impl),
                              _p(tinv), _p(logdet), _p(info, ctypes.c_int), ctypes.c_int(7), _p(U),
                              ctypes.c_int64(n))
    assert rc == 0
    return A, tinv, logdet, int(info[0]), U


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('n', [128, 300, 77])
def test_leaf_cholesky_and_inverse_of_diagonal_blocks(emu, impl, n):
    S = _spd(n, n + impl)
    marker = 123.456
    A0 = np.tril(S) + np.triu(np.full((n, n), marker), 1)       # the strict upper triangle is not the kernel's
    A, tinv, logdet, info, U = run_blocks(emu, A0.copy(), 1, impl)
    assert info == 0
    for b in range((n + NB - 1) // NB):
        lo, hi = b * NB, min(n, (b + 1) * NB)
        m = hi - lo
        L = np.linalg.cholesky(S[lo:hi, lo:hi])
        got = A[lo:hi, lo:hi]
        np.testing.assert_allclose(np.tril(got), L, rtol=0, atol=1e-12 * np.abs(L).max())
        assert (got[np.triu_indices(m, 1)] == marker).all(), 'strict upper triangle must stay untouched'
        T = np.linalg.inv(L)
        np.testing.assert_allclose(tinv[b][:m, :m], T, rtol=0, atol=1e-10 * np.abs(T).max())
        assert np.abs(np.triu(tinv[b], 1)).max() == 0.0, 'T must carry explicit zeros above the diagonal'
        if m < NB:    # identity padding of a partial last block
            np.testing.assert_allclose(tinv[b][m:, m:], np.eye(NB - m), atol=1e-14)
            assert np.abs(tinv[b][m:, :m]).max() < 1e-14
        np.testing.assert_allclose(U[lo:hi, lo:hi], T.T, rtol=0, atol=1e-10 * np.abs(T).max())
        assert abs(logdet[b] - np.log(np.diag(L)).sum()) < 1e-11 * max(1.0, abs(logdet[b]))


@pytest.mark.parametrize('impl', [0, 1])
def test_leaf_reports_the_first_bad_pivot(emu, impl):
    n = 200
    S = _spd(n, 5)
    S[150, 150] = -1.0                      # leading minor of order 151 is not positive definite
    _, _, _, info, _ = run_blocks(emu, np.tril(S), 1, impl, want_u=False)
    assert info == 7 + 151                  # row0 + 1-based order


@pytest.mark.parametrize('impl', [0, 1])
def test_block_inverses_of_a_factored_matrix(emu, impl):
    n = 260
    L = np.linalg.cholesky(_spd(n, 9))
    _, tinv, _, _, _ = run_blocks(emu, L.copy(), 0, impl, want_u=False)
    for b in range(3):
        lo, hi = b * NB, min(n, (b + 1) * NB)
        T = np.linalg.inv(L[lo:hi, lo:hi])
        np.testing.assert_allclose(tinv[b][:hi - lo, :hi - lo], T, rtol=0, atol=1e-10 * np.abs(T).max())


@pytest.mark.parametrize('notrans', [0, 1])
@pytest.mark.parametrize('m,n', [(150, 128), (64, 128), (70, 44)])
def test_strip_trsm_on_the_tensor_core_emulation(emu, notrans, m, n):
    rng = np.random.default_rng(m + n + notrans)
    L = np.linalg.cholesky(_spd(n, 3))
    T = np.zeros((NB, NB))
    T[:n, :n] = np.linalg.inv(L)
    T[n:, n:] = np.eye(NB - n)
    B0 = rng.standard_normal((m, n + 3))                       # ld > n: columns beyond n stay untouched
    B = B0.copy()
    rc = emu.emu_trsm_strip(_p(B), ctypes.c_int64(B.shape[1]), ctypes.c_int(m), ctypes.c_int(n), _p(T),
                            ctypes.c_int(notrans))
    assert rc == 0
    want = B0[:, :n] @ (T[:n, :n] if notrans else T[:n, :n].T)
    np.testing.assert_allclose(B[:, :n], want, rtol=0, atol=1e-12 * np.abs(want).max())
    assert (B[:, n:] == B0[:, n:]).all()
