"""The SOURCE TEXT of the cp.async FP64 tensor-core GEMM (csrc/gemm.cu: gemm_nt_dmma_kernel with
tile_coords / k_range / load_stage / gemm_epilogue -- the tile order, the triangular K ranges, the
ragged-edge zero fill and ALL epilogue variants are shared with the TMA kernel, whose main loop
alone cannot be emulated) executed on the CPU with the mma.sync emulation of
tests/emu/harness_prelude_warp.h and held to numpy.

Covers every (a_tri, b_tri, c_uplo) combination -- including the ones the opt-in
triangular-aware adjoints (ops.TRI_AWARE_ADJOINTS) introduce --, beta != 0 read-modify-write,
16-byte and 8-byte staging, the block-row `rowmap` mask and the two per-row-start modes of the
distributed path.  Substitutions on the kernel text:
    extern __shared__ __align__(16) double smem[];   ->  double* smem = emu_smem;
    __syncthreads()                                  ->  emu_barrier()
    asm volatile("prefetch.global.L2 ...")           ->  (void)0        (a hint, no semantics)
    __ldcg(p)                                        ->  *(p)           (a cache hint)"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GEMM_CU = os.path.join(ROOT, 'gpflow-slim_b200', 'csrc', 'gemm.cu')
TRI_NONE, TRI_LOWER, TRI_UPPER = 0, 1, 2
C_ALL, C_LOWER, C_ROWMAP = 0, 1, 2


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    src = open(GEMM_CU).read()
    start = src.index('#include "internal.cuh"') + len('#include "internal.cuh"')
    end = src.index('// ------------------------------------------------------------------ TMA + mbarrier variant')
    region = src[start:end]
    region, n1 = re.subn(r'extern __shared__ __align__\(16\) double smem\[\];', 'double* smem = emu_smem;', region)
    region, n2 = re.subn(r'__syncthreads\(\)', 'emu_barrier()', region)
    region, n3 = re.subn(r'asm volatile\("prefetch\.global\.L2.*?\)\);', '(void)0;', region, flags=re.S)
    region, n4 = re.subn(r'__ldcg\(', '*(', region)
    assert (n1, n3, n4) == (1, 1, 1) and n2 >= 1, (n1, n2, n3, n4)
    assert 'asm' not in region and '<<<' not in region
    d = tmp_path_factory.mktemp('gemm_emu')
    tu = d / 'gemm_emu.cpp'
    tu.write_text('#include "harness_prelude_warp.h"\n'
                  'enum { TRI_NONE = 0, TRI_LOWER = 1, TRI_UPPER = 2 };\n'
                  'enum { C_ALL = 0, C_LOWER = 1, C_ROWMAP = 2 };\n' + region + '\n' +
                  open(os.path.join(HERE, 'emu', 'gemm_driver.inc')).read())
    so = d / 'libgemm_emu.so'
    cmd = ['g++', '-std=c++17', '-O1', '-fPIC', '-shared', '-pthread', '-Wno-attributes', '-Wno-unused-function',
           '-I', os.path.join(HERE, 'emu'), '-I', os.path.join(ROOT, 'include'), str(tu), '-o', str(so)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-6000:]
    return ctypes.CDLL(str(so))


def _p(a, t=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def gemm(lib, A, B, C, alpha=1.0, beta=0.0, a_tri=0, b_tri=0, c_uplo=0, rowlim=None, coff=0, rowlo=None,
         lo_off=0, lo_mode=0, vec16=None):
    M, K = A.shape
    N = B.shape[0]
    lda, ldb, ldc = A.strides[0] // 8, B.strides[0] // 8, C.strides[0] // 8
    if vec16 is None:
        vec16 = int(lda % 2 == 0 and ldb % 2 == 0 and A.ctypes.data % 16 == 0 and B.ctypes.data % 16 == 0)
    rc = lib.emu_gemm_nt(ctypes.c_double(alpha), _p(A), ctypes.c_int64(lda), _p(B), ctypes.c_int64(ldb),
                         ctypes.c_double(beta), _p(C), ctypes.c_int64(ldc), M, N, K, a_tri, b_tri, c_uplo,
                         _p(rowlim, ctypes.c_int64), ctypes.c_int64(coff), _p(rowlo, ctypes.c_int64),
                         ctypes.c_int64(lo_off), lo_mode, vec16)
    assert rc == 0
    return C


def tri(X, mode):
    return np.tril(X) if mode == TRI_LOWER else (np.triu(X) if mode == TRI_UPPER else X)


def close(got, want):
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-13 * max(1.0, np.abs(want).max()))


@pytest.mark.parametrize('m,n,k,ld_extra', [(140, 150, 38, 0), (129, 127, 33, 1), (5, 3, 2, 0), (130, 260, 16, 2)])
def test_general_product_with_beta_and_both_staging_widths(emu, m, n, k, ld_extra):
    rng = np.random.default_rng(m + n + k)
    A = rng.standard_normal((m, k + ld_extra))[:, :k]          # ld_extra = 1: odd leading dimension -> 8-byte path
    B = rng.standard_normal((n, k + ld_extra))[:, :k]
    C0 = rng.standard_normal((m, n))
    got = gemm(emu, A, B, C0.copy(), alpha=0.7, beta=-0.3)
    close(got, 0.7 * A @ B.T - 0.3 * C0)
    got = gemm(emu, A, B, np.full((m, n), np.nan), alpha=-1.0, beta=0.0)      # beta = 0 never reads C
    close(got, -A @ B.T)


@pytest.mark.parametrize('a_tri', [TRI_NONE, TRI_LOWER, TRI_UPPER])
@pytest.mark.parametrize('b_tri', [TRI_NONE, TRI_LOWER, TRI_UPPER])
@pytest.mark.parametrize('c_uplo', [C_ALL, C_LOWER])
def test_every_triangular_combination(emu, a_tri, b_tri, c_uplo):
    """Square n = 160 (two tile rows, ragged last tile; the three-tile-row case is the test
    below).  Operands are TRULY triangular (the kernel skips whole zero tiles but reads diagonal
    tiles in full); with c_uplo = lower only the lower triangle of C may be written."""
    n = 160
    rng = np.random.default_rng(a_tri * 9 + b_tri * 3 + c_uplo)
    A, B = tri(rng.standard_normal((n, n)), a_tri), tri(rng.standard_normal((n, n)), b_tri)
    C0 = rng.standard_normal((n, n))
    got = gemm(emu, A, B, C0.copy(), alpha=1.3, beta=0.5, a_tri=a_tri, b_tri=b_tri, c_uplo=c_uplo)
    full = 1.3 * A @ B.T + 0.5 * C0
    if c_uplo == C_LOWER:
        close(np.tril(got), np.tril(full))
        iu = np.triu_indices(n, 1)
        assert (got[iu] == C0[iu]).all(), 'strict upper triangle of C must stay untouched'
    else:
        close(got, full)


def test_three_tile_rows_triangular_inverse_shape(emu):
    """U U^T with both operands upper triangular and lower output (K^-1 = U U^T, potrf.cu) on three
    tile rows: interior tiles, diagonal tiles and skipped tiles all occur."""
    n = 300
    U = np.triu(np.random.default_rng(2).standard_normal((n, n)))
    got = gemm(emu, U, U, np.zeros((n, n)), a_tri=TRI_UPPER, b_tri=TRI_UPPER, c_uplo=C_LOWER)
    close(np.tril(got), np.tril(U @ U.T))
    assert np.abs(np.triu(got, 1)).max() == 0.0


def test_rectangular_operands_with_lower_output(emu):
    """The shape of the TRSM / triangular-aware matmul adjoints: C (n x n, lower) = A^T-like
    [n x K] times [n x K]^T with K = data rows, b_tri flipped."""
    n, K = 200, 40
    rng = np.random.default_rng(4)
    A, B = rng.standard_normal((n, K)), rng.standard_normal((n, K))
    got = gemm(emu, A, B, np.zeros((n, n)), alpha=-1.0, c_uplo=C_LOWER)
    close(np.tril(got), np.tril(-A @ B.T))
    assert np.abs(np.triu(got, 1)).max() == 0.0
    Bl = np.tril(rng.standard_normal((n, n)))
    G = rng.standard_normal((140, n))
    close(gemm(emu, G, Bl, np.empty((140, n)), b_tri=TRI_LOWER), G @ Bl.T)
    Bu = np.triu(rng.standard_normal((n, n)))
    close(gemm(emu, G, Bu, np.empty((140, n)), b_tri=TRI_UPPER), G @ Bu.T)


def test_rowmap_mask_of_the_block_row_distribution(emu):
    """C_ROWMAP: element (r, c) is updated iff c + coff <= rowlim[r] (rowlim non-decreasing)."""
    m, n, k, coff = 260, 200, 24, 64
    rng = np.random.default_rng(8)
    A, B, C0 = rng.standard_normal((m, k)), rng.standard_normal((n, k)), rng.standard_normal((m, n))
    rowlim = np.sort(rng.integers(0, n + coff + 40, m)).astype(np.int64)
    got = gemm(emu, A, B, C0.copy(), alpha=-1.0, beta=1.0, c_uplo=C_ROWMAP, rowlim=rowlim, coff=coff)
    mask = (np.arange(n)[None, :] + coff) <= rowlim[:, None]
    close(got, np.where(mask, C0 - A @ B.T, C0))


def test_per_row_start_modes_of_the_prefix_solves(emu):
    """lo_mode 1: A[r][k] == 0 for lo_off + k < rowlo[r] (K range of a tile starts later);
    lo_mode 2: C[r][c] is not wanted for lo_off + c < rowlo[r] (tiles wholly left are skipped)."""
    m, n, k = 384, 140, 384
    rng = np.random.default_rng(12)
    rowlo = (np.sort(rng.integers(0, 3, m)) * 128).astype(np.int64)          # multiples of 128, sorted
    A = rng.standard_normal((m, k))
    A[np.arange(k)[None, :] < rowlo[:, None]] = 0.0
    B = rng.standard_normal((n, k))
    close(gemm(emu, A, B, np.empty((m, n)), rowlo=rowlo, lo_off=0, lo_mode=1), A @ B.T)
    A2 = rng.standard_normal((m, k))
    C0 = rng.standard_normal((m, n))
    got = gemm(emu, A2, B, C0.copy(), alpha=1.0, beta=0.0, rowlo=rowlo, lo_off=0, lo_mode=2)
    want = A2 @ B.T
    wanted = np.arange(n)[None, :] >= rowlo[:, None]
    close(got[wanted], want[wanted])          # entries left of a row's start may or may not be written
