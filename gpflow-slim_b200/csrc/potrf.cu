// Blocked recursive FP64 Cholesky, triangular solves and triangular inverse.
//
// Replaces tf.cholesky / tf.matrix_triangular_solve of the reference (call sites listed in
// include/gpslim_b200.h).  Row-major, lower.  Structure:
//   * leaves: one CTA factors a 128x128 diagonal block in shared memory AND inverts it
//     (T = L_kk^-1); the rows beneath the block are then solved in place against T^T by
//     64-row strips on the FP64 tensor cores (trsm_strip_kernel);
//   * inner nodes: the trailing update  A22 -= A21 A21^T  (including all rows beneath, so the
//     recursive TRSM updates are the same launch) is one lower-masked DMMA GEMM whose K is the
//     size of the left half -- half of all flops run with K >= N/4.
// The same leaves/inner-node split gives B L^-T (gps_trsm_rec) and U = L^-T
// (gps_inv_upper_rec); K^-1 = U U^T is a single triangular-aware GEMM.
#include "internal.cuh"

namespace {

constexpr int NB = GPS_NB;        // 128
constexpr int SLD = 130;          // smem row stride of the base kernel (see bank analysis below)
constexpr int BASE_THREADS = 256;
constexpr int BASE_SMEM = (NB * SLD + 3 * NB) * (int)sizeof(double);

// ------------------------------------------------------------------ 128x128 leaf
// Thread pair (2i, 2i+1) owns row i; each half-sums the odd / even k terms of the dot
// products.  SLD = 130: for a half-warp (8 rows x 2 parities) the 8-byte words
// (i*130 + k + h) mod 16 are all distinct -> conflict-free.
template <bool DO_CHOL>
__global__ void __launch_bounds__(BASE_THREADS, 1)
potrf_base_kernel(double* __restrict__ Abase, int64_t lda, int n_total, double* __restrict__ tinv,
                  double* __restrict__ logdet, int* __restrict__ info, int row0,
                  double* __restrict__ Ubase, int64_t ldu) {
  extern __shared__ __align__(16) double sm[];
  double* S = sm;                 // [NB][SLD]; lower: L, strict upper (transposed): T
  double* rinv = sm + NB * SLD;   // 1 / L_ii
  double* piv = rinv + NB;        // raw pivots
  double* red = piv + NB;         // scratch for the log-det reduction
  const int blk = blockIdx.x;
  const int tid = threadIdx.x;
  const int n = min(NB, n_total - blk * NB);
  double* A = Abase + (int64_t)blk * NB * (lda + 1);

  for (int idx = tid; idx < NB * NB; idx += BASE_THREADS) {
    int i = idx >> 7, j = idx & (NB - 1);
    double v = (i == j) ? 1.0 : 0.0;          // identity padding for a partial last block
    if (i < n && j <= i) v = A[(int64_t)i * lda + j];
    S[i * SLD + j] = v;
  }
  __syncthreads();

  const int i = tid >> 1, hh = tid & 1;
  if (DO_CHOL) {
    int bad = 0;
    for (int j = 0; j < NB; ++j) {
      double p0 = 0.0, p1 = 0.0;
      if (i >= j) {
        const double* ri = S + i * SLD;
        const double* rj = S + j * SLD;
        int k = hh;
        for (; k + 2 < j; k += 4) {
          p0 = fma(ri[k], rj[k], p0);
          p1 = fma(ri[k + 2], rj[k + 2], p1);
        }
        if (k < j) p0 = fma(ri[k], rj[k], p0);
      }
      double part = p0 + p1;
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      double s = 0.0;
      if (i >= j) s = S[i * SLD + j] - part;
      if (i == j && hh == 0) piv[j] = s;             // raw pivot
      __syncthreads();
      double d = piv[j];
      if (!(d > 0.0)) {                              // also catches NaN
        if (!bad && tid == 0 && j < n) {
          int val = row0 + blk * NB + j + 1;
          int old = atomicCAS(info, 0, val);
          if (old != 0 && old > val) atomicMin(info, val);
        }
        bad = 1;
      }
      double r = sqrt(d);
      if (hh == 0) {
        if (i == j) {
          S[j * SLD + j] = r;
          rinv[j] = 1.0 / r;
        } else if (i > j) {
          S[i * SLD + j] = s / r;
        }
      }
      __syncthreads();
    }
  } else {
    if (tid < NB) rinv[tid] = 1.0 / S[tid * SLD + tid];
    __syncthreads();
  }

  // T = L^-1, column c by thread pair c (forward substitution down the rows); T[i][c] (i > c)
  // is kept at S[c][i], i.e. in the unused strict upper triangle.
  {
    const int c = i;
    const int tmax = NB - 1 - (tid >> 5) * 16;   // warp-uniform trip count (16 columns / warp)
    const double* tc = S + c * SLD;
    for (int t = 0; t < tmax; ++t) {
      const int r = c + 1 + t;
      const bool act = r < NB;
      const double* lr = S + (act ? r : c) * SLD;
      double p0 = 0.0, p1 = 0.0;
      // the k = c term uses T[c][c] = rinv[c] and is added below
      int k = c + 1 + hh;
      const int kend = act ? r : 0;
      for (; k + 2 < kend; k += 4) {
        p0 = fma(lr[k], tc[k], p0);
        p1 = fma(lr[k + 2], tc[k + 2], p1);
      }
      if (k < kend) p0 = fma(lr[k], tc[k], p0);
      double part = p0 + p1;
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      if (act && hh == 0) S[c * SLD + r] = -(part + lr[c] * rinv[c]) * rinv[r];
      __syncwarp();
    }
  }
  __syncthreads();

  // write back: L (lower part of A), T (dense, zero upper), U diag tile = T^T (zero lower)
  if (DO_CHOL) {
    for (int idx = tid; idx < NB * NB; idx += BASE_THREADS) {
      int r = idx >> 7, cc = idx & (NB - 1);
      if (r < n && cc <= r) A[(int64_t)r * lda + cc] = S[r * SLD + cc];
    }
  }
  if (tinv) {
    double* T = tinv + (int64_t)blk * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += BASE_THREADS) {
      int r = idx >> 7, cc = idx & (NB - 1);
      double v = 0.0;
      if (cc < r) v = S[cc * SLD + r];
      else if (cc == r) v = rinv[r];
      T[idx] = v;
    }
  }
  if (Ubase) {
    double* U = Ubase + (int64_t)blk * NB * (ldu + 1);
    for (int idx = tid; idx < NB * NB; idx += BASE_THREADS) {
      int r = idx >> 7, cc = idx & (NB - 1);
      if (r < n && cc < n) {
        double v = 0.0;
        if (cc > r) v = S[r * SLD + cc];
        else if (cc == r) v = rinv[r];
        U[(int64_t)r * ldu + cc] = v;
      }
    }
  }
  if (logdet) {
    double s = 0.0;
    if (tid < n) s = -log(rinv[tid]);
    s = warp_sum(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < BASE_THREADS / 32; ++w) t += red[w];
      logdet[blk] = t;
    }
  }
}

// ------------------------------------------------------------------ 128x128 leaf, blocked (DMMA)
// The production leaf.  Right-looking Cholesky over 8-column panels entirely in shared memory:
//   * the 8x8 diagonal block is factored AND inverted by warp 0 in registers (row per lane,
//     warp shuffles), one panel ahead of everybody else (look-ahead inside the CTA);
//   * panel solve  X = B Td^T  and trailing update  C -= X X^T  are m8n8k4 DMMAs on 8x8 tiles,
//     spread over the 8 warps;
// then T = L^-1 by block forward substitution (one block column per warp, DMMA products,
// T_ij kept transposed in the unused strict upper triangle of the same array).
constexpr int LLD = 132;   // 132 = 4 mod 16: conflict-free m8n8k4 fragment loads
constexpr int LEAF_SMEM = (NB * LLD + 16 * 64 + 8 * 64 + 32) * (int)sizeof(double);

// inverse of the 8x8 lower block at S[c0.., c0..] -> Tdp[i*8 + c]; lane c = lane & 7 owns column c
__device__ __forceinline__ void inv8(const double* S, double* Tdp, int c0, int lane, double myrinv) {
  const int c = lane & 7;
  double t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double sacc = 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) sacc = fma(S[(c0 + i) * LLD + c0 + k], t[k], sacc);
    double ri = __shfl_sync(0xffffffffu, myrinv, i);
    t[i] = (i == c) ? ri : ((i > c) ? -sacc * ri : 0.0);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) Tdp[i * 8 + c] = t[i];
  }
}

// 1 / sqrt(a) to FP64 rounding accuracy from the single-precision MUFU seed and two Newton steps
// (relative error 2^-22 -> 1e-13 -> 1e-26): the library rsqrt(double) costs ~3x more in the
// serial chain of the 8x8 factorisation.  Out-of-range arguments take the library routine.
__device__ __forceinline__ double fast_rsqrt(double a) {
  if (!(a > 1e-30 && a < 1e30)) return rsqrt(a);
  double y = (double)rsqrtf((float)a);
  const double h = 0.5 * a;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}

// Cholesky + inverse of the 8x8 diagonal block (executed by one full warp).  EVERY lane holds the
// whole lower triangle in registers and factors it redundantly: no shuffle sits on the serial
// chain (sqrt -> scale -> rank-1 update, 8 times), which is the critical path of the whole leaf --
// the row-per-lane variant spent ~4500 cycles here per 8 columns, mostly in shuffles and the
// library rsqrt.  Lane c < 8 then inverts column c of the factor from the same registers.
__device__ __forceinline__ void factor8(double* S, double* Tdp, int c0, int lane, int* s_bad) {
  double a[8][8], rinv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) a[i][j] = S[(c0 + i) * LLD + c0 + j];      // broadcast loads
  __syncwarp();      // every lane has read the block before lanes 0..7 overwrite it below
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double akk = a[k][k];
    if (!(akk > 0.0) && lane == 0 && *s_bad == 0) *s_bad = c0 + k + 1;
    const double r = fast_rsqrt(akk);
    rinv[k] = r;
    a[k][k] = akk * r;
#pragma unroll
    for (int i = k + 1; i < 8; ++i) a[i][k] *= r;
#pragma unroll
    for (int j = k + 1; j < 8; ++j)
#pragma unroll
      for (int i = j; i < 8; ++i) a[i][j] = fma(-a[i][k], a[j][k], a[i][j]);
  }
  const int l = lane & 7;
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {           // lane l stores row l (static register indices only)
      if (i == l) {
#pragma unroll
        for (int c = 0; c <= i; ++c) S[(c0 + i) * LLD + c0 + c] = a[i][c];
      }
    }
  }
  // T = L^-1, column c = l by forward substitution: t[i] = -rinv[i] * sum_{c <= k < i} a[i][k] t[k]
  double t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double sacc = 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) sacc = fma(a[i][k], t[k], sacc);           // t[k] == 0 for k < c
    t[i] = (i == l) ? rinv[i] : ((i > l) ? -sacc * rinv[i] : 0.0);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) Tdp[i * 8 + l] = t[i];
  }
}

template <bool DO_CHOL>
__global__ void __launch_bounds__(256, 1)
potrf_leaf_kernel(double* __restrict__ Abase, int64_t lda, int n_total, double* __restrict__ tinv,
                  double* __restrict__ logdet, int* __restrict__ info, int row0,
                  double* __restrict__ Ubase, int64_t ldu) {
  extern __shared__ __align__(16) double sm[];
  double* S = sm;                    // [NB][LLD]
  double* Td = S + NB * LLD;         // [16][8][8] inverses of the diagonal 8x8 blocks
  double* scr = Td + 16 * 64;        // [8 warps][8][8]
  double* red = scr + 8 * 64;        // [32]
  __shared__ int s_bad;
  const int blk = blockIdx.x;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int n = min(NB, n_total - blk * NB);
  double* A = Abase + (int64_t)blk * NB * (lda + 1);

  for (int idx = tid; idx < NB * NB; idx += 256) {
    int i = idx >> 7, j = idx & (NB - 1);
    double v = (i == j) ? 1.0 : 0.0;
    if (i < n && j <= i) v = A[(int64_t)i * lda + j];
    S[i * LLD + j] = v;
  }
  if (tid == 0) s_bad = 0;
  // (row, column) of the t-th tile of a lower triangle enumerated by rows: decoded once instead of
  // with a square root per tile in the trailing-update loop
  __shared__ unsigned char tile_r[120], tile_c[120];
  if (tid < 120) {
    int r = (int)((sqrtf(8.0f * (float)tid + 1.0f) - 1.0f) * 0.5f);
    while (r * (r + 1) / 2 > tid) --r;
    while ((r + 1) * (r + 2) / 2 <= tid) ++r;
    tile_r[tid] = (unsigned char)r;
    tile_c[tid] = (unsigned char)(tid - r * (r + 1) / 2);
  }
  __syncthreads();

  if (DO_CHOL) {
    if (warp == 0) factor8(S, Td, 0, lane, &s_bad);
    __syncthreads();
    for (int p = 0; p < 16; ++p) {
      const int c0 = 8 * p;
      // (a) panel solve: tile rows bi > p
      for (int bi = p + 1 + warp; bi < 16; bi += 8) {
        const double* ap = S + (8 * bi + lr) * LLD + c0 + lc;
        const double* tp = Td + p * 64 + lr * 8 + lc;
        double x0 = 0.0, x1 = 0.0;
        double a0 = ap[0], a1 = ap[4], b0 = tp[0], b1 = tp[4];
        dmma884(x0, x1, a0, b0);
        dmma884(x0, x1, a1, b1);
        *reinterpret_cast<double2*>(S + (8 * bi + lr) * LLD + c0 + 2 * lc) = make_double2(x0, x1);
      }
      __syncthreads();
      if (p == 15) break;
      // (b) trailing update with panel p; warp 0 runs one panel ahead on the diagonal
      const int m = 15 - p;
      const int ntiles = m * (m + 1) / 2;
      const int tstart = (warp == 0) ? 0 : warp;
      const int tstep = (warp == 0) ? ntiles : 7;     // warp 0 takes tile 0 only
      for (int t = tstart; t < ntiles; t += tstep) {
        const int r = tile_r[t], c = tile_c[t];
        const int bi = p + 1 + r, bj = p + 1 + c;
        const double* ap = S + (8 * bi + lr) * LLD + c0 + lc;
        const double* bp = S + (8 * bj + lr) * LLD + c0 + lc;
        double2* cp = reinterpret_cast<double2*>(S + (8 * bi + lr) * LLD + 8 * bj + 2 * lc);
        double a0 = -ap[0], a1 = -ap[4], b0 = bp[0], b1 = bp[4];
        double2 cv = *cp;
        dmma884(cv.x, cv.y, a0, b0);
        dmma884(cv.x, cv.y, a1, b1);
        *cp = cv;
      }
      if (warp == 0) {
        __syncwarp();
        factor8(S, Td + (p + 1) * 64, 8 * (p + 1), lane, &s_bad);
      }
      __syncthreads();
    }
    if (tid == 0 && s_bad != 0 && s_bad <= n) {
      int val = row0 + blk * NB + s_bad;
      int old = atomicCAS(info, 0, val);
      if (old != 0 && old > val) atomicMin(info, val);
    }
  } else {
    for (int b = 2 * warp; b < 2 * warp + 2; ++b) {
      const int l = lane & 7;
      double myrinv = 1.0 / S[(8 * b + l) * LLD + 8 * b + l];
      inv8(S, Td + b * 64, 8 * b, lane, myrinv);
    }
    __syncthreads();
  }

  // T = L^-1: block column j by one warp; T_ij (i > j) stored transposed at S[8j.., 8i..]
  for (int jj = 0; jj < 2; ++jj) {
    const int j = (jj == 0) ? warp : 15 - warp;
    double* sc = scr + warp * 64;
    for (int i = j + 1; i < 16; ++i) {
      double c0a = 0.0, c1a = 0.0, c0b = 0.0, c1b = 0.0;
      {
        const double* ap = S + (8 * i + lr) * LLD + 8 * j + lc;
        const double* bp = Td + j * 64 + lc * 8 + lr;
        dmma884(c0a, c1a, ap[0], bp[0]);
        dmma884(c0b, c1b, ap[4], bp[32]);
      }
      for (int k = j + 1; k < i; ++k) {
        const double* ap = S + (8 * i + lr) * LLD + 8 * k + lc;
        const double* bp = S + (8 * j + lr) * LLD + 8 * k + lc;
        dmma884(c0a, c1a, ap[0], bp[0]);
        dmma884(c0b, c1b, ap[4], bp[4]);
      }
      sc[lr * 8 + 2 * lc] = c0a + c0b;
      sc[lr * 8 + 2 * lc + 1] = c1a + c1b;
      __syncwarp();
      double t0 = 0.0, t1 = 0.0;
      const double* tp = Td + i * 64 + lr * 8 + lc;
      dmma884(t0, t1, tp[0], sc[lc * 8 + lr]);
      dmma884(t0, t1, tp[4], sc[(lc + 4) * 8 + lr]);
      __syncwarp();
      S[(8 * j + 2 * lc) * LLD + 8 * i + lr] = -t0;
      S[(8 * j + 2 * lc + 1) * LLD + 8 * i + lr] = -t1;
      __syncwarp();
    }
  }
  __syncthreads();

  if (DO_CHOL) {
    for (int idx = tid; idx < NB * NB; idx += 256) {
      int r = idx >> 7, cc = idx & (NB - 1);
      if (r < n && cc <= r) A[(int64_t)r * lda + cc] = S[r * LLD + cc];
    }
  }
  if (tinv) {
    double* T = tinv + (int64_t)blk * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += 256) {
      int r = idx >> 7, cc = idx & (NB - 1);
      double v = 0.0;
      if ((r >> 3) == (cc >> 3)) v = Td[(r >> 3) * 64 + (r & 7) * 8 + (cc & 7)];
      else if (cc < r) v = S[cc * LLD + r];
      T[idx] = v;
    }
  }
  if (Ubase) {
    double* U = Ubase + (int64_t)blk * NB * (ldu + 1);
    for (int idx = tid; idx < NB * NB; idx += 256) {
      int r = idx >> 7, cc = idx & (NB - 1);
      if (r < n && cc < n) {
        double v = 0.0;   // U[r][cc] = T[cc][r]
        if ((r >> 3) == (cc >> 3)) v = Td[(r >> 3) * 64 + (cc & 7) * 8 + (r & 7)];
        else if (cc > r) v = S[r * LLD + cc];
        U[(int64_t)r * ldu + cc] = v;
      }
    }
  }
  if (logdet) {
    double sdet = 0.0;
    if (tid < n) sdet = log(S[tid * LLD + tid]);
    sdet = warp_sum(sdet);
    if (lane == 0) red[warp] = sdet;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      logdet[blk] = t;
    }
  }
}

// ------------------------------------------------------------------ strip TRSM on DMMA
// B[r0:r0+64, 0:n] <- B[r0:r0+64, 0:n] * T^T   (T = 128x128 dense inverse of the diagonal
// block, lower triangular with explicit zeros above the diagonal).  In place: a CTA reads
// only the strip it overwrites.
constexpr int TS_LD = 132;   // 132 = 4 mod 16 -> conflict-free DMMA fragment loads
constexpr int STRIP = 64;
constexpr int STRIP_SMEM = (NB + STRIP) * TS_LD * (int)sizeof(double);

// NOTRANS = false:  B <- B T^T   (forward solve  B L^-T, T = L_kk^-1 lower)
// NOTRANS = true :  B <- B T     (solve against the untransposed factor, B L^-1): the tile is
//                   staged transposed, i.e. the kernel multiplies by (T^T)^T with T^T upper.
template <bool NOTRANS>
__global__ void __launch_bounds__(256, 1)
trsm_strip_kernel(double* __restrict__ B, int64_t ldb, int m, int n, const double* __restrict__ T) {
  extern __shared__ __align__(16) double sm[];
  double* Ts = sm;                  // [NB][TS_LD]
  double* Bs = sm + NB * TS_LD;     // [STRIP][TS_LD]
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * STRIP;
  // all copies in flight at once (cp.async): T in 16-byte chunks, the strip in 8-byte ones
  // (a user-supplied B may have an odd leading dimension)
  if (NOTRANS) {
    for (int idx = tid; idx < NB * NB; idx += 256) {
      int r = idx >> 7, c = idx & (NB - 1);
      Ts[c * TS_LD + r] = T[idx];   // Ts = T^T (upper); T is L2-resident (128 KiB)
    }
  } else {
    for (int idx = tid; idx < NB * NB / 2; idx += 256) {
      int r = idx >> 6, c2 = (idx & 63) * 2;
      cp_async16(Ts + r * TS_LD + c2, T + r * NB + c2, 16);
    }
  }
  for (int idx = tid; idx < STRIP * NB; idx += 256) {
    int r = idx >> 7, c = idx & (NB - 1);
    bool ok = (r0 + r < m) && (c < n);
    cp_async8(Bs + r * TS_LD + c, ok ? B + (int64_t)(r0 + r) * ldb + c : B, ok ? 8 : 0);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;
  const int mrow = (warp & 3) * 16;     // 2 m-subtiles of 8 rows
  const int nbase = (warp >> 2) * 64;   // 8 n-subtiles of 8 columns
  double acc[2][8][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[a][j][0] = acc[a][j][1] = 0.0;
  // lower tile: Ts[c][k] == 0 for k > c; upper tile (NOTRANS): Ts[c][k] == 0 for k < c
  const int kbeg = NOTRANS ? nbase : 0;
  const int kend = NOTRANS ? n : min(n, nbase + 64);
  for (int k0 = kbeg; k0 < kend; k0 += 4) {
    double a0 = Bs[(mrow + lr) * TS_LD + k0 + lc];
    double a1 = Bs[(mrow + 8 + lr) * TS_LD + k0 + lc];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (NOTRANS ? (k0 + 3 >= nbase + j * 8) : (k0 <= nbase + j * 8 + 7)) {     // warp-uniform
        double b = Ts[(nbase + j * 8 + lr) * TS_LD + k0 + lc];
        dmma884(acc[0][j][0], acc[0][j][1], a0, b);
        dmma884(acc[1][j][0], acc[1][j][1], a1, b);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    int row = r0 + mrow + a * 8 + lr;
    if (row >= m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int col = nbase + j * 8 + lc * 2;
      double* bp = B + (int64_t)row * ldb + col;
      if (col < n) bp[0] = acc[a][j][0];
      if (col + 1 < n) bp[1] = acc[a][j][1];
    }
  }
}

void set_attrs(gps_handle* h) {
  if (h->attr_potrf) return;
  cudaFuncSetAttribute(potrf_base_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BASE_SMEM);
  cudaFuncSetAttribute(potrf_base_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BASE_SMEM);
  cudaFuncSetAttribute(trsm_strip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, STRIP_SMEM);
  cudaFuncSetAttribute(trsm_strip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, STRIP_SMEM);
  cudaFuncSetAttribute(potrf_leaf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM);
  cudaFuncSetAttribute(potrf_leaf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM);
  h->attr_potrf = true;
}

int64_t split_point(int64_t n) {
  int64_t nb = (n + NB - 1) / NB;
  return (nb / 2) * NB;
}

int strip_launch(gps_handle* h, Mat B, int64_t n, const double* T, bool notrans = false) {
  if (B.rows <= 0) return 0;
  set_attrs(h);
  unsigned grid = (unsigned)((B.rows + STRIP - 1) / STRIP);
  if (notrans)
    trsm_strip_kernel<true><<<grid, 256, STRIP_SMEM, h->stream>>>(B.p, B.ld, (int)B.rows, (int)n, T);
  else
    trsm_strip_kernel<false><<<grid, 256, STRIP_SMEM, h->stream>>>(B.p, B.ld, (int)B.rows, (int)n, T);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

// ---- row structure of the prefix solves -------------------------------------------------
// The rows of B are sorted by a per-row first column start[r] (a multiple of 128): left of it
// the row is identically zero (forward solve) / not wanted (solve against L itself).  Held as
// segments of equal start; a device copy of the per-row array feeds the GEMM masks.
struct RowSegs {
  std::vector<int64_t> begin, count, start;
  int64_t rows = 0;
  const int64_t* dev = nullptr;   // [rows]
  // rows with start < colend (a prefix, since start is non-decreasing)
  int64_t active(int64_t colend) const {
    int64_t n = 0;
    for (size_t i = 0; i < start.size() && start[i] < colend; ++i) n += count[i];
    return n;
  }
  // sum over rows r < nrows of (hi - max(lo, start[r]))  clipped at 0
  double span(int64_t nrows, int64_t lo, int64_t hi) const {
    double t = 0;
    for (size_t i = 0; i < start.size() && begin[i] < nrows; ++i) {
      int64_t c = count[i] < nrows - begin[i] ? count[i] : nrows - begin[i];
      int64_t l = start[i] > lo ? start[i] : lo;
      if (hi > l) t += (double)c * (double)(hi - l);
    }
    return t;
  }
};
}  // namespace

int gps_potrf_rec(gps_handle* h, Mat A, int64_t n, int64_t below, int64_t blk0, double* tinv,
                  double* logdet, int* info_dev, int64_t row0) {
  int rc;
  if (n <= 0) return 0;
  if (n <= NB) {
    set_attrs(h);
    if (h->leaf_impl == 1)
      potrf_base_kernel<true><<<1, BASE_THREADS, BASE_SMEM, h->stream>>>(
          A.p, A.ld, (int)n, tinv + blk0 * NB * NB, logdet ? logdet + blk0 : nullptr, info_dev,
          (int)row0, nullptr, 0);
    else
      potrf_leaf_kernel<true><<<1, 256, LEAF_SMEM, h->stream>>>(
          A.p, A.ld, (int)n, tinv + blk0 * NB * NB, logdet ? logdet + blk0 : nullptr, info_dev,
          (int)row0, nullptr, 0);
    GPS_LAUNCH_CHECK(h);
    if (below > 0) {
      if ((rc = strip_launch(h, A.sub(n, 0, below, n), n, tinv + blk0 * NB * NB))) return rc;
    }
    return 0;
  }
  int64_t n1 = split_point(n), n2 = n - n1;
  if ((rc = gps_potrf_rec(h, A, n1, n2 + below, blk0, tinv, logdet, info_dev, row0))) return rc;
  // trailing update incl. the rows beneath:  C -= P * P[0:n2]^T  (lower-masked)
  Mat P = A.sub(n1, 0, n2 + below, n1);
  Mat Pb = A.sub(n1, 0, n2, n1);
  Mat C = A.sub(n1, n1, n2 + below, n2);
  if ((rc = gps_gemm_nt_launch(h, -1.0, P, Pb, 1.0, C, TRI_NONE, TRI_NONE, C_LOWER))) return rc;
  return gps_potrf_rec(h, A.sub(n1, n1, n2 + below, n2), n2, below, blk0 + n1 / NB, tinv, logdet,
                       info_dev, row0 + n1);
}

int gps_trsm_rec(gps_handle* h, Mat L, Mat B, int64_t blk0, const double* tinv) {
  int rc;
  int64_t n = L.rows;
  if (n <= 0 || B.rows <= 0) return 0;
  if (n <= NB) return strip_launch(h, B, n, tinv + blk0 * NB * NB);
  int64_t n1 = split_point(n), n2 = n - n1;
  Mat B1 = B.sub(0, 0, B.rows, n1), B2 = B.sub(0, n1, B.rows, n2);
  if ((rc = gps_trsm_rec(h, L.sub(0, 0, n1, n1), B1, blk0, tinv))) return rc;
  if ((rc = gps_gemm_nt_launch(h, -1.0, B1, L.sub(n1, 0, n2, n1), 1.0, B2, TRI_NONE, TRI_NONE,
                               C_ALL)))
    return rc;
  return gps_trsm_rec(h, L.sub(n1, n1, n2, n2), B2, blk0 + n1 / NB, tinv);
}

namespace {
// ---- explicit inverses of the aligned diagonal blocks ("big leaves") --------------------
// For every full diagonal block b of size bs (bs = 128 * 2^q) of the factor L:
//   U_b = L_bb^-T (upper)  and  T_b = L_bb^-1 = U_b^T (lower),  stacked densely [b][bs][bs].
// Built bottom-up from the 128 x 128 block inverses, ALL blocks of a level in one strided-batch
// launch per product:   W = -U11 L21^T,   U12 = W U22 = W T22^T,   T21 = U12^T = T22 W^T
// (three NT products per level; keeping T next to U removes every transpose).
__global__ void leaf_diag_kernel(const double* __restrict__ tinv, double* __restrict__ Uo, int64_t ldu,
                                 int64_t zu, double* __restrict__ To, int64_t ldt, int64_t zt, int per) {
  __shared__ double tile[32][33];
  const int blk = blockIdx.z;                      // global 128-block index
  const int z = blk / per, j = blk % per;
  const int64_t offu = (int64_t)z * zu + (int64_t)j * NB * (ldu + 1);
  const int64_t offt = (int64_t)z * zt + (int64_t)j * NB * (ldt + 1);
  const double* T = tinv + (int64_t)blk * NB * NB;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const double v = T[(r0 + i) * NB + c0 + threadIdx.x];
    tile[i][threadIdx.x] = v;
    To[offt + (int64_t)(r0 + i) * ldt + c0 + threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8)      // U[r][c] = T[c][r]
    Uo[offu + (int64_t)(c0 + i) * ldu + r0 + threadIdx.x] = tile[threadIdx.x][i];
}

// U / T: `nfull` blocks of size bs, block z at U + z * zu with leading dimension ldu (same for T);
// W: scratch of nfull * bs * bs / 4 doubles.  The diagonal 128-tiles are written in full (zeros
// on the other side of the diagonal); off-diagonal tiles on the zero side are NOT touched.
int block_inverses_batched(gps_handle* h, Mat L, const double* tinv, int bs, int64_t nfull, double* U,
                           int64_t ldu, int64_t zu, double* T, int64_t ldt, int64_t zt, double* W) {
  int rc;
  const int per = bs / NB;
  leaf_diag_kernel<<<dim3(NB / 32, NB / 32, (unsigned)(nfull * per)), dim3(32, 8), 0, h->stream>>>(tinv, U, ldu, zu, T,
                                                                                                ldt, zt, per);
  GPS_LAUNCH_CHECK(h);
  for (int sz = NB; sz < bs; sz *= 2) {
    const int m = bs / (2 * sz);                       // nodes per block at this level
    GemmBatch b;
    b.ny = m; b.nz = (int)nfull;
    const int64_t uy = (int64_t)2 * sz * (ldu + 1), ty = (int64_t)2 * sz * (ldt + 1);
    const int64_t ly = (int64_t)2 * sz * (L.ld + 1), lz = (int64_t)bs * (L.ld + 1);
    const int64_t wy = (int64_t)sz * sz, wz = (int64_t)m * sz * sz;
    Mat U11(U, sz, sz, ldu), L21(L.p + (int64_t)sz * L.ld, sz, sz, L.ld), Wm(W, sz, sz, sz);
    Mat T22(T + (int64_t)sz * (ldt + 1), sz, sz, ldt), U12(U + sz, sz, sz, ldu), T21(T + (int64_t)sz * ldt, sz, sz, ldt);
    // W = -U11 L21^T
    b.sAy = uy; b.sAz = zu; b.sBy = ly; b.sBz = lz; b.sCy = wy; b.sCz = wz;
    if ((rc = gps_gemm_nt_launch(h, -1.0, U11, L21, 0.0, Wm, TRI_UPPER, TRI_NONE, C_ALL, nullptr, 0, -1.0, nullptr,
                                 0, 0, &b)))
      return rc;
    // U12 = W T22^T
    b.sAy = wy; b.sAz = wz; b.sBy = ty; b.sBz = zt; b.sCy = uy; b.sCz = zu;
    if ((rc = gps_gemm_nt_launch(h, 1.0, Wm, T22, 0.0, U12, TRI_NONE, TRI_LOWER, C_ALL, nullptr, 0, -1.0, nullptr, 0,
                                 0, &b)))
      return rc;
    // T21 = T22 W^T
    b.sAy = ty; b.sAz = zt; b.sBy = wy; b.sBz = wz; b.sCy = ty; b.sCz = zt;
    if ((rc = gps_gemm_nt_launch(h, 1.0, T22, Wm, 0.0, T21, TRI_LOWER, TRI_NONE, C_ALL, nullptr, 0, -1.0, nullptr, 0,
                                 0, &b)))
      return rc;
  }
  return 0;
}

struct LeafInv {
  int bs = 0;               // 0: no big leaves (strips only)
  int64_t nfull = 0;        // number of full aligned blocks
  const double* U = nullptr;
  const double* T = nullptr;
  double* X = nullptr;      // [maxrows][bs] scratch
};

int build_leaf_inverses(gps_handle* h, Mat L, const double* tinv, int64_t maxrows, LeafInv* out) {
  int rc;
  const int bs = h->trsm_leaf;
  out->bs = 0;
  if (bs <= NB || L.rows < bs) return 0;
  const int64_t nfull = L.rows / bs;
  const size_t bytes = (size_t)nfull * bs * bs * sizeof(double);
  double* U = (double*)gps_ws(h, WS_LEAF_U, bytes);
  double* T = (double*)gps_ws(h, WS_LEAF_T, bytes);
  double* W = (double*)gps_ws(h, WS_LEAF_W, bytes / 4);
  double* X = (double*)gps_ws(h, WS_LEAF_X, (size_t)maxrows * bs * sizeof(double));
  if (!U || !T || !W || !X) return -102;
  GPS_CUDA(h, cudaMemsetAsync(U, 0, bytes, h->stream));
  GPS_CUDA(h, cudaMemsetAsync(T, 0, bytes, h->stream));
  if ((rc = block_inverses_batched(h, L, tinv, bs, nfull, U, bs, (int64_t)bs * bs, T, bs, (int64_t)bs * bs, W)))
    return rc;
  out->bs = bs; out->nfull = nfull; out->U = U; out->T = T; out->X = X;
  return 0;
}

// a node of the prefix recursions is split at a multiple of the big-leaf size while it is larger
// than one leaf (so that leaves stay aligned); below that, at 128-block granularity as before
int64_t split_node(int64_t n, int leaf) {
  if (leaf > NB && n > leaf) return ((n + leaf - 1) / leaf / 2) * (int64_t)leaf;
  return split_point(n);
}

// rows x bs leaf:  B <- B * op(inverse)  by one product into scratch + copy back
int leaf_apply(gps_handle* h, Mat B, int64_t rows, const LeafInv& lv, int64_t c0, bool notrans) {
  int rc;
  const int bs = lv.bs;
  const int64_t b = c0 / bs;
  Mat X(lv.X, rows, bs, bs);
  Mat Bl = B.sub(0, 0, rows, bs);
  if (!notrans) {   // B L^-T = B T^T-form: X[r][j] = sum_{k<=j} B[r][k] T[j][k]
    Mat T(const_cast<double*>(lv.T) + b * bs * bs, bs, bs, bs);
    if ((rc = gps_gemm_nt_launch(h, 1.0, Bl, T, 0.0, X, TRI_NONE, TRI_LOWER, C_ALL))) return rc;
  } else {          // B L^-1 = B T: X[r][j] = sum_{k>=j} B[r][k] U[j][k]
    Mat U(const_cast<double*>(lv.U) + b * bs * bs, bs, bs, bs);
    if ((rc = gps_gemm_nt_launch(h, 1.0, Bl, U, 0.0, X, TRI_NONE, TRI_UPPER, C_ALL))) return rc;
  }
  GPS_CUDA(h, cudaMemcpy2DAsync(Bl.p, Bl.ld * sizeof(double), X.p, X.ld * sizeof(double), bs * sizeof(double),
                                rows, cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}

// Forward solve B <- B L^-T where row r of B is zero left of column start[r]: at the node that
// covers columns [c0, c0 + n) only the rows with start < c0 + n take part, and the update GEMM
// skips the leading zero K-range of each row tile.  With B = rows of the identity this yields
// rows of U = L^-T at N^3/3 flops overall.
int trsm_rlt_segs(gps_handle* h, Mat L, Mat B, int64_t c0, const double* tinv, const RowSegs& rs,
                  const LeafInv& lv) {
  int rc;
  int64_t n = L.rows;
  if (n <= 0) return 0;
  const int64_t rows = rs.active(c0 + n);
  if (rows <= 0) return 0;
  if (lv.bs > NB && n == lv.bs && c0 % lv.bs == 0 && c0 / lv.bs < lv.nfull)
    return leaf_apply(h, B, rows, lv, c0, false);
  if (n <= NB) return strip_launch(h, B.sub(0, 0, rows, n), n, tinv + (c0 / NB) * NB * NB);
  int64_t n1 = split_node(n, lv.bs), n2 = n - n1;
  if ((rc = trsm_rlt_segs(h, L.sub(0, 0, n1, n1), B.sub(0, 0, rows, n1), c0, tinv, rs, lv))) return rc;
  const int64_t rows1 = rs.active(c0 + n1);
  if (rows1 > 0) {
    double flops = 2.0 * (double)n2 * rs.span(rows1, c0, c0 + n1);
    if ((rc = gps_gemm_nt_launch(h, -1.0, B.sub(0, 0, rows1, n1), L.sub(n1, 0, n2, n1), 1.0,
                                 B.sub(0, n1, rows1, n2), TRI_NONE, TRI_NONE, C_ALL, nullptr, 0, flops,
                                 rs.dev, c0, 1)))
      return rc;
  }
  return trsm_rlt_segs(h, L.sub(n1, n1, n2, n2), B.sub(0, n1, rows, n2), c0 + n1, tinv, rs, lv);
}

// B <- B L^-1 (solve X L = B) given Lt = L^T (upper, row-major) so that every product is NT;
// columns are resolved right to left; row r wants the columns >= start[r] only.
int trsm_rln_segs(gps_handle* h, Mat Lt, Mat B, int64_t c0, const double* tinv, const RowSegs& rs,
                  const LeafInv& lv) {
  int rc;
  int64_t n = Lt.rows;
  if (n <= 0) return 0;
  const int64_t rows = rs.active(c0 + n);
  if (rows <= 0) return 0;
  if (lv.bs > NB && n == lv.bs && c0 % lv.bs == 0 && c0 / lv.bs < lv.nfull)
    return leaf_apply(h, B, rows, lv, c0, true);
  if (n <= NB) return strip_launch(h, B.sub(0, 0, rows, n), n, tinv + (c0 / NB) * NB * NB, true);
  int64_t n1 = split_node(n, lv.bs), n2 = n - n1;
  if ((rc = trsm_rln_segs(h, Lt.sub(n1, n1, n2, n2), B.sub(0, n1, rows, n2), c0 + n1, tinv, rs, lv)))
    return rc;
  const int64_t rows1 = rs.active(c0 + n1);
  if (rows1 <= 0) return 0;
  // B1 -= X2 L21:  C[r][j] -= sum_k X2[r][k] Lt[j][n1 + k]   (tiles left of start[r] skipped)
  double flops = 2.0 * (double)n2 * rs.span(rows1, c0, c0 + n1);
  if ((rc = gps_gemm_nt_launch(h, -1.0, B.sub(0, n1, rows1, n2), Lt.sub(0, n1, n1, n2), 1.0,
                               B.sub(0, 0, rows1, n1), TRI_NONE, TRI_NONE, C_ALL, nullptr, 0, flops,
                               rs.dev, c0, 2)))
    return rc;
  return trsm_rln_segs(h, Lt.sub(0, 0, n1, n1), B.sub(0, 0, rows1, n1), c0, tinv, rs, lv);
}

int make_segs(gps_handle* h, const int64_t* row_start, int64_t rows, int64_t ncols, RowSegs* out) {
  out->rows = rows;
  if (!row_start) {
    out->begin.push_back(0); out->count.push_back(rows); out->start.push_back(0);
    out->dev = nullptr;
    return 0;
  }
  for (int64_t r = 0; r < rows; ++r) {
    int64_t s = row_start[r];
    if (s < 0 || s % NB || s >= ncols + NB || (r > 0 && s < row_start[r - 1]))
      return gps_fail(h, -4, "row_start must be non-decreasing multiples of %d inside the matrix", NB);
    if (out->start.empty() || out->start.back() != s) {
      out->begin.push_back(r); out->count.push_back(0); out->start.push_back(s);
    }
    out->count.back()++;
  }
  int64_t* dev = (int64_t*)gps_ws(h, WS_ROWLO, (size_t)rows * sizeof(int64_t));
  if (!dev) return -102;
  GPS_CUDA(h, cudaMemcpyAsync(dev, row_start, (size_t)rows * sizeof(int64_t), cudaMemcpyHostToDevice,
                              h->stream));
  out->dev = dev;
  return 0;
}
}  // namespace

int gps_block_inverses(gps_handle* h, Mat L, double* tinv) {
  if (L.rows <= 0) return 0;
  set_attrs(h);
  unsigned nblk = (unsigned)((L.rows + NB - 1) / NB);
  if (h->leaf_impl == 1)
    potrf_base_kernel<false><<<nblk, BASE_THREADS, BASE_SMEM, h->stream>>>(
        L.p, L.ld, (int)L.rows, tinv, nullptr, nullptr, 0, nullptr, 0);
  else
    potrf_leaf_kernel<false><<<nblk, 256, LEAF_SMEM, h->stream>>>(
        L.p, L.ld, (int)L.rows, tinv, nullptr, nullptr, 0, nullptr, 0);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

// diagonal tiles of U from the block inverses (U_kk = T_kk^T)
__global__ void u_diag_kernel(const double* __restrict__ tinv, double* __restrict__ U, int64_t ldu,
                              int n_total) {
  __shared__ double tile[32][33];
  const int blk = blockIdx.z;
  const int n = min(NB, n_total - blk * NB);
  const double* T = tinv + (int64_t)blk * NB * NB;
  double* Ub = U + (int64_t)blk * NB * (ldu + 1);
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;  // tile of T
  for (int i = threadIdx.y; i < 32; i += 8) tile[i][threadIdx.x] = T[(r0 + i) * NB + c0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = c0 + i, c = r0 + threadIdx.x;  // U[r][c] = T[c][r]
    if (r < n && c < n) Ub[(int64_t)r * ldu + c] = tile[threadIdx.x][i];
  }
}

int gps_inv_upper_rec(gps_handle* h, Mat L, Mat U, int64_t blk0, const double* tinv) {
  int rc;
  int64_t n = L.rows;
  if (n <= NB) return 0;  // diagonal tiles are written up front by the caller
  int64_t n1 = split_point(n), n2 = n - n1;
  if ((rc = gps_inv_upper_rec(h, L.sub(0, 0, n1, n1), U.sub(0, 0, n1, n1), blk0, tinv))) return rc;
  if ((rc = gps_inv_upper_rec(h, L.sub(n1, n1, n2, n2), U.sub(n1, n1, n2, n2), blk0 + n1 / NB, tinv)))
    return rc;
  // U12 = -(U11 L21^T) U22.  Both products are triangular-aware NT GEMMs (n^3/3 flops overall):
  //   W   = -U11 L21^T                      (U11 upper)       -> scratch
  //   U12 =  W T22^T with T22 = U22^T       (T22 lower)       -> in place in U
  // T22 is produced by an O(n^2) transpose; its garbage upper tiles are never read.
  Mat U12 = U.sub(0, n1, n1, n2);
  const int64_t ldw = (n2 + 15) / 16 * 16;
  double* wbuf = (double*)gps_ws(h, WS_INVW, (size_t)n1 * ldw * sizeof(double));
  double* tbuf = (double*)gps_ws(h, WS_INVT, (size_t)n2 * ldw * sizeof(double));
  if (!wbuf || !tbuf) return -102;
  Mat W(wbuf, n1, n2, ldw), T22(tbuf, n2, n2, ldw);
  if ((rc = gps_gemm_nt_launch(h, -1.0, U.sub(0, 0, n1, n1), L.sub(n1, 0, n2, n1), 0.0, W,
                               TRI_UPPER, TRI_NONE, C_ALL)))
    return rc;
  if ((rc = gps_transpose_launch(h, U.sub(n1, n1, n2, n2), T22))) return rc;
  return gps_gemm_nt_launch(h, 1.0, W, T22, 0.0, U12, TRI_NONE, TRI_LOWER, C_ALL);
}

static int u_diag_launch(gps_handle* h, const double* tinv, Mat U) {
  unsigned nblk = (unsigned)((U.rows + NB - 1) / NB);
  u_diag_kernel<<<dim3(NB / 32, NB / 32, nblk), dim3(32, 8), 0, h->stream>>>(tinv, U.p, U.ld,
                                                                           (int)U.rows);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

int gps_inv_upper_full(gps_handle* h, Mat L, Mat U, const double* tinv) {
  int rc;
  if (L.rows > NB) {
    // size the per-node scratch for the top node up front (it is the largest)
    int64_t n1 = split_point(L.rows), n2 = L.rows - n1;
    int64_t ldw = (n2 + 15) / 16 * 16;
    if (!gps_ws(h, WS_INVW, (size_t)n1 * ldw * sizeof(double)) ||
        !gps_ws(h, WS_INVT, (size_t)n2 * ldw * sizeof(double)))
      return -102;
  }
  if ((rc = u_diag_launch(h, tinv, U))) return rc;
  return gps_inv_upper_rec(h, L, U, 0, tinv);
}

// ------------------------------------------------------------------------- C ABI
extern "C" {

int gps_potrf(gps_handle* h, DLTensor* A_inout, int zero_upper, int* info_host) {
  if (!h) return -1;
  Mat A;
  int rc;
  if ((rc = gps_as_mat(h, A_inout, 2, "A_inout", &A, false))) return rc;
  if (A.rows != A.cols) return gps_fail(h, -2, "potrf: matrix must be square");
  GPS_CUDA(h, cudaSetDevice(h->device));
  int64_t n = A.rows;
  if (n == 0) {
    if (info_host) *info_host = 0;
    return 0;
  }
  int64_t nblk = (n + NB - 1) / NB;
  double* tinv = (double*)gps_ws(h, WS_TINV, (size_t)nblk * NB * NB * sizeof(double));
  int* info_dev = (int*)gps_ws(h, WS_INFO, 64);
  if (!tinv || !info_dev) return -102;
  GPS_CUDA(h, cudaMemsetAsync(info_dev, 0, sizeof(int), h->stream));
  if ((rc = gps_potrf_rec(h, A, n, 0, 0, tinv, nullptr, info_dev, 0))) return rc;
  if (zero_upper && (rc = gps_zero_upper_launch(h, A))) return rc;
  if (info_host) {
    GPS_CUDA(h, cudaMemcpyAsync(info_host, info_dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    GPS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (*info_host > 0) {
      gps_fail(h, *info_host, "potrf: leading minor of order %d is not positive definite", *info_host);
      return *info_host;
    }
  }
  return 0;
}

int gps_trsm_rlt(gps_handle* h, const DLTensor* Lt, DLTensor* B_inout) {
  if (!h) return -1;
  Mat L, B;
  int rc;
  if ((rc = gps_as_mat(h, Lt, 2, "L", &L, false))) return rc;
  if ((rc = gps_as_mat(h, B_inout, 3, "B_inout", &B, false))) return rc;
  if (L.rows != L.cols) return gps_fail(h, -2, "trsm: L must be square");
  if (B.cols != L.rows) return gps_fail(h, -3, "trsm: B has %lld columns, L is %lld x %lld",
                                        (long long)B.cols, (long long)L.rows, (long long)L.rows);
  GPS_CUDA(h, cudaSetDevice(h->device));
  if (L.rows == 0 || B.rows == 0) return 0;
  int64_t nblk = (L.rows + NB - 1) / NB;
  double* tinv = (double*)gps_ws(h, WS_TINV, (size_t)nblk * NB * NB * sizeof(double));
  if (!tinv) return -102;
  if ((rc = gps_block_inverses(h, L, tinv))) return rc;
  return gps_trsm_rec(h, L, B, 0, tinv);
}

int gps_trsm_rlt_prefix(gps_handle* h, const DLTensor* Lt, DLTensor* B_inout,
                        const int64_t* row_start_host) {
  if (!h) return -1;
  Mat L, B;
  int rc;
  if ((rc = gps_as_mat(h, Lt, 2, "L", &L, false))) return rc;
  if ((rc = gps_as_mat(h, B_inout, 3, "B_inout", &B, false))) return rc;
  if (L.rows != L.cols) return gps_fail(h, -2, "trsm: L must be square");
  if (B.cols != L.rows) return gps_fail(h, -3, "trsm: B has %lld columns, L is %lld x %lld",
                                        (long long)B.cols, (long long)L.rows, (long long)L.rows);
  GPS_CUDA(h, cudaSetDevice(h->device));
  if (L.rows == 0 || B.rows == 0) return 0;
  RowSegs rs;
  if ((rc = make_segs(h, row_start_host, B.rows, L.rows, &rs))) return rc;
  int64_t nblk = (L.rows + NB - 1) / NB;
  double* tinv = (double*)gps_ws(h, WS_TINV, (size_t)nblk * NB * NB * sizeof(double));
  if (!tinv) return -102;
  if ((rc = gps_block_inverses(h, L, tinv))) return rc;
  LeafInv lv;
  if ((rc = build_leaf_inverses(h, L, tinv, B.rows, &lv))) return rc;
  return trsm_rlt_segs(h, L, B, 0, tinv, rs, lv);
}

int gps_trsm_rln_prefix(gps_handle* h, const DLTensor* Lt, const DLTensor* Ltt, DLTensor* B_inout,
                        const int64_t* row_start_host) {
  if (!h) return -1;
  Mat L, LT, B;
  int rc;
  if ((rc = gps_as_mat(h, Lt, 2, "L", &L, false))) return rc;
  if ((rc = gps_as_mat(h, Ltt, 3, "Lt", &LT, false))) return rc;
  if ((rc = gps_as_mat(h, B_inout, 4, "B_inout", &B, false))) return rc;
  if (L.rows != L.cols || LT.rows != L.rows || LT.cols != L.cols)
    return gps_fail(h, -3, "trsm_rln: L and Lt must be square and of equal size");
  if (B.cols != L.rows) return gps_fail(h, -4, "trsm_rln: B has %lld columns, L is %lld x %lld",
                                        (long long)B.cols, (long long)L.rows, (long long)L.rows);
  GPS_CUDA(h, cudaSetDevice(h->device));
  if (L.rows == 0 || B.rows == 0) return 0;
  RowSegs rs;
  if ((rc = make_segs(h, row_start_host, B.rows, L.rows, &rs))) return rc;
  int64_t nblk = (L.rows + NB - 1) / NB;
  double* tinv = (double*)gps_ws(h, WS_TINV, (size_t)nblk * NB * NB * sizeof(double));
  if (!tinv) return -102;
  if ((rc = gps_block_inverses(h, L, tinv))) return rc;
  LeafInv lv;
  if ((rc = build_leaf_inverses(h, L, tinv, B.rows, &lv))) return rc;
  return trsm_rln_segs(h, LT, B, 0, tinv, rs, lv);
}

int gps_tri_inv_t(gps_handle* h, const DLTensor* Lt, DLTensor* U_out) {
  if (!h) return -1;
  Mat L, U;
  int rc;
  if ((rc = gps_as_mat(h, Lt, 2, "L", &L, false))) return rc;
  if ((rc = gps_as_mat(h, U_out, 3, "U_out", &U, false))) return rc;
  if (L.rows != L.cols || U.rows != L.rows || U.cols != L.cols)
    return gps_fail(h, -3, "tri_inv_t: shape mismatch");
  GPS_CUDA(h, cudaSetDevice(h->device));
  if (L.rows == 0) return 0;
  int64_t nblk = (L.rows + NB - 1) / NB;
  double* tinv = (double*)gps_ws(h, WS_TINV, (size_t)nblk * NB * NB * sizeof(double));
  if (!tinv) return -102;
  // strict lower part of U is defined to be zero
  GPS_CUDA(h, cudaMemset2DAsync(U.p, U.ld * sizeof(double), 0, U.cols * sizeof(double), U.rows,
                                h->stream));
  if ((rc = gps_block_inverses(h, L, tinv))) return rc;
  const int64_t n = L.rows;
  if (n > NB && n <= 4096 && n % NB == 0 && ((n / NB) & (n / NB - 1)) == 0) {
    // a power-of-two number of 128-blocks: all nodes of a level in ONE strided-batch launch
    // (3 launches per level instead of 3 per node: 9 instead of 21 at n = 1024)
    double* T = (double*)gps_ws(h, WS_LEAF_T, (size_t)n * n * sizeof(double));
    double* W = (double*)gps_ws(h, WS_LEAF_W, (size_t)n * n / 4 * sizeof(double));
    if (!T || !W) return -102;
    return block_inverses_batched(h, L, tinv, (int)n, 1, U.p, U.ld, 0, T, n, 0, W);
  }
  return gps_inv_upper_full(h, L, U, tinv);
}

}  // extern "C"
