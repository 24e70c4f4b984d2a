// FP64 tensor-core GEMM family:  C = alpha * A * B^T + beta * C,  all row-major, both operands
// with K contiguous ("NT").  This one kernel carries every O(n^3) flop of the hot path: the
// SYRK/GEMM trailing updates and panel updates of the blocked Cholesky (tf.cholesky call
// sites, see gpslim_b200.h), the triangular inverse, K^-1 = U U^T, and the conditional /
// predictive products.
//
// Blackwell note: tcgen05.mma has no f64 kind; FP64 tensor work on sm_100a is the warp-level
// DMMA (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4) with register accumulators, so this is a
// warp-MMA kernel: 128x128 CTA tile, 8 warps of 64x32, BK=16, 4-stage cp.async pipeline into
// padded shared memory (row stride 20 doubles -> conflict-free fragment loads).
//
// Triangular operands: a_tri / b_tri restrict the K range per output tile so that all-zero
// 128-wide operand tiles are never loaded or multiplied (this is what makes TRMM-, TRTRI- and
// LAUUM-shaped products cost their true flop count).  c_uplo = lower computes only tiles that
// intersect the lower triangle and masks stores above the diagonal.
#include <cuda.h>

#include "internal.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4;
constexpr int LDS = BK + 4;                       // padded smem row stride (doubles)
constexpr int STAGE_ELEMS = (BM + BN) * LDS;      // doubles per stage
constexpr int GEMM_SMEM = STAGES * STAGE_ELEMS * (int)sizeof(double);  // 163840 B
constexpr int GEMM_THREADS = 256;

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  int64_t lda, ldb, ldc;
  int M, N, K;
  double alpha, beta;
  int a_tri, b_tri, c_uplo;
  int tiles_m, tiles_n;
  // C_ROWMAP: element (r, c) is updated iff c + coff <= rowlim[r]; rowlim is non-decreasing
  // (block-row distributions: rowlim[r] = global row index of local row r)
  const int64_t* rowlim;
  int64_t coff;
  // rows sorted by a per-row first column rowlo[r] (global index, multiple of 128):
  // lo_mode 1: A[r][k] == 0 for lo_off + k < rowlo[r]  -> the K range of a tile starts later
  // lo_mode 2: C[r][c] is not wanted for lo_off + c < rowlo[r] -> tiles wholly left are skipped
  const int64_t* rowlo;
  int64_t lo_off;
  int lo_mode;
  // split-K (option "gemm_splitk", on by default; both tensor-core kernels): blockIdx.y = slice s works on
  // k in [s * k_chunk, (s + 1) * k_chunk) and writes its partial tile (alpha = 1, beta = 0) to
  // part + s * part_stride; splitk_reduce_kernel then forms alpha * sum_s + beta * C.
  int split_k, k_chunk;
  double* part;
  int64_t ldp, part_stride;
  // strided batch (cp.async kernel only): blockIdx.y / blockIdx.z select one of by x bz independent
  // products whose operands sit at uniform strides (diagonal blocks of a matrix, stacked tiles)
  int batch_y;                 // 0: blockIdx.y is the split-K slice; > 0: blockIdx.y is a batch index
  int64_t sAy, sBy, sCy, sAz, sBz, sCz;
};

__device__ __forceinline__ void tile_coords(const GemmArgs& g, int bid, int& ti, int& tj) {
  // grouped ordering: GROUP tile-rows are walked column by column so that concurrently
  // resident CTAs share A row-panels and B row-panels in L2.
  constexpr int GROUP = 8;
  int per_group = GROUP * g.tiles_n;
  int grp = bid / per_group;
  int first = grp * GROUP;
  int gsize = min(GROUP, g.tiles_m - first);
  int rem = bid - grp * per_group;
  ti = first + rem % gsize;
  tj = rem / gsize;
}

__device__ __forceinline__ void k_range(const GemmArgs& g, int ti, int tj, int& klo, int& khi) {
  klo = 0;
  khi = g.K;
  if (g.a_tri == TRI_LOWER) khi = min(khi, (ti + 1) * BM);
  if (g.a_tri == TRI_UPPER) klo = max(klo, ti * BM);
  if (g.b_tri == TRI_LOWER) khi = min(khi, (tj + 1) * BN);
  if (g.b_tri == TRI_UPPER) klo = max(klo, tj * BN);
}

// VEC16: operands are 16-byte aligned with even leading dimensions -> 16 B cp.async
template <bool VEC16>
__device__ __forceinline__ void load_stage(const GemmArgs& g, double* stage, int m0, int n0, int k0,
                                           int khi, int tid) {
  double* As = stage;
  double* Bs = stage + BM * LDS;
  if (VEC16) {
    // 128 rows x 8 chunks of 2 doubles per operand; 256 threads x 4 chunks
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = tid + i * GEMM_THREADS;
      int row = c >> 3, kc = (c & 7) * 2;
      int k = k0 + kc;
      int kb = (k + 1 < khi) ? 16 : (k < khi ? 8 : 0);
      {
        int r = m0 + row;
        int nb = (r < g.M) ? kb : 0;
        const double* src = nb ? g.A + (int64_t)r * g.lda + k : g.A;
        cp_async16(As + row * LDS + kc, src, nb);
      }
      {
        int r = n0 + row;
        int nb = (r < g.N) ? kb : 0;
        const double* src = nb ? g.B + (int64_t)r * g.ldb + k : g.B;
        cp_async16(Bs + row * LDS + kc, src, nb);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c = tid + i * GEMM_THREADS;
      int row = c >> 4, kc = c & 15;
      int k = k0 + kc;
      int kb = (k < khi) ? 8 : 0;
      {
        int r = m0 + row;
        int nb = (r < g.M) ? kb : 0;
        const double* src = nb ? g.A + (int64_t)r * g.lda + k : g.A;
        cp_async8(As + row * LDS + kc, src, nb);
      }
      {
        int r = n0 + row;
        int nb = (r < g.N) ? kb : 0;
        const double* src = nb ? g.B + (int64_t)r * g.ldb + k : g.B;
        cp_async8(Bs + row * LDS + kc, src, nb);
      }
    }
  }
}

// Ask L2 for the C tile at the start of a read-modify-write tile (beta != 0), so that the
// epilogue's loads find it there: 128 rows x 8 lines of 128 B, 4 lines per thread.
__device__ __forceinline__ void prefetch_c_tile(const GemmArgs& g, int m0, int n0, int tid) {
  if (g.beta == 0.0) return;
  const int row = m0 + (tid >> 1);
  if (row >= g.M) return;
  const double* p = g.C + (int64_t)row * g.ldc + n0 + (tid & 1) * 64;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (n0 + (tid & 1) * 64 + q * 16 < g.N) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p + q * 16));
}

__device__ __forceinline__ void gemm_epilogue(const GemmArgs& g, double (&acc)[8][4][2], int m0, int n0,
                                              int wm, int wn, int lr, int lc) {
  const bool vec_ok = ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
  const bool diag_tile = (g.c_uplo == C_LOWER) && (n0 + BN - 1 > m0);
  const bool full_tile = m0 + BM <= g.M && n0 + BN <= g.N;
  // C_ROWMAP: rowlim is non-decreasing, so a tile whose last column is allowed in its first
  // row is unmasked throughout (almost all tiles of a trailing update are)
  const bool rowmap_masked =
      (g.c_uplo == C_ROWMAP) && !(full_tile && (int64_t)n0 + BN - 1 + g.coff <= g.rowlim[m0]);
  if (rowmap_masked) {
    // block-row distributed lower update: per-row column limit
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int row = m0 + wm * 64 + i * 8 + lr;
      if (row >= g.M) continue;
      const int64_t lim = g.rowlim[row] - g.coff;   // last column this row may touch
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int col = n0 + wn * 32 + j * 8 + lc * 2;
        if (col >= g.N) continue;
        double* cp = g.C + (int64_t)row * g.ldc + col;
        const double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
        if (col <= lim) cp[0] = (g.beta != 0.0) ? fma(g.beta, cp[0], v0) : v0;
        if (col + 1 < g.N && col + 1 <= lim) cp[1] = (g.beta != 0.0) ? fma(g.beta, cp[1], v1) : v1;
      }
    }
    return;
  }
  if (vec_ok && !diag_tile && full_tile) {
    // interior tile: all the old values of half a warp tile are requested before the first
    // store, so the read-modify-write costs two memory round trips instead of eight
    double* cbase = g.C + (int64_t)(m0 + wm * 64 + lr) * g.ldc + n0 + wn * 32 + lc * 2;
    if (g.beta != 0.0) {
#pragma unroll
      for (int ih = 0; ih < 2; ++ih) {
        double2 old[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            old[i][j] = __ldcg(reinterpret_cast<const double2*>(cbase + (int64_t)(ih * 4 + i) * 8 * g.ldc + j * 8));
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            double v0 = fma(g.beta, old[i][j].x, g.alpha * acc[ih * 4 + i][j][0]);
            double v1 = fma(g.beta, old[i][j].y, g.alpha * acc[ih * 4 + i][j][1]);
            *reinterpret_cast<double2*>(cbase + (int64_t)(ih * 4 + i) * 8 * g.ldc + j * 8) = make_double2(v0, v1);
          }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<double2*>(cbase + (int64_t)i * 8 * g.ldc + j * 8) =
              make_double2(g.alpha * acc[i][j][0], g.alpha * acc[i][j][1]);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int row = m0 + wm * 64 + i * 8 + lr;
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + wn * 32 + j * 8 + lc * 2;
      if (col >= g.N) continue;
      bool ok0 = !diag_tile || col <= row;
      bool ok1 = (col + 1 < g.N) && (!diag_tile || col + 1 <= row);
      double* cp = g.C + (int64_t)row * g.ldc + col;
      double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
      if (vec_ok && ok0 && ok1) {
        if (g.beta != 0.0) {
          double2 old = *reinterpret_cast<const double2*>(cp);
          v0 = fma(g.beta, old.x, v0);
          v1 = fma(g.beta, old.y, v1);
        }
        *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
      } else {
        if (ok0) cp[0] = (g.beta != 0.0) ? fma(g.beta, cp[0], v0) : v0;
        if (ok1) cp[1] = (g.beta != 0.0) ? fma(g.beta, cp[1], v1) : v1;
      }
    }
  }
}

template <bool VEC16>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_nt_dmma_kernel(const GemmArgs g_in) {
  extern __shared__ __align__(16) double smem[];
  GemmArgs g = g_in;
  if (g.batch_y > 0) {
    const int64_t by = blockIdx.y, bz = blockIdx.z;
    g.A += by * g.sAy + bz * g.sAz;
    g.B += by * g.sBy + bz * g.sBz;
    g.C += by * g.sCy + bz * g.sCz;
  }
  int ti, tj;
  tile_coords(g, blockIdx.x, ti, tj);
  const int m0 = ti * BM, n0 = tj * BN;
  if (g.c_uplo == C_LOWER && m0 + BM - 1 < n0) return;  // tile entirely above the diagonal
  if (g.c_uplo == C_ROWMAP && (int64_t)n0 + g.coff > g.rowlim[min(m0 + BM, g.M) - 1]) return;
  if (g.lo_mode == 2 && (int64_t)n0 + BN - 1 + g.lo_off < g.rowlo[m0]) return;

  int klo, khi;
  k_range(g, ti, tj, klo, khi);
  if (g.lo_mode == 1) {
    int64_t lo = g.rowlo[m0] - g.lo_off;     // rows are sorted: the tile's first row starts first
    if (lo > klo) klo = (int)(lo < khi ? lo / BK * BK : khi);
  }
  if (g.split_k > 1) {                      // this CTA's slice of K (k_chunk is a multiple of BK)
    klo = max(klo, (int)blockIdx.y * g.k_chunk);
    khi = min(khi, ((int)blockIdx.y + 1) * g.k_chunk);
  }
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps, warp tile 64 x 32
  const int lr = lane >> 2, lc = lane & 3;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (khi > klo) ? (khi - klo + BK - 1) / BK : 0;

  // prologue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage<VEC16>(g, smem + s * STAGE_ELEMS, m0, n0, klo + s * BK, khi, tid);
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nxt = kt + STAGES - 1;
      if (nxt < nk)
        load_stage<VEC16>(g, smem + (nxt % STAGES) * STAGE_ELEMS, m0, n0, klo + nxt * BK, khi, tid);
      cp_async_commit();
    }
    const double* As = smem + (kt % STAGES) * STAGE_ELEMS + (wm * 64 + lr) * LDS + lc;
    const double* Bs = smem + (kt % STAGES) * STAGE_ELEMS + BM * LDS + (wn * 32 + lr) * LDS + lc;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[i * 8 * LDS + kk * 4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[j * 8 * LDS + kk * 4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  if (g.split_k > 1) {                      // partial tile of slice blockIdx.y, plain store
    GemmArgs gs = g;
    gs.C = g.part + (int64_t)blockIdx.y * g.part_stride;
    gs.ldc = g.ldp;
    gs.alpha = 1.0;
    gs.beta = 0.0;
    gemm_epilogue(gs, acc, m0, n0, wm, wn, lr, lc);
    return;
  }
  gemm_epilogue(g, acc, m0, n0, wm, wn, lr, lc);
}

// C = alpha * sum_s part[s] + beta * C over the elements the product writes (all, or j <= i)
__global__ void splitk_reduce_kernel(const double* __restrict__ part, int64_t ldp, int64_t stride, int nsplit,
                                     double* __restrict__ C, int64_t ldc, int M, int N, double alpha, double beta,
                                     int lower, const int64_t* __restrict__ rowlo, int64_t lo_off) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;      // blockDim.x == BN: one tile column per block
  for (int i = blockIdx.y; i < M; i += gridDim.y) {
    if (j >= N || (lower && j > i)) continue;
    // lo_mode 2: tiles wholly left of the first wanted column of their first row were never computed
    if (rowlo && (int64_t)blockIdx.x * BN + BN - 1 + lo_off < rowlo[i / BM * BM]) continue;
    double s = 0.0;
    for (int q = 0; q < nsplit; ++q) s += part[(int64_t)q * stride + (int64_t)i * ldp + j];
    double* c = C + (int64_t)i * ldc + j;
    *c = (beta != 0.0) ? fma(beta, *c, alpha * s) : alpha * s;
  }
}

// ------------------------------------------------------------------ TMA + mbarrier variant
// The production kernel when both operands are 16-byte aligned with even leading dimensions
// (every handle-owned buffer is).  Same tile / warp layout and the same tile-skipping rules as
// above, but the operand tiles are staged by the TMA unit (cp.async.bulk.tensor.2d, one elected
// producer thread) into a 6-deep ring of dense 128 x 16 boxes with the 128-byte hardware
// swizzle, and the eight DMMA warps synchronise with the ring through per-stage full / empty
// mbarriers only: there is no CTA-wide barrier in the main loop, the warps drift freely, and
// no DMMA warp spends issue slots on address arithmetic or cp.async.
//   smem: stage s = [A box 16 KiB][B box 16 KiB] at 1024-byte aligned offsets; element (r, k)
//   of a box lives at  r*128 + (((k >> 1) ^ (r & 7)) << 4) + (k & 1)*8  (SWIZZLE_128B), which
//   makes the m8n8k4 fragment loads (8 rows x 4 consecutive k per warp) bank-conflict free.
constexpr int TBK = 16;                          // doubles per k-tile = one 128-byte swizzle row
constexpr int TSTAGES = 6;
constexpr int TBOX_BYTES = BM * TBK * 8;         // 16384
constexpr int TSTAGE_BYTES = 2 * TBOX_BYTES;     // 32768
constexpr int TMA_THREADS = GEMM_THREADS;        // 8 DMMA warps; lane 0 of warp 0 also drives the TMA
constexpr int TMA_SMEM = TSTAGES * TSTAGE_BYTES + 1024 + 2 * 8 * TSTAGES;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (it > (1u << 24)) __trap();   // a lost arrival must fail loudly, never hang the GPU
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
gemm_nt_tma_kernel(const GemmArgs g, const __grid_constant__ CUtensorMap tmA,
                   const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ __align__(16) double smem[];
  int ti, tj;
  tile_coords(g, blockIdx.x, ti, tj);
  const int m0 = ti * BM, n0 = tj * BN;
  if (g.c_uplo == C_LOWER && m0 + BM - 1 < n0) return;
  if (g.c_uplo == C_ROWMAP && (int64_t)n0 + g.coff > g.rowlim[min(m0 + BM, g.M) - 1]) return;
  if (g.lo_mode == 2 && (int64_t)n0 + BN - 1 + g.lo_off < g.rowlo[m0]) return;

  int klo, khi;
  k_range(g, ti, tj, klo, khi);
  if (g.lo_mode == 1) {
    int64_t lo = g.rowlo[m0] - g.lo_off;
    if (lo > klo) klo = (int)(lo < khi ? lo / TBK * TBK : khi);
  }
  if (g.split_k > 1) {   // this CTA's slice of K; k_chunk is a multiple of TBK, so no box straddles a slice
    klo = max(klo, (int)blockIdx.y * g.k_chunk);
    khi = min(khi, ((int)blockIdx.y + 1) * g.k_chunk);
  }
  const int nk = (khi > klo) ? (khi - klo + TBK - 1) / TBK : 0;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t bar_full = base + TSTAGES * TSTAGE_BYTES;
  const uint32_t bar_empty = bar_full + 8 * TSTAGES;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);                    // the producer's arrive.expect_tx
      mbar_init(bar_empty + 8 * s, GEMM_THREADS / 32);   // one arrival per DMMA warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (g.split_k == 1) prefetch_c_tile(g, m0, n0, tid);

  // ---------------- producer duty: lane 0 of warp 0 feeds the ring.  Load L goes to slot
  // L % TSTAGES; it is issued in the middle of iteration L - (TSTAGES - 1), i.e. one full
  // iteration after the slot's previous contents were consumed, so the wait on the empty
  // barrier is normally already satisfied and never holds up DMMA issue.
  auto issue_load = [&](int L) {
    const int sl = L % TSTAGES;
    const uint32_t use = (uint32_t)(L / TSTAGES);
    mbar_wait(bar_empty + 8 * sl, (use & 1u) ^ 1u);      // first use of a slot: returns at once
    mbar_arrive_expect_tx(bar_full + 8 * sl, TSTAGE_BYTES);
    const int k = klo + L * TBK;
    tma_load_2d(base + sl * TSTAGE_BYTES, &tmA, k, m0, bar_full + 8 * sl);
    tma_load_2d(base + sl * TSTAGE_BYTES + TBOX_BYTES, &tmB, k, n0, bar_full + 8 * sl);
  };
  if (tid == 0) {
    for (int L = 0; L < TSTAGES - 1 && L < nk; ++L) issue_load(L);
  }

  // ---------------- consumers: 2 x 4 DMMA warps, warp tile 64 x 32
  const int wm = warp >> 2, wn = warp & 3;
  const int lr = lane >> 2, lc = lane & 3;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // byte offset of this lane's element inside a box row, for the four k-steps of a k-tile
  // A side: fragment row lr of every 8-row group reads box row pr = {0,2,4,6,1,3,5,7}[lr], so the
  // four rows a half-warp touches per 64-bit load sit in four different 32-byte bank groups of
  // the swizzled box (consecutive rows would pair up on the same banks).  The C rows follow.
  const int pr = ((lr & 3) << 1) | (lr >> 2);
  uint32_t koff[TBK / 4], koffa[TBK / 4];
#pragma unroll
  for (int kk = 0; kk < TBK / 4; ++kk) {
    koff[kk] = (uint32_t)((((2 * kk + (lc >> 1)) ^ lr) << 4) | ((lc & 1) << 3));
    koffa[kk] = (uint32_t)((((2 * kk + (lc >> 1)) ^ pr) << 4) | ((lc & 1) << 3));
  }
  const uint32_t a_row = (uint32_t)(wm * 64 + pr) * 128u;
  const uint32_t b_row = (uint32_t)TBOX_BYTES + (uint32_t)(wn * 32 + lr) * 128u;

  double fa[2][8], fb[2][4];
  int s = 0;
  uint32_t ph = 0;
  if (nk > 0) {
    mbar_wait(bar_full, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) fa[0][i] = lds_f64(base + a_row + i * 1024 + koffa[0]);
#pragma unroll
    for (int j = 0; j < 4; ++j) fb[0][j] = lds_f64(base + b_row + j * 1024 + koff[0]);
  }
  for (int kt = 0; kt < nk; ++kt) {
    const uint32_t sb = base + s * TSTAGE_BYTES;
    int s2 = s + 1;
    uint32_t ph2 = ph;
    if (s2 == TSTAGES) { s2 = 0; ph2 ^= 1u; }
#pragma unroll
    for (int kk = 0; kk < TBK / 4; ++kk) {
      const int cur = kk & 1, nxt = cur ^ 1;
      // fragments of the next k-step are fetched while this one's DMMAs issue
      if (kk + 1 < TBK / 4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) fa[nxt][i] = lds_f64(sb + a_row + i * 1024 + koffa[kk + 1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) fb[nxt][j] = lds_f64(sb + b_row + j * 1024 + koff[kk + 1]);
      } else if (kt + 1 < nk) {
        const uint32_t nb = base + s2 * TSTAGE_BYTES;
        mbar_wait(bar_full + 8 * s2, ph2);
#pragma unroll
        for (int i = 0; i < 8; ++i) fa[nxt][i] = lds_f64(nb + a_row + i * 1024 + koffa[0]);
#pragma unroll
        for (int j = 0; j < 4; ++j) fb[nxt][j] = lds_f64(nb + b_row + j * 1024 + koff[0]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[cur][i], fb[cur][j]);
      if (kk == 1 && tid == 0 && kt + TSTAGES - 1 < nk) issue_load(kt + TSTAGES - 1);
    }
    // every fragment of stage s has been consumed by a DMMA: hand the slot back
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * s);
    s = s2;
    ph = ph2;
  }

  if (g.split_k > 1) {                      // partial tile of slice blockIdx.y, plain store
    GemmArgs gs = g;
    gs.C = g.part + (int64_t)blockIdx.y * g.part_stride;
    gs.ldc = g.ldp;
    gs.alpha = 1.0;
    gs.beta = 0.0;
    gemm_epilogue(gs, acc, m0, n0, wm, wn, pr, lc);
    return;
  }
  gemm_epilogue(g, acc, m0, n0, wm, wn, pr, lc);
}

// plain-FMA check kernel with identical semantics (gps_set_option("gemm_impl", 1)); used by
// the GPU tests to validate the DMMA kernel independently of the oracle.
__global__ void gemm_nt_naive_kernel(const GemmArgs g) {
  int col = blockIdx.x * 16 + threadIdx.x;
  int row = blockIdx.y * 16 + threadIdx.y;
  if (row >= g.M || col >= g.N) return;
  if (g.c_uplo == C_LOWER && col > row) return;
  if (g.c_uplo == C_ROWMAP && (int64_t)col + g.coff > g.rowlim[row]) return;
  int klo = 0, khi = g.K;
  if (g.a_tri == TRI_LOWER) khi = min(khi, row + 1);
  if (g.a_tri == TRI_UPPER) klo = max(klo, row);
  if (g.b_tri == TRI_LOWER) khi = min(khi, col + 1);
  if (g.b_tri == TRI_UPPER) klo = max(klo, col);
  const double* a = g.A + (int64_t)row * g.lda;
  const double* b = g.B + (int64_t)col * g.ldb;
  double s = 0;
  for (int k = klo; k < khi; ++k) s = fma(a[k], b[k], s);
  double* cp = g.C + (int64_t)row * g.ldc + col;
  *cp = (g.beta != 0.0) ? fma(g.beta, *cp, g.alpha * s) : g.alpha * s;
}

// 2-D tensor map of a row-major operand view: dims (K, rows), row stride ld, box 16 x 128,
// 128-byte swizzle, out-of-bounds elements read as zero (edge tiles need no predicates).
// cuTensorMapEncodeTiled is taken from the driver through the runtime (no -lcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

bool make_tensor_map(CUtensorMap* tm, const Mat& X) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)X.cols, (cuuint64_t)X.rows};
  cuuint64_t strides[1] = {(cuuint64_t)X.ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)X.p, dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

double gemm_flops(const GemmArgs& g) {
  // algorithmic flops honouring the triangular structure (tile granularity is not counted)
  double full = 2.0 * (double)g.M * (double)g.N * (double)g.K;
  double f = full;
  // lower-masked C with M >= N (rows beneath ride along): N^2/2 + (M - N) N outputs
  if (g.c_uplo == C_LOWER && g.M >= g.N)
    f = 2.0 * (double)g.K * (0.5 * (double)g.N * (double)g.N + (double)(g.M - g.N) * (double)g.N);
  if (g.a_tri != TRI_NONE && g.b_tri != TRI_NONE)
    f = (g.c_uplo == C_LOWER) ? full / 6.0 : full / 3.0;
  else if (g.a_tri != TRI_NONE || g.b_tri != TRI_NONE)
    f *= 0.5;
  return f;
}

}  // namespace

int gps_gemm_nt_launch(gps_handle* h, double alpha, Mat A, Mat B, double beta, Mat C, int a_tri,
                       int b_tri, int c_uplo, const int64_t* rowlim, int64_t coff, double flops,
                       const int64_t* rowlo, int64_t lo_off, int lo_mode, const GemmBatch* batch) {
  if (A.cols != B.cols) return gps_fail(h, -3, "gemm_nt: K mismatch (%lld vs %lld)",
                                        (long long)A.cols, (long long)B.cols);
  if (C.rows != A.rows || C.cols != B.rows) return gps_fail(h, -6, "gemm_nt: C shape mismatch");
  if (C.rows == 0 || C.cols == 0) return 0;
  if (A.rows > INT32_MAX || B.rows > INT32_MAX || A.cols > INT32_MAX)
    return gps_fail(h, -3, "gemm_nt: dimension too large");
  GemmArgs g;
  g.A = A.p; g.B = B.p; g.C = C.p;
  g.lda = A.ld; g.ldb = B.ld; g.ldc = C.ld;
  g.M = (int)A.rows; g.N = (int)B.rows; g.K = (int)A.cols;
  g.alpha = alpha; g.beta = beta;
  g.a_tri = a_tri; g.b_tri = b_tri; g.c_uplo = c_uplo;
  g.tiles_m = (g.M + BM - 1) / BM;
  g.tiles_n = (g.N + BN - 1) / BN;
  g.rowlim = rowlim; g.coff = coff;
  g.rowlo = rowlo; g.lo_off = lo_off; g.lo_mode = rowlo ? lo_mode : 0;
  g.split_k = 1; g.k_chunk = 0; g.part = nullptr; g.ldp = 0; g.part_stride = 0;
  g.batch_y = 0; g.sAy = g.sBy = g.sCy = g.sAz = g.sBz = g.sCz = 0;
  int batch_z = 1;
  if (batch && batch->ny * batch->nz > 1) {
    if (batch->ny < 1 || batch->nz < 1 || batch->ny > 65535 || batch->nz > 65535)
      return gps_fail(h, -3, "gemm_nt: bad batch counts");
    g.batch_y = batch->ny; batch_z = batch->nz;
    g.sAy = batch->sAy; g.sBy = batch->sBy; g.sCy = batch->sCy;
    g.sAz = batch->sAz; g.sBz = batch->sBz; g.sCz = batch->sCz;
  }
  if (c_uplo == C_ROWMAP && !rowlim) return gps_fail(h, -7, "gemm_nt: C_ROWMAP needs a row-limit array");
  // split-K: a product with a long K and too few output tiles to fill the SMs -- the 1024 x 1024 x 8192
  // products of the SVGP backward are 64 (36 lower) tiles on 148 SMs, an M x R product with R <= 128 is
  // M/128 tiles -- is cut into K slices (blockIdx.y); the partial tiles go to a per-stream scratch and
  // one reduction pass applies alpha / beta and the lower-output mask.  One wave of CTAs at most.
  if (h->gemm_splitk && h->gemm_impl != 1 && (c_uplo == C_ALL || c_uplo == C_LOWER) && g.batch_y == 0 && g.K >= 256) {
    const int tiles = (c_uplo == C_LOWER && g.M == g.N) ? g.tiles_m * (g.tiles_m + 1) / 2 : g.tiles_m * g.tiles_n;
    // slices at least 128 deep for the small latency-bound products (one 128 x 128 tile of depth 512
    // keeps a single SM busy for 70 us), 256 deep otherwise; triangular operands are sliced too (a
    // slice outside a tile's K range just stores zeros)
    const int min_chunk = (tiles <= 32) ? 128 : 256;
    int split = h->sm_count / (tiles > 0 ? tiles : 1);
    if (split > g.K / min_chunk) split = g.K / min_chunk;
    if (split > 32) split = 32;
    if (split >= 2) {
      const int64_t ldp = ((int64_t)g.N + 15) / 16 * 16;
      g.k_chunk = ((g.K + split - 1) / split + BK - 1) / BK * BK;
      split = (g.K + g.k_chunk - 1) / g.k_chunk;      // no empty slices
      double* part = (double*)gps_ws_splitk(h, (size_t)split * g.M * ldp * sizeof(double));
      if (!part) return -102;
      g.split_k = split;
      g.part = part; g.ldp = ldp; g.part_stride = (int64_t)g.M * ldp;
    }
  }

  GemmEvent* ev = nullptr;
  if (h->profile) {
    if (h->events_used == h->events.size()) {
      GemmEvent e;
      if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess)
        return gps_fail(h, -100, "cudaEventCreate failed");
      e.flops = 0;
      h->events.push_back(e);
    }
    ev = &h->events[h->events_used++];
    ev->flops = flops >= 0.0 ? flops : gemm_flops(g) * (g.batch_y > 0 ? (double)g.batch_y * batch_z : 1.0);
    cudaEventRecord(ev->a, h->stream);
  }

  if (h->gemm_impl == 1) {
    dim3 grid((g.N + 15) / 16, (g.M + 15) / 16);
    for (int bz = 0; bz < batch_z; ++bz)
      for (int by = 0; by < (g.batch_y > 0 ? g.batch_y : 1); ++by) {
        GemmArgs gb = g;
        gb.A += by * g.sAy + bz * g.sAz;
        gb.B += by * g.sBy + bz * g.sBz;
        gb.C += by * g.sCy + bz * g.sCz;
        gemm_nt_naive_kernel<<<grid, dim3(16, 16), 0, h->stream>>>(gb);
      }
  } else {
    if (!h->attr_gemm) {
      cudaFuncSetAttribute(gemm_nt_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           GEMM_SMEM);
      cudaFuncSetAttribute(gemm_nt_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           GEMM_SMEM);
      h->attr_gemm = true;
    }
    bool vec16 = ((A.ld & 1) == 0) && ((B.ld & 1) == 0) &&
                 ((reinterpret_cast<uintptr_t>(A.p) & 15) == 0) &&
                 ((reinterpret_cast<uintptr_t>(B.p) & 15) == 0);
    if (g.batch_y > 0 && ((g.sAy | g.sBy | g.sAz | g.sBz) & 1)) vec16 = false;
    unsigned grid = (unsigned)(g.tiles_m * g.tiles_n);
    // gemm_impl 0: TMA kernel whenever the operands qualify; 2: force the cp.async kernel
    CUtensorMap tmA, tmB;
    bool tma = vec16 && h->gemm_impl == 0 && g.K > 0 && g.batch_y == 0 && make_tensor_map(&tmA, A) &&
               make_tensor_map(&tmB, B);
    const dim3 grid2(grid, (unsigned)(g.batch_y > 0 ? g.batch_y : g.split_k), (unsigned)batch_z);
    if (tma) {
      if (!h->attr_tma) {
        cudaFuncSetAttribute(gemm_nt_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM);
        h->attr_tma = true;
      }
      gemm_nt_tma_kernel<<<grid2, TMA_THREADS, TMA_SMEM, h->stream>>>(g, tmA, tmB);
    } else if (vec16) {
      gemm_nt_dmma_kernel<true><<<grid2, GEMM_THREADS, GEMM_SMEM, h->stream>>>(g);
    } else {
      gemm_nt_dmma_kernel<false><<<grid2, GEMM_THREADS, GEMM_SMEM, h->stream>>>(g);
    }
    if (g.split_k > 1) {
      h->launches++;
      splitk_reduce_kernel<<<dim3((unsigned)((g.N + 127) / 128), (unsigned)(g.M < 4096 ? g.M : 4096)), 128, 0,
                             h->stream>>>(g.part, g.ldp, g.part_stride, g.split_k, g.C, g.ldc, g.M, g.N,
                                          g.alpha, g.beta, c_uplo == C_LOWER, g.lo_mode == 2 ? g.rowlo : nullptr,
                                          g.lo_off);
    }
  }
  if (ev) cudaEventRecord(ev->b, h->stream);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

extern "C" int gps_gemm_nt(gps_handle* h, double alpha, const DLTensor* A, const DLTensor* B,
                           double beta, DLTensor* C, int a_tri, int b_tri, int c_uplo) {
  if (!h) return -1;
  Mat a, b, c;
  int rc;
  if ((rc = gps_as_mat(h, A, 3, "A", &a, false))) return rc;
  if ((rc = gps_as_mat(h, B, 4, "B", &b, false))) return rc;
  if ((rc = gps_as_mat(h, C, 6, "C", &c, false))) return rc;
  if (a_tri < 0 || a_tri > 2 || b_tri < 0 || b_tri > 2 || c_uplo < 0 || c_uplo > 1)
    return gps_fail(h, -7, "gemm_nt: bad tri/uplo flag");
  GPS_CUDA(h, cudaSetDevice(h->device));
  return gps_gemm_nt_launch(h, alpha, a, b, beta, c, a_tri, b_tri, c_uplo);
}

extern "C" int gps_gemm_nt_rowmap(gps_handle* h, double alpha, const DLTensor* A, const DLTensor* B,
                                  double beta, DLTensor* C, const DLTensor* row_limit,
                                  int64_t col_offset, double flops) {
  if (!h) return -1;
  Mat a, b, c;
  int rc;
  const int64_t* lim;
  if ((rc = gps_as_mat(h, A, 3, "A", &a, false))) return rc;
  if ((rc = gps_as_mat(h, B, 4, "B", &b, false))) return rc;
  if ((rc = gps_as_mat(h, C, 6, "C", &c, false))) return rc;
  if ((rc = gps_as_i64(h, row_limit, 7, "row_limit", c.rows, &lim))) return rc;
  GPS_CUDA(h, cudaSetDevice(h->device));
  return gps_gemm_nt_launch(h, alpha, a, b, beta, c, TRI_NONE, TRI_NONE, C_ROWMAP, lim, col_offset,
                            flops);
}
