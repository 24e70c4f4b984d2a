// Internal declarations shared by the translation units of libgpslim_b200.so.
// Not part of the public ABI (that is include/gpslim_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <string>
#include <vector>

#include "gpslim_b200.h"

#define GPS_NB 128  // diagonal block / tile size of the blocked factorisations

// ----------------------------------------------------------------------------- host side
struct Mat {  // row-major matrix view
  double* p;
  int64_t rows, cols, ld;
  __host__ __device__ Mat() : p(nullptr), rows(0), cols(0), ld(0) {}
  __host__ __device__ Mat(double* p_, int64_t r, int64_t c, int64_t l) : p(p_), rows(r), cols(c), ld(l) {}
  __host__ __device__ Mat sub(int64_t r0, int64_t c0, int64_t nr, int64_t nc) const {
    return Mat(p + r0 * ld + c0, nr, nc, ld);
  }
};

enum WsSlot {
  WS_TINV = 0,    // block inverses of the diagonal blocks  [nblk][128][128]
  WS_LOGDET,      // per-block sum log diag
  WS_FEAT_L,      // Gram features of X
  WS_FEAT_R,      // Gram features of X2
  WS_PARTIAL,     // per-CTA partial sums of reductions
  WS_PARTIAL2,
  WS_ABUF,        // fused GPR: (N+R) x ld matrix  K -> L -> K^-1
  WS_UBUF,        // fused GPR: U = L^-T
  WS_VEC,         // fused GPR: beta etc.
  WS_TRSM,        // predict: A^T
  WS_MISC,
  WS_INFO,
  WS_THETA,
  WS_INVW,        // triangular inverse: W = -U11 L21^T of the current recursion node
  WS_INVT,        // triangular inverse: T22 = U22^T of the current recursion node
  WS_ROWLO,       // prefix solves: per-row first column (device copy)
  WS_ADJ_U,       // adjoint entry points (adjoint.cu): U = L^-T when the caller does not supply it
  WS_ADJ_A,       // ... three n x ld scratch matrices
  WS_ADJ_B,
  WS_ADJ_C,
  WS_SPLITK,      // (unused: split-K scratch is per stream, gps_ws_splitk)
  WS_LEAF_U,      // prefix solves: inverses of the aligned diagonal blocks of size trsm_leaf,
  WS_LEAF_T,      //   U_b = L_bb^-T (upper) and T_b = L_bb^-1 (lower), stacked [nblk][leaf][leaf]
  WS_LEAF_W,      //   per-level scratch of the batched inverse
  WS_LEAF_X,      //   output of a leaf product before it is copied back over its input
  WS_GRAM_G,      // Gram fast-path backward with input gradient: G = dObj/d(d2)  [N][ldM]
  WS_GRAM_B,      //   [F_R | 1]^T
  WS_GRAM_P,      //   P = G [F_R | 1]
  WS_COUNT
};

struct GemmEvent { cudaEvent_t a, b; double flops; };

struct gps_handle {
  int device = 0;
  cudaStream_t stream = 0;
  std::string err;
  int gemm_impl = 0;
  int leaf_impl = 0;   // 0 = blocked DMMA leaf, 1 = simple check kernel
  int gram_impl = 0;   // 0 = specialised kernels where they apply (single stationary kernel; NKN networks), 1 = interpreter only,
                       // 2 = experimental shared-memory-accumulator interpreter backward
  int profile = 0;
  int trsm_leaf = 512; // prefix solves: aligned diagonal blocks of this size (a power-of-two multiple of 128)
                       // are solved by ONE product with their explicit inverse; 128 = strips only
  int gemm_splitk = 1; // split-K for long-K products with few output tiles (0 switches it off)
  // function attributes (opt-in shared memory sizes) are per device: one flag set per handle
  bool attr_gemm = false, attr_tma = false, attr_potrf = false;
  // split-K partial tiles, one buffer per stream (a handle may be driven from several streams)
  std::map<cudaStream_t, std::pair<void*, size_t>> splitk_ws;
  // small scratch slots that calls running concurrently on different streams of one handle must
  // not share (block inverses, info flag, row tables): kept per (slot, stream)
  std::map<std::pair<int, cudaStream_t>, std::pair<void*, size_t>> stream_ws;
  void* ws_ptr[WS_COUNT] = {};
  size_t ws_bytes[WS_COUNT] = {};
  std::vector<GemmEvent> events;   // pool
  size_t events_used = 0;
  double gemm_ms_acc = 0, gemm_flops_acc = 0;
  int64_t launches = 0;
  int sm_count = 148;
};

int gps_fail(gps_handle* h, int code, const char* fmt, ...);
void* gps_ws(gps_handle* h, int slot, size_t bytes);  // nullptr on failure (err set)
void* gps_ws_splitk(gps_handle* h, size_t bytes);      // per-stream (h->stream) scratch

#define GPS_CUDA(h, expr)                                                          \
  do {                                                                             \
    cudaError_t e__ = (expr);                                                      \
    if (e__ != cudaSuccess)                                                        \
      return gps_fail((h), -100, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                      __FILE__, __LINE__);                                         \
  } while (0)

#define GPS_LAUNCH_CHECK(h)                                                        \
  do {                                                                             \
    (h)->launches++;                                                               \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess)                                                        \
      return gps_fail((h), -101, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                      __FILE__, __LINE__);                                         \
  } while (0)

// argument checking: returns 0 or sets the error and returns -(argidx)
int gps_as_mat(gps_handle* h, const DLTensor* t, int argidx, const char* name, Mat* out,
               bool allow_vec = true);
// contiguous int64 device vector with exactly n entries
int gps_as_i64(gps_handle* h, const DLTensor* t, int argidx, const char* name, int64_t n,
               const int64_t** out);

// ----------------------------------------------------------------------------- GEMM family
enum { TRI_NONE = 0, TRI_LOWER = 1, TRI_UPPER = 2 };
enum { C_ALL = 0, C_LOWER = 1, C_ROWMAP = 2 };

// strided batch of independent products: member (y, z) uses A + y*sAy + z*sAz etc.
struct GemmBatch { int ny, nz; int64_t sAy, sBy, sCy, sAz, sBz, sCz; };

// C = alpha * A * B^T + beta * C   (A: MxK, B: NxK, C: MxN, all row-major)
// c_uplo == C_ROWMAP: element (r, c) is updated iff c + coff <= rowlim[r] (device array,
// non-decreasing).  flops >= 0 overrides the profile's flop count for this launch.
int gps_gemm_nt_launch(gps_handle* h, double alpha, Mat A, Mat B, double beta, Mat C, int a_tri,
                       int b_tri, int c_uplo, const int64_t* rowlim = nullptr, int64_t coff = 0,
                       double flops = -1.0, const int64_t* rowlo = nullptr, int64_t lo_off = 0,
                       int lo_mode = 0, const GemmBatch* batch = nullptr);

// ----------------------------------------------------------------------------- factorisation
// Factor the n x n block at A (lower, in place) and solve the `below` rows under it:
// A[n:n+below, 0:n] <- A[n:n+below, 0:n] L^-T.  blk0 = index of the first 128-block (into
// the Tinv / logdet workspaces).  info_dev: device int, first failing 1-based order (0 = ok).
int gps_potrf_rec(gps_handle* h, Mat A, int64_t n, int64_t below, int64_t blk0, double* tinv,
                  double* logdet, int* info_dev, int64_t row0);
// B (m x n) <- B L^-T for an n x n lower L whose block inverses are tinv[blk0...]
int gps_trsm_rec(gps_handle* h, Mat L, Mat B, int64_t blk0, const double* tinv);
// block inverses of an already-factored L (all diagonal 128-blocks, one launch)
int gps_block_inverses(gps_handle* h, Mat L, double* tinv);
// U = L^-T (upper; diagonal tiles fully written incl. zero lower part; off-diagonal lower
// tiles are NOT touched)
int gps_inv_upper_rec(gps_handle* h, Mat L, Mat U, int64_t blk0, const double* tinv);
int gps_sum_partials(gps_handle* h, const double* parts, int64_t n, double scale, double* out,
                     int accumulate);

// ----------------------------------------------------------------------------- misc kernels
int gps_transpose_launch(gps_handle* h, Mat A, Mat At);
int gps_zero_upper_launch(gps_handle* h, Mat A);
int gps_fill_launch(gps_handle* h, double* p, int64_t n, double v);
// y[i] = sum_{k in range(i)} A[i,k] x[k]; tri: TRI_NONE all k, TRI_UPPER k>=i, TRI_LOWER k<=i
int gps_gemv_launch(gps_handle* h, Mat A, const double* x, double* y, int tri);
int gps_row_sumsq_launch(gps_handle* h, double alpha, Mat A, double beta, double* out);
int gps_sumsq_launch(gps_handle* h, const double* x, int64_t n, double scale, double* out,
                     int accumulate);

// ----------------------------------------------------------------------------- Gram
struct GramPlan;  // gram.cu
int gps_gram_fwd_mat(gps_handle* h, const gps_kernel_desc* desc, const double* theta_dev, Mat X,
                     const Mat* X2, double diag_add, int uplo, Mat K);
// W modes for the backward contraction
enum { W_DENSE = 0, W_GPR = 1 };
struct GramW {
  int mode;
  Mat W;               // W_DENSE: the weights; W_GPR: K^-1 (lower)
  const double* beta;  // W_GPR: [R][N] (row r = beta_r)
  int R;
  int sym_lower;       // 1: only the lower triangle of W is valid (symmetric problem)
};
int gps_gram_bwd_mat(gps_handle* h, const gps_kernel_desc* desc, const double* theta_dev, Mat X,
                     const Mat* X2, GramW w, double* dtheta_out, Mat* dX, double* trace_out);
int gps_kdiag_fwd_vec(gps_handle* h, const gps_kernel_desc* desc, const double* theta_dev, Mat X,
                      double* out);

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte async copy global->shared; src_bytes in {0,8,16}, the rest is zero-filled
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
