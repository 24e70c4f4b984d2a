// Fused GPR objective + gradient and GPR prediction (models/gpr.py:55-72, :118-131,
// densities.py:73-95 of the reference) -- the whole evaluation stays on the device.
//
// Data layout in HBM (handle-owned, WS_ABUF):  an (N + R) x ld row-major matrix, ld = N rounded
// up to 16.  Rows 0..N-1 hold K + noise*I (lower triangle only), rows N..N+R-1 hold Yc^T.
// Factoring the leading N x N block in place and treating the R extra rows as part of every
// panel turns them into alpha^T = (L^-1 Yc)^T for free: no separate triangular solve for the
// quadratic form.  For the gradient, U = L^-T goes to a second N x ld buffer (WS_UBUF) and
// K^-1 = U U^T overwrites the lower triangle of the first; the weight matrix
// W = 1/2 (R K^-1 - beta beta^T) is never formed: the Gram backward kernel builds it per tile.
// Flops: N^3/3 (POTRF) + N^3/2 (U) + N^3/3 (U U^T) versus ~7 N^3 for TensorFlow's
// Cholesky-gradient route.
#include <math.h>

#include "internal.cuh"

int gps_inv_upper_full(gps_handle* h, Mat L, Mat U, const double* tinv);

namespace {

__global__ void put_yt_kernel(const double* __restrict__ Y, int64_t ldy, int64_t N, int R,
                              double* __restrict__ dst, int64_t ld) {
  // dst[r][i] = Y[i][r]
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  for (int r = 0; r < R; ++r) dst[(int64_t)r * ld + i] = Y[i * ldy + r];
}

__global__ void get_yt_kernel(const double* __restrict__ src, int64_t lds, int64_t N, int R,
                              double* __restrict__ Y, int64_t ldy) {
  // Y[i][r] = src[r][i]
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  for (int r = 0; r < R; ++r) Y[i * ldy + r] = src[(int64_t)r * lds + i];
}

// nlml = 0.5 N R log(2 pi) + R * sum(logdet parts) + 0.5 * sum alpha^2   (densities.py:92-94)
__global__ void nlml_kernel(const double* __restrict__ logdet, int64_t nblk,
                            const double* __restrict__ alpha, int64_t ld, int64_t N, int R,
                            double* __restrict__ out) {
  __shared__ double sm[2][32];
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = threadIdx.x; i < nblk; i += blockDim.x) s1 += logdet[i];
  for (int r = 0; r < R; ++r)
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
      double a = alpha[(int64_t)r * ld + i];
      s2 = fma(a, a, s2);
    }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    sm[0][threadIdx.x >> 5] = s1;
    sm[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double t1 = threadIdx.x < (blockDim.x >> 5) ? sm[0][threadIdx.x] : 0.0;
    double t2 = threadIdx.x < (blockDim.x >> 5) ? sm[1][threadIdx.x] : 0.0;
    t1 = warp_sum(t1);
    t2 = warp_sum(t2);
    if (threadIdx.x == 0)
      out[0] = 0.5 * (double)N * (double)R * 1.8378770664093453 + (double)R * t1 + 0.5 * t2;
  }
}

// var[i] = kdiag[i] - rowsumsq(At)[i]
__global__ void var_kernel(const double* __restrict__ kdiag, const double* __restrict__ ss,
                           int64_t n, double* __restrict__ var) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) var[i] = kdiag[i] - ss[i];
}

// Row-panel form of W = 1/2 (R K^-1 - beta beta^T) for a block-row distributed K^-1: row r of
// the panel is global row g = grow[r] and holds K^-1[g, j] for j >= c0 (c0 = first column of
// g's distribution block).  By symmetry, entries right of the diagonal block stand for (g, j)
// and (j, g): they weigh double; entries left of c0 are not this rank's and become zero.
__global__ void gpr_weight_rows_kernel(double* __restrict__ W, int64_t ldw, int64_t m, int64_t N,
                                       const int64_t* __restrict__ grow, const double* __restrict__ beta,
                                       int64_t ldb, int R, int64_t bs) {
  const int64_t r = blockIdx.y;
  if (r >= m) return;
  const int64_t g = grow[r];
  const int64_t c0 = g / bs * bs, c1 = c0 + bs;
  double bg[16];
  for (int q = 0; q < R; ++q) bg[q] = beta[(int64_t)q * ldb + g];
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N;
       j += (int64_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    if (j >= c0) {
      double bb = 0.0;
      for (int q = 0; q < R; ++q) bb = fma(bg[q], beta[(int64_t)q * ldb + j], bb);
      v = 0.5 * ((double)R * W[r * ldw + j] - bb);
      if (j >= c1) v *= 2.0;
    }
    W[r * ldw + j] = v;
  }
}

int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Build K + noise I (lower) and Yc^T into the work matrix and factor it.
int factor(gps_handle* h, const gps_kernel_desc* desc, const double* th, Mat X, Mat Yc, double noise,
           Mat* Aout, double** tinv_out, double** logdet_out, int** info_out, Mat* Ubuf) {
  int rc;
  const int64_t N = X.rows, R = Yc.cols;
  const int64_t ld = round_up(N, 16);
  const int64_t nblk = (N + GPS_NB - 1) / GPS_NB;
  double* a = (double*)gps_ws(h, WS_ABUF, (size_t)(N + R) * ld * sizeof(double));
  double* tinv = (double*)gps_ws(h, WS_TINV, (size_t)nblk * GPS_NB * GPS_NB * sizeof(double));
  double* logdet = (double*)gps_ws(h, WS_LOGDET, (size_t)nblk * sizeof(double));
  int* info = (int*)gps_ws(h, WS_INFO, 64);
  if (!a || !tinv || !logdet || !info) return -102;
  Mat A(a, N + R, N, ld);
  if ((rc = gps_gram_fwd_mat(h, desc, th, X, nullptr, noise, 1, A.sub(0, 0, N, N)))) return rc;
  put_yt_kernel<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>(Yc.p, Yc.ld, N, (int)R, a + N * ld, ld);
  GPS_LAUNCH_CHECK(h);
  GPS_CUDA(h, cudaMemsetAsync(info, 0, sizeof(int), h->stream));
  (void)Ubuf;
  if ((rc = gps_potrf_rec(h, A, N, R, 0, tinv, logdet, info, 0))) return rc;
  *Aout = A;
  *tinv_out = tinv;
  *logdet_out = logdet;
  *info_out = info;
  return 0;
}

int read_info(gps_handle* h, const int* info_dev, int* info_host) {
  if (!info_host) return 0;
  GPS_CUDA(h, cudaMemcpyAsync(info_host, info_dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  GPS_CUDA(h, cudaStreamSynchronize(h->stream));
  if (*info_host > 0) {
    gps_fail(h, *info_host, "Cholesky failed: leading minor of order %d is not positive definite", *info_host);
    return *info_host;
  }
  return 0;
}

}  // namespace

extern "C" {

int gps_gpr_nlml_fwd_bwd(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta,
                         const DLTensor* Xt, const DLTensor* Yt, double noise, int want_grad,
                         DLTensor* out_scalars, DLTensor* dtheta_out, DLTensor* dY_out, int* info_host) {
  if (!h) return -1;
  Mat th, X, Yc, outs, dth, dY;
  int rc;
  if ((rc = gps_as_mat(h, theta, 3, "theta", &th))) return rc;
  if (!desc) return gps_fail(h, -2, "null kernel descriptor");
  if (th.rows * th.cols < desc->n_theta) return gps_fail(h, -3, "theta too short");
  if ((rc = gps_as_mat(h, Xt, 4, "X", &X, false))) return rc;
  if ((rc = gps_as_mat(h, Yt, 5, "Yc", &Yc, false))) return rc;
  if (Yc.rows != X.rows) return gps_fail(h, -5, "Yc has %lld rows, X has %lld", (long long)Yc.rows, (long long)X.rows);
  if (Yc.cols < 1 || Yc.cols > 16) return gps_fail(h, -5, "Yc must have 1..16 columns");
  if ((rc = gps_as_mat(h, out_scalars, 8, "out_scalars", &outs))) return rc;
  if (outs.rows * outs.cols < 2) return gps_fail(h, -8, "out_scalars needs 2 entries");
  if (want_grad) {
    if ((rc = gps_as_mat(h, dtheta_out, 9, "dtheta_out", &dth))) return rc;
    if (dth.rows * dth.cols < desc->n_theta) return gps_fail(h, -9, "dtheta_out too small");
    if (dY_out) {
      if ((rc = gps_as_mat(h, dY_out, 10, "dY_out", &dY, false))) return rc;
      if (dY.rows != Yc.rows || dY.cols != Yc.cols) return gps_fail(h, -10, "dY_out shape mismatch");
    }
  }
  GPS_CUDA(h, cudaSetDevice(h->device));
  const int64_t N = X.rows, R = Yc.cols;
  if (N == 0) return gps_fail(h, -4, "X is empty");
  // Split-K pays where the products are few and latency-bound (N <= 8192: +7 % on BASELINE config C2);
  // at the headline size the extra partial-tile passes of the thousands of small products inside the
  // recursions cost 1.5 % of the step (bench line 0.945 -> 0.931 evals/s), so it is off from N = 16384 on.
  struct SplitkGuard {
    gps_handle* h; int saved;
    SplitkGuard(gps_handle* h_, bool off) : h(h_), saved(h_->gemm_splitk) { if (off) h->gemm_splitk = 0; }
    ~SplitkGuard() { h->gemm_splitk = saved; }
  } splitk_guard(h, N >= 16384);

  Mat A;
  double *tinv, *logdet;
  int* info;
  if ((rc = factor(h, desc, th.p, X, Yc, noise, &A, &tinv, &logdet, &info, nullptr))) return rc;
  const int64_t ld = A.ld;
  const int64_t nblk = (N + GPS_NB - 1) / GPS_NB;
  double* alpha = A.p + N * ld;   // [R][ld]
  nlml_kernel<<<1, 1024, 0, h->stream>>>(logdet, nblk, alpha, ld, N, (int)R, outs.p);
  GPS_LAUNCH_CHECK(h);
  if (!want_grad) return read_info(h, info, info_host);

  // U = L^-T
  double* u = (double*)gps_ws(h, WS_UBUF, (size_t)N * ld * sizeof(double));
  double* beta = (double*)gps_ws(h, WS_VEC, (size_t)R * N * sizeof(double));
  if (!u || !beta) return -102;
  Mat L = A.sub(0, 0, N, N);
  Mat U(u, N, N, ld);
  if ((rc = gps_inv_upper_full(h, L, U, tinv))) return rc;
  // beta_r = U alpha_r
  for (int r = 0; r < R; ++r)
    if ((rc = gps_gemv_launch(h, U, alpha + r * ld, beta + r * N, TRI_UPPER))) return rc;
  // K^-1 = U U^T  (lower) -> overwrites L
  if ((rc = gps_gemm_nt_launch(h, 1.0, U, U, 0.0, L, TRI_UPPER, TRI_UPPER, C_LOWER))) return rc;
  // dtheta = sum W dK/dtheta, dnoise = tr W
  GramW gw;
  gw.mode = W_GPR; gw.W = L; gw.beta = beta; gw.R = (int)R; gw.sym_lower = 1;
  if ((rc = gps_gram_bwd_mat(h, desc, th.p, X, nullptr, gw, dth.p, nullptr, outs.p + 1))) return rc;
  if (dY_out) {
    get_yt_kernel<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>(beta, N, N, (int)R, dY.p, dY.ld);
    GPS_LAUNCH_CHECK(h);
  }
  return read_info(h, info, info_host);
}

int gps_gpr_weight_rows(gps_handle* h, DLTensor* W_inout, const DLTensor* row_index,
                        const DLTensor* beta_t, int64_t block) {
  if (!h) return -1;
  Mat W, B;
  const int64_t* grow;
  int rc;
  if ((rc = gps_as_mat(h, W_inout, 2, "W_inout", &W, false))) return rc;
  if ((rc = gps_as_i64(h, row_index, 3, "row_index", W.rows, &grow))) return rc;
  if ((rc = gps_as_mat(h, beta_t, 4, "beta", &B, false))) return rc;
  if (B.cols != W.cols) return gps_fail(h, -4, "beta must be R x %lld", (long long)W.cols);
  if (B.rows < 1 || B.rows > 16) return gps_fail(h, -4, "beta must have 1..16 rows");
  if (block < 1) return gps_fail(h, -5, "block must be positive");
  GPS_CUDA(h, cudaSetDevice(h->device));
  if (W.rows == 0 || W.cols == 0) return 0;
  if (W.rows > 65535 * 64) return gps_fail(h, -2, "too many rows");
  for (int64_t r0 = 0; r0 < W.rows; r0 += 65535) {
    int64_t mr = W.rows - r0 < 65535 ? W.rows - r0 : 65535;
    dim3 grid((unsigned)((W.cols + 1023) / 1024 < 64 ? (W.cols + 1023) / 1024 : 64), (unsigned)mr);
    gpr_weight_rows_kernel<<<grid, 256, 0, h->stream>>>(W.p + r0 * W.ld, W.ld, mr, W.cols, grow + r0,
                                                        B.p, B.ld, (int)B.rows, block);
    GPS_LAUNCH_CHECK(h);
  }
  return 0;
}

int gps_gpr_predict(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta,
                    const DLTensor* Xt, const DLTensor* Yt, double noise, const DLTensor* Xnew_t,
                    int full_cov, DLTensor* mean_out, DLTensor* var_out, int* info_host) {
  if (!h) return -1;
  Mat th, X, Yc, Xn, mean, var;
  int rc;
  if ((rc = gps_as_mat(h, theta, 3, "theta", &th))) return rc;
  if (!desc) return gps_fail(h, -2, "null kernel descriptor");
  if (th.rows * th.cols < desc->n_theta) return gps_fail(h, -3, "theta too short");
  if ((rc = gps_as_mat(h, Xt, 4, "X", &X, false))) return rc;
  if ((rc = gps_as_mat(h, Yt, 5, "Yc", &Yc, false))) return rc;
  if ((rc = gps_as_mat(h, Xnew_t, 7, "Xnew", &Xn, false))) return rc;
  if ((rc = gps_as_mat(h, mean_out, 9, "mean_out", &mean, false))) return rc;
  if ((rc = gps_as_mat(h, var_out, 10, "var_out", &var))) return rc;
  if (Yc.rows != X.rows) return gps_fail(h, -5, "Yc/X row mismatch");
  if (Yc.cols < 1 || Yc.cols > 16) return gps_fail(h, -5, "Yc must have 1..16 columns");
  if (Xn.cols != X.cols) return gps_fail(h, -7, "Xnew column mismatch");
  const int64_t N = X.rows, R = Yc.cols, Ns = Xn.rows;
  if (mean.rows != Ns || mean.cols != R) return gps_fail(h, -9, "mean_out must be %lld x %lld", (long long)Ns, (long long)R);
  if (full_cov ? (var.rows != Ns || var.cols != Ns) : (var.rows * var.cols != Ns))
    return gps_fail(h, -10, "var_out shape mismatch");
  GPS_CUDA(h, cudaSetDevice(h->device));
  if (N == 0) return gps_fail(h, -4, "X is empty");

  Mat A;
  double *tinv, *logdet;
  int* info;
  if ((rc = factor(h, desc, th.p, X, Yc, noise, &A, &tinv, &logdet, &info, nullptr))) return rc;
  const int64_t ld = A.ld;
  Mat L = A.sub(0, 0, N, N);
  Mat V(A.p + N * ld, R, N, ld);   // V^T = alpha^T: [R][N]
  if (Ns == 0) return read_info(h, info, info_host);
  // At = K(Xnew, X) L^-T   ([N*, N]);  A = L^-1 Kx of the reference is At^T
  const int64_t ldt = round_up(N, 16);
  double* at = (double*)gps_ws(h, WS_TRSM, (size_t)Ns * ldt * sizeof(double));
  if (!at) return -102;
  Mat At(at, Ns, N, ldt);
  if ((rc = gps_gram_fwd_mat(h, desc, th.p, Xn, &X, 0.0, 0, At))) return rc;
  if ((rc = gps_trsm_rec(h, L, At, 0, tinv))) return rc;
  // fmean = At V     (models/gpr.py:124)
  if ((rc = gps_gemm_nt_launch(h, 1.0, At, V, 0.0, mean, TRI_NONE, TRI_NONE, C_ALL))) return rc;
  if (full_cov) {
    // K(Xnew) - At At^T   (models/gpr.py:126)
    if ((rc = gps_gram_fwd_mat(h, desc, th.p, Xn, nullptr, 0.0, 0, var))) return rc;
    if ((rc = gps_gemm_nt_launch(h, -1.0, At, At, 1.0, var, TRI_NONE, TRI_NONE, C_ALL))) return rc;
  } else {
    // Kdiag(Xnew) - colsum(A^2)   (models/gpr.py:130)
    double* tmp = (double*)gps_ws(h, WS_MISC, (size_t)2 * Ns * sizeof(double));
    if (!tmp) return -102;
    if ((rc = gps_kdiag_fwd_vec(h, desc, th.p, Xn, tmp))) return rc;
    if ((rc = gps_row_sumsq_launch(h, 1.0, At, 0.0, tmp + Ns))) return rc;
    var_kernel<<<(unsigned)((Ns + 255) / 256), 256, 0, h->stream>>>(tmp, tmp + Ns, Ns, var.p);
    GPS_LAUNCH_CHECK(h);
  }
  return read_info(h, info, info_host);
}

}  // extern "C"
