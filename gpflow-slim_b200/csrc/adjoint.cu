// Library-side adjoints of the factorisation and of the triangular solve, and K^-1 from the
// factor -- the gps_potri / gps_chol_bwd / gps_trsm_bwd entry points sketched in SURVEY.md 8(b).
//
// TensorFlow differentiates tf.cholesky and tf.matrix_triangular_solve with its registered
// gradients (the reference never writes them down: optimizer.minimize, examples/gpr.py:53-54).
// Here each adjoint is "one triangular inverse U = L^-T plus two or three triangular-aware
// tensor-core GEMMs" (Murray 2016 for the Cholesky):
//   chol:  Abar = sym( U Phi(L^T Lbar) U^T ),  Phi = lower triangle with the diagonal halved
//   trsm:  X = B L^-T  =>  Bbar = Xbar L^-1 = Xbar U^T,   Lbar = -tril(Bbar^T X)
//   potri: K^-1 = U U^T (lower triangle)
// The same formulas exist as torch.autograd.Functions over gps_tri_inv_t / gps_gemm_nt in
// gpflowSlim/_backend/ops.py; these entry points do the whole adjoint in ONE call (one ctypes
// round trip instead of ~10 for the launch-bound SVGP step) and let a non-torch host bind them.
#include "internal.cuh"

int gps_inv_upper_full(gps_handle* h, Mat L, Mat U, const double* tinv);   // potrf.cu

namespace {

int64_t round16(int64_t x) { return (x + 15) / 16 * 16; }

// out = tril(src)^T  (upper triangular; the strict lower part of out is written as zero)
__global__ void tril_transpose_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ out,
                                      int64_t ldo, int64_t n) {
  __shared__ double tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < n && c < n && c <= r) ? src[r * lds + c] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = c0 + i, c = r0 + threadIdx.x;   // out[r][c] = tril(src)[c][r]
    if (r < n && c < n) out[r * ldo + c] = tile[threadIdx.x][i];
  }
}

// P <- Phi(P): strict upper part zeroed, diagonal halved
__global__ void phi_kernel(double* __restrict__ P, int64_t ld, int64_t n) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t r = blockIdx.y; r < n; r += gridDim.y) {
    if (c >= n) continue;
    if (c > r) P[r * ld + c] = 0.0;
    else if (c == r) P[r * ld + c] *= 0.5;
  }
}

// out = 1/2 (S + S^T)
__global__ void symmetrize_kernel(const double* __restrict__ S, int64_t lds, double* __restrict__ out,
                                  int64_t ldo, int64_t n) {
  __shared__ double tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {      // tile of S^T: rows c0.., columns r0..
    const int64_t r = c0 + i, c = r0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < n && c < n) ? S[r * lds + c] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < n && c < n) out[r * ldo + c] = 0.5 * (S[r * lds + c] + tile[threadIdx.x][i]);
  }
}

// U = L^-T: the caller's (gps_tri_inv_t output, zero strict lower part) or computed into WS_ADJ_U
int get_U(gps_handle* h, Mat L, const DLTensor* U_opt, int argidx, Mat* U) {
  int rc;
  const int64_t n = L.rows;
  if (U_opt) {
    if ((rc = gps_as_mat(h, U_opt, argidx, "U", U, false))) return rc;
    if (U->rows != n || U->cols != n) return gps_fail(h, -argidx, "U must be %lld x %lld", (long long)n, (long long)n);
    return 0;
  }
  const int64_t ld = round16(n);
  const int64_t nblk = (n + GPS_NB - 1) / GPS_NB;
  double* u = (double*)gps_ws(h, WS_ADJ_U, (size_t)n * ld * sizeof(double));
  double* tinv = (double*)gps_ws(h, WS_TINV, (size_t)nblk * GPS_NB * GPS_NB * sizeof(double));
  if (!u || !tinv) return -102;
  *U = Mat(u, n, n, ld);
  GPS_CUDA(h, cudaMemsetAsync(u, 0, (size_t)n * ld * sizeof(double), h->stream));
  if ((rc = gps_block_inverses(h, L, tinv))) return rc;
  return gps_inv_upper_full(h, L, *U, tinv);
}

int square_lower(gps_handle* h, const DLTensor* t, int argidx, const char* name, Mat* out) {
  int rc;
  if ((rc = gps_as_mat(h, t, argidx, name, out, false))) return rc;
  if (out->rows != out->cols) return gps_fail(h, -argidx, "argument %d (%s) must be square", argidx, name);
  return 0;
}

}  // namespace

extern "C" {

int gps_potri(gps_handle* h, const DLTensor* Lt, DLTensor* Kinv_out) {
  if (!h) return -1;
  Mat L, Kinv, U;
  int rc;
  if ((rc = square_lower(h, Lt, 2, "L", &L))) return rc;
  if ((rc = gps_as_mat(h, Kinv_out, 3, "Kinv_out", &Kinv, false))) return rc;
  if (Kinv.rows != L.rows || Kinv.cols != L.rows) return gps_fail(h, -3, "potri: Kinv_out shape mismatch");
  GPS_CUDA(h, cudaSetDevice(h->device));
  if (L.rows == 0) return 0;
  if ((rc = get_U(h, L, nullptr, 0, &U))) return rc;
  return gps_gemm_nt_launch(h, 1.0, U, U, 0.0, Kinv, TRI_UPPER, TRI_UPPER, C_LOWER);
}

int gps_chol_bwd(gps_handle* h, const DLTensor* Lt, const DLTensor* Lbar_t, const DLTensor* U_opt,
                 DLTensor* Abar_out) {
  if (!h) return -1;
  Mat L, Lbar, U, Abar;
  int rc;
  if ((rc = square_lower(h, Lt, 2, "L", &L))) return rc;
  if ((rc = gps_as_mat(h, Lbar_t, 3, "Lbar", &Lbar, false))) return rc;
  if ((rc = gps_as_mat(h, Abar_out, 5, "Abar_out", &Abar, false))) return rc;
  const int64_t n = L.rows;
  if (Lbar.rows != n || Lbar.cols != n || Abar.rows != n || Abar.cols != n)
    return gps_fail(h, -3, "chol_bwd: Lbar / Abar_out must be %lld x %lld", (long long)n, (long long)n);
  GPS_CUDA(h, cudaSetDevice(h->device));
  if (n == 0) return 0;
  if ((rc = get_U(h, L, U_opt, 4, &U))) return rc;
  const int64_t ld = round16(n);
  double* a = (double*)gps_ws(h, WS_ADJ_A, (size_t)n * ld * sizeof(double));
  double* b = (double*)gps_ws(h, WS_ADJ_B, (size_t)n * ld * sizeof(double));
  double* c = (double*)gps_ws(h, WS_ADJ_C, (size_t)n * ld * sizeof(double));
  if (!a || !b || !c) return -102;
  Mat LT(a, n, n, ld), LbarT(b, n, n, ld), P(c, n, n, ld), Qt(a, n, n, ld), S(b, n, n, ld);
  const dim3 tgrid((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32));
  if (tgrid.y > 65535) return gps_fail(h, -103, "chol_bwd: matrix too large");
  // L^T and tril(Lbar)^T: both upper triangular with explicit zeros below the diagonal
  tril_transpose_kernel<<<tgrid, dim3(32, 8), 0, h->stream>>>(L.p, L.ld, LT.p, LT.ld, n);
  GPS_LAUNCH_CHECK(h);
  tril_transpose_kernel<<<tgrid, dim3(32, 8), 0, h->stream>>>(Lbar.p, Lbar.ld, LbarT.p, LbarT.ld, n);
  GPS_LAUNCH_CHECK(h);
  // P = L^T Lbar  ((L^T)[m,k] (Lbar^T)[n',k] summed over k), then Phi
  if ((rc = gps_gemm_nt_launch(h, 1.0, LT, LbarT, 0.0, P, TRI_UPPER, TRI_UPPER, C_ALL))) return rc;
  phi_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)(n < 65535 ? n : 65535)), 256, 0, h->stream>>>(P.p, P.ld, n);
  GPS_LAUNCH_CHECK(h);
  // S = U P U^T:  Qt = U P^T,  S = U Qt^T
  if ((rc = gps_gemm_nt_launch(h, 1.0, U, P, 0.0, Qt, TRI_UPPER, TRI_LOWER, C_ALL))) return rc;
  if ((rc = gps_gemm_nt_launch(h, 1.0, U, Qt, 0.0, S, TRI_UPPER, TRI_NONE, C_ALL))) return rc;
  symmetrize_kernel<<<tgrid, dim3(32, 8), 0, h->stream>>>(S.p, S.ld, Abar.p, Abar.ld, n);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

int gps_trsm_bwd(gps_handle* h, const DLTensor* Lt, const DLTensor* X_t, const DLTensor* Xbar_t,
                 const DLTensor* U_opt, DLTensor* Bbar_out, DLTensor* Lbar_out) {
  if (!h) return -1;
  Mat L, X, Xbar, U, Bbar, Lbar;
  int rc;
  if ((rc = square_lower(h, Lt, 2, "L", &L))) return rc;
  if ((rc = gps_as_mat(h, X_t, 3, "X", &X, false))) return rc;
  if ((rc = gps_as_mat(h, Xbar_t, 4, "Xbar", &Xbar, false))) return rc;
  if ((rc = gps_as_mat(h, Bbar_out, 6, "Bbar_out", &Bbar, false))) return rc;
  const int64_t n = L.rows, m = X.rows;
  if (X.cols != n || Xbar.rows != m || Xbar.cols != n || Bbar.rows != m || Bbar.cols != n)
    return gps_fail(h, -3, "trsm_bwd: X, Xbar and Bbar_out must be %lld x %lld", (long long)m, (long long)n);
  if (Lbar_out) {
    if ((rc = gps_as_mat(h, Lbar_out, 7, "Lbar_out", &Lbar, false))) return rc;
    if (Lbar.rows != n || Lbar.cols != n) return gps_fail(h, -7, "trsm_bwd: Lbar_out must be %lld x %lld", (long long)n, (long long)n);
  }
  GPS_CUDA(h, cudaSetDevice(h->device));
  if (n == 0 || m == 0) {
    if (Lbar_out && n > 0)
      GPS_CUDA(h, cudaMemset2DAsync(Lbar.p, Lbar.ld * sizeof(double), 0, n * sizeof(double), n, h->stream));
    return 0;
  }
  if ((rc = get_U(h, L, U_opt, 5, &U))) return rc;
  // Bbar = Xbar L^-1 = Xbar U^T
  if ((rc = gps_gemm_nt_launch(h, 1.0, Xbar, U, 0.0, Bbar, TRI_NONE, TRI_UPPER, C_ALL))) return rc;
  if (!Lbar_out) return 0;
  // Lbar = -tril(Bbar^T X): an NT product of the two transposes, lower output only
  const int64_t ldm = round16(m);
  double* a = (double*)gps_ws(h, WS_ADJ_A, (size_t)n * ldm * sizeof(double));
  double* b = (double*)gps_ws(h, WS_ADJ_B, (size_t)n * ldm * sizeof(double));
  if (!a || !b) return -102;
  Mat BbarT(a, n, m, ldm), XT(b, n, m, ldm);
  if ((rc = gps_transpose_launch(h, Bbar, BbarT))) return rc;
  if ((rc = gps_transpose_launch(h, X, XT))) return rc;
  GPS_CUDA(h, cudaMemset2DAsync(Lbar.p, Lbar.ld * sizeof(double), 0, n * sizeof(double), n, h->stream));
  return gps_gemm_nt_launch(h, -1.0, BbarT, XT, 0.0, Lbar, TRI_NONE, TRI_NONE, C_LOWER);
}

}  // extern "C"
