/*
 * gpslim_b200.h -- C ABI of libgpslim_b200.so: hand-written sm_100a CUDA kernels for the
 * Gaussian-process inference hot path of GPflow-Slim
 *     Gram build -> jittered FP64 Cholesky -> triangular solves -> NLML/ELBO + gradient
 *     -> predictive mean / variance.
 *
 * The reference (ssydasheng/GPflow-Slim) has no FFI of its own: it calls TensorFlow ops
 * directly.  Each entry point below therefore cites the reference call site(s) (file:line
 * relative to the reference's gpflowSlim/ package) whose TensorFlow op(s) it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - Every tensor is a BORROWED, non-owning DLPack `DLTensor` view (ABI of dlpack.h v0.8,
 *     restated below): device_type kDLCUDA, dtype float64, ndim 1 or 2, innermost stride 1
 *     (row-major with an arbitrary leading dimension).  Nothing is retained after a call.
 *   - Outputs are caller-allocated.  Scratch memory is owned by the handle, grown lazily and
 *     freed by gps_destroy().
 *   - Every function returns an int status: 0 ok; <0 = -(index of the offending argument,
 *     1-based) for a bad dtype/shape/stride/device; >0 = LAPACK-style `info` (the leading
 *     minor of that order is not positive definite).  gps_last_error(h) gives the text.
 *     No C++ exception crosses this boundary.
 *   - Calls enqueue work on the handle's stream (gps_set_stream) and return without a host
 *     synchronisation unless a host scalar or `info` is requested (documented per function).
 *     A handle is not re-entrant; distinct handles are independent.
 *   - "lower" matrices: only the lower triangle (row >= col) is read / written; the strict
 *     upper triangle of an in-place factor is left untouched.
 */
#ifndef GPSLIM_B200_H_
#define GPSLIM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- DLPack tensor view (layout-identical to dlpack.h v0.8 `DLTensor`) -------------- */
#ifndef DLPACK_DLPACK_H_
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;   /* kDLCUDA == 2 */
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType; /* kDLFloat == 2 */
typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides;      /* in elements; NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;
#endif

typedef struct gps_handle gps_handle;

/* ---- kernel descriptor ---------------------------------------------------------------
 * A covariance function is a list of PRIMITIVES (kernels.py: Stationary :360-429, RBF
 * :432-439, Exponential/Matern12/32/52 :555-610, Linear :474-510, Periodic :769-819) plus
 * a straight-line PROGRAM over value slots that combines them (kernels.py Sum/Product
 * :1071-1084 incl. scalar constants; neural_kernel_network_wrapper.py Linear :90-129,
 * Product :132-155).  Slot p (p < n_prims) holds primitive p's value; ops write new slots;
 * the result is slot `out_slot`.  All parameter VALUES (constrained) live in one device
 * vector `theta`; the descriptor only stores offsets into it.
 *
 * theta layout of a primitive at theta_off:
 *   stationary: [variance, lengthscale[0..nls)]   nls = ard ? ndims : 1
 *   linear    : [variance[0..nv)]                 nv  = ard ? ndims : 1
 *   periodic  : [variance, lengthscale, period]
 */
#define GPS_MAX_PRIMS 16
#define GPS_MAX_DIMS 32
#define GPS_MAX_OPS 48
#define GPS_MAX_SLOTS 96
#define GPS_MAX_THETA 512

enum gps_prim_type {
  GPS_RBF = 0, GPS_EXPONENTIAL = 1, GPS_MATERN12 = 2, GPS_MATERN32 = 3, GPS_MATERN52 = 4,
  GPS_LINEAR = 5, GPS_PERIODIC = 6
};

enum gps_op_type {
  GPS_OP_CONST = 0,   /* slot[dst] = theta[a]                                  (kernels.py:1060-1063) */
  GPS_OP_ADD = 1,     /* slot[dst] = slot[a] + slot[b]                         (Sum  :1071-1076) */
  GPS_OP_MUL = 2,     /* slot[dst] = slot[a] * slot[b]                         (Product :1079-1084) */
  GPS_OP_COPY = 3,    /* slot[dst] = slot[a] */
  GPS_OP_LINEAR = 4,  /* slot[dst+o] = sum_i theta[c+o*b+i]*slot[a+i] + theta[d+o], o<n_out
                         a=src slot, b=n_in, c=weights offset, d=bias offset, n=n_out (wrapper.py:114-115) */
  GPS_OP_PRODUCT = 5  /* slot[dst+g] = prod_{s<b} slot[a+g*b+s], g<n            (wrapper.py:142-145) */
};

typedef struct {
  int32_t type;                 /* gps_prim_type */
  int32_t ndims;                /* input_dim of the primitive */
  int32_t ard;                  /* 1: one lengthscale (or Linear variance) per dimension */
  int32_t theta_off;
  int32_t dims[GPS_MAX_DIMS];   /* active columns of X (kernels.py:217-253 _slice) */
} gps_prim;

typedef struct { int32_t op, dst, a, b, c, d, n, pad; } gps_op;

typedef struct {
  int32_t n_prims, n_ops, n_theta, out_slot;
  gps_prim prims[GPS_MAX_PRIMS];
  gps_op ops[GPS_MAX_OPS];
} gps_kernel_desc;

/* ---- handle --------------------------------------------------------------------------*/
int gps_create(int device, gps_handle** out);
int gps_destroy(gps_handle* h);
int gps_set_stream(gps_handle* h, void* cuda_stream);      /* cudaStream_t, 0 = legacy default */
const char* gps_last_error(gps_handle* h);
int gps_version(void);
/* options: "gemm_impl" 0 = DMMA tensor-core GEMM, operands staged by TMA when they are 16-byte
 *                          aligned with even leading dimensions, else by cp.async (default);
 *                      1 = plain-FMA check kernel; 2 = always the cp.async DMMA kernel;
 *          "gram_impl" 0 = specialised Gram kernels where they apply (default): register-tiled for a
 *                          single stationary covariance, tensor-core (mma.sync f64) kernels for
 *                          Linear/Product(2) neural-kernel networks; interpreter otherwise,
 *                          1 = the generic interpreter kernels only,
 *                          2 = interpreter kernels with their slot / accumulator arrays in shared
 *                          memory (correct, 7 % faster on the NKN config, not the default;
 *                          tests/test_gpu_switches.py holds it to the default path);
 *          "gemm_splitk" 1 (default) = split-K: products with fewer output tiles than SMs are cut
 *                          into up to 32 K slices (both tensor-core kernels; partial tiles in a
 *                          per-stream scratch + one reduction pass); 0 = off;
 *          "trsm_leaf" 128..4096 (power of two, default 512): the prefix solves handle aligned
 *                          diagonal blocks of this size by one product with their explicit inverse;
 *          "leaf_impl" 0 = blocked DMMA 128x128 Cholesky leaf (default), 1 = scalar check kernel;
 *          "profile"   1 = bracket every GEMM-class launch with CUDA events. */
int gps_set_option(gps_handle* h, const char* name, int64_t value);
/* Sums since the last reset: milliseconds and algorithmic flops of the DMMA GEMM launches,
 * number of kernels this library launched.  Synchronises the stream. */
int gps_profile_read(gps_handle* h, double* gemm_ms, double* gemm_flops, int64_t* launches,
                     int reset);

/* ---- Gram matrices -------------------------------------------------------------------
 * gps_gram_fwd: K[i,j] = k(X[i,:], X2[j,:]) (+ diag_add on i==j when X2 is NULL).
 *   Replaces Kernel.K(X, X2): square_dist kernels.py:408-421 (scale by lengthscale,
 *   -2XX'^T + norms, clip at 0), euclid_dist :424-426 (sqrt(d+1e-12)), the exp/Matern/Linear/
 *   Periodic bodies and the Sum/Product/NKN composition -- one fused kernel, the distance
 *   matrix is never materialised.  X2 == NULL is the symmetric case; uplo = 1 then writes
 *   only the lower triangle.  `diag_add` replaces `+ eye(N)*likelihood.variance`
 *   (models/gpr.py:69,120) and `+ jitter*eye` (features.py:76, conditionals.py:60).
 * gps_gram_bwd: dtheta[t] = sum_ij W[i,j] dK[i,j]/dtheta[t]; optionally dX (gradient w.r.t.
 *   the FIRST argument's rows, e.g. inducing inputs Z).  For X2 == NULL, W is taken as
 *   symmetric and both roles of X contribute to dX.  Replaces TensorFlow's autodiff of the
 *   ops above (optimizer.minimize, examples/gpr.py:53-54).  K is recomputed tile by tile.
 * gps_kdiag_fwd/bwd: Kernel.Kdiag (kernels.py:428-429, :507-510, :803-804).
 */
int gps_gram_fwd(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta,
                 const DLTensor* X, const DLTensor* X2, double diag_add, int uplo,
                 DLTensor* K_out);
int gps_gram_bwd(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta,
                 const DLTensor* X, const DLTensor* X2, const DLTensor* W,
                 DLTensor* dtheta_out, DLTensor* dX_out /* may be NULL */);
int gps_kdiag_fwd(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta,
                  const DLTensor* X, DLTensor* out);
int gps_kdiag_bwd(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta,
                  const DLTensor* X, const DLTensor* w, DLTensor* dtheta_out,
                  DLTensor* dX_out /* may be NULL */);

/* ---- dense factor / solve / multiply ---------------------------------------------------
 * gps_potrf: in-place lower Cholesky A = L L^T (tf.cholesky: models/gpr.py:70,121,
 *   conditionals.py:84, kullback_leiblers.py:53, models/sgpr.py:136,143,170,173).  Blocked
 *   recursive right-looking: 128x128 diagonal blocks factored (and inverted) by one CTA,
 *   panels solved against the block inverse, trailing SYRK/GEMM updates on FP64 tensor
 *   cores (DMMA).  If `info_host` is non-NULL the call synchronises and stores LAPACK info.
 *   zero_upper != 0 additionally zeroes the strict upper triangle (tf.cholesky's output).
 * gps_trsm_rlt: B <- B L^-T  (row-major restatement of A = L^-1 B^T: the
 *   tf.matrix_triangular_solve(L, ., lower=True) call sites densities.py:82,
 *   models/gpr.py:122-123, conditionals.py:87, kullback_leiblers.py:54,93,
 *   models/sgpr.py:140,145,171,175-177).
 * gps_tri_inv_t: U = L^-T (upper triangular, strict lower part zeroed).  Used for the
 *   transposed solve conditionals.py:100 and by the backward kernels.
 * gps_gemm_nt: C = alpha * A B^T + beta * C on FP64 tensor cores; a_tri / b_tri (0 none,
 *   1 lower, 2 upper) declare A / B triangular so that zero tiles are skipped; c_uplo
 *   (0 all, 1 lower only).  (tf.matmul call sites: conditionals.py:90,103,111,
 *   models/gpr.py:124,127, models/sgpr.py:141,144.)
 * gps_transpose: Bt = A^T.
 */
int gps_potrf(gps_handle* h, DLTensor* A_inout, int zero_upper, int* info_host);
int gps_trsm_rlt(gps_handle* h, const DLTensor* L, DLTensor* B_inout);
int gps_tri_inv_t(gps_handle* h, const DLTensor* L, DLTensor* U_out);
int gps_gemm_nt(gps_handle* h, double alpha, const DLTensor* A, const DLTensor* B, double beta,
                DLTensor* C, int a_tri, int b_tri, int c_uplo);
int gps_transpose(gps_handle* h, const DLTensor* A, DLTensor* At_out);

/* ---- adjoints of the factorisation / the triangular solve, and K^-1 from the factor ------
 * (the gps_potri / gps_chol_bwd / gps_trsm_bwd of the SURVEY.md 8(b) sketch; csrc/adjoint.cu.)
 * The reference has no such code: TensorFlow's registered gradients of tf.cholesky and
 * tf.matrix_triangular_solve run when optimizer.minimize differentiates the objective
 * (examples/gpr.py:53-54, examples/svgp.py:160-161).  L must be lower triangular WITH zeros above
 * the diagonal (gps_potrf with zero_upper = 1).  `U` is optional everywhere: pass the output of
 * gps_tri_inv_t(L) to share one triangular inverse between several adjoints of the same factor,
 * or NULL to have it computed into the handle's workspace.
 * gps_potri:    Kinv_out (lower triangle only; the strict upper part is not written) = (L L^T)^-1.
 * gps_chol_bwd: Abar_out = sym(L^-T Phi(L^T tril(Lbar)) L^-1), the adjoint of L = chol(A)
 *               (tf.cholesky call sites: models/gpr.py:70,121, conditionals.py:84,
 *               kullback_leiblers.py:53, models/sgpr.py:136,143,170,173).
 * gps_trsm_bwd: for X = B L^-T (gps_trsm_rlt):  Bbar_out = Xbar L^-1  and, if Lbar_out is not
 *               NULL,  Lbar_out = -tril(Bbar^T X)  (the tf.matrix_triangular_solve call sites
 *               densities.py:82, conditionals.py:87, kullback_leiblers.py:54,93, ...).
 * Status: written at the end of round 1 without GPU access; exercised through the CPU build of
 * the library (tests/test_library_on_cpu.py); the Python layer calls them only when
 * ops.FUSED_ADJOINTS is switched on. */
int gps_potri(gps_handle* h, const DLTensor* L, DLTensor* Kinv_out);
int gps_chol_bwd(gps_handle* h, const DLTensor* L, const DLTensor* Lbar, const DLTensor* U /* may be NULL */,
                 DLTensor* Abar_out);
int gps_trsm_bwd(gps_handle* h, const DLTensor* L, const DLTensor* X, const DLTensor* Xbar,
                 const DLTensor* U /* may be NULL */, DLTensor* Bbar_out, DLTensor* Lbar_out /* may be NULL */);

/* ---- building blocks of the block-row distributed GPR (one process per GPU) ---------------
 * The multi-GPU path (gpflowSlim/_backend/dist_gpr.py) distributes the rows of K over the
 * ranks in blocks; NCCL moves the panels, these entry points do the arithmetic.
 * gps_gemm_nt_rowmap: C[r,c] = alpha (A B^T)[r,c] + beta C[r,c] only where
 *   c + col_offset <= row_limit[r]  (row_limit: int64 device vector, one entry per row of C,
 *   non-decreasing = global row index of each local row).  This is the trailing update
 *   A22 -= L21 L21^T of tf.cholesky (models/gpr.py:70) restricted to the lower triangle when
 *   the rows of A22 are a block-cyclic subset.  Tiles wholly above the limit are skipped.
 *   `flops` (>= 0) is what the profile option books for the launch.
 * gps_trsm_rlt_prefix: B <- B L^-T where row r of B is identically zero left of column
 *   row_start[r] (HOST int64 array, one entry per row of B, non-decreasing multiples of 128;
 *   NULL = all zeros, i.e. the plain solve).  Rows that have not started yet are skipped and
 *   the leading zero K-range of every update tile is never read: with B = rows of the
 *   identity sorted by column this yields rows of U = L^-T at their true flop count.
 * gps_trsm_rln_prefix: B <- B L^-1 (the solve against the untransposed factor), columns right
 *   to left; row r only WANTS the columns >= row_start[r] (what lies left of it is left
 *   undefined).  Lt = L^T (row-major upper) must be supplied so that every product is
 *   K-contiguous.  Rows of U go in, rows of K^-1 = L^-T L^-1 come out (the O(N^3) part of
 *   TensorFlow's Cholesky gradient, examples/gpr.py:53-54).
 * gps_gpr_weight_rows: W[r,j] <- m(r,j)/2 (R K^-1[g_r,j] - sum_q beta_q[g_r] beta_q[j]) in
 *   place on a row panel of K^-1; g_r = row_index[r]; m = 0 left of g_r's block (of size
 *   `block`), 1 inside it, 2 to the right (symmetric counterpart).
 */
int gps_gemm_nt_rowmap(gps_handle* h, double alpha, const DLTensor* A, const DLTensor* B,
                       double beta, DLTensor* C, const DLTensor* row_limit, int64_t col_offset,
                       double flops);
int gps_trsm_rlt_prefix(gps_handle* h, const DLTensor* L, DLTensor* B_inout,
                        const int64_t* row_start);
int gps_trsm_rln_prefix(gps_handle* h, const DLTensor* L, const DLTensor* Lt, DLTensor* B_inout,
                        const int64_t* row_start);
int gps_gpr_weight_rows(gps_handle* h, DLTensor* W_inout, const DLTensor* row_index,
                        const DLTensor* beta, int64_t block);

/* ---- reductions on the hot path ----------------------------------------------------------
 * gps_sum_log_diag: out[0] = sum_i log(L[i,i])           (densities.py:93)
 * gps_row_sumsq:    out[i] = beta*out[i] + alpha * sum_j A[i,j]^2
 *                   (conditionals.py:94,118, models/gpr.py:130, models/sgpr.py:185-186)
 */
int gps_sum_log_diag(gps_handle* h, const DLTensor* L, DLTensor* out);
int gps_row_sumsq(gps_handle* h, double alpha, const DLTensor* A, double beta, DLTensor* out);

/* ---- fused GPR objective ------------------------------------------------------------------
 * gps_gpr_nlml_fwd_bwd: the whole of GPR._build_likelihood (models/gpr.py:55-72) +
 *   densities.multivariate_normal (densities.py:73-95) + Model.objective (models/model.py:
 *   67-73) and its gradient, on the device:
 *     K = k(X,X) + noise*I ; L = chol(K) ; alpha = L^-1 Yc ;
 *     nlml = N R/2 log(2 pi) + R sum log L_ii + 1/2 sum alpha^2
 *     dnlml/dtheta = sum_ij W_ij dK_ij/dtheta,  W = 1/2 (R K^-1 - beta beta^T), beta = L^-T alpha
 *     dnlml/dnoise = tr W ;  dnlml/dYc = beta.
 *   Yc = Y - mean_function(X), [N, R].  Outputs are device tensors: out_scalars[0] = nlml,
 *   out_scalars[1] = dnlml/dnoise; dtheta_out [n_theta]; dY_out [N,R] or NULL.  When
 *   want_grad == 0 only nlml is produced (N^3/3 flops instead of N^3).  `info_host`, if
 *   non-NULL, synchronises and receives the Cholesky info.
 * gps_gpr_predict: GPR._build_predict (models/gpr.py:118-131): mean_out [N*,R] =
 *   A^T V, var_out [N*] = Kdiag(Xnew) - colsum(A^2) (full_cov == 0) or [N*,N*] =
 *   K(Xnew) - A^T A (full_cov == 1), with A = L^-1 K(X,Xnew), V = L^-1 Yc.
 */
int gps_gpr_nlml_fwd_bwd(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta,
                         const DLTensor* X, const DLTensor* Yc, double noise, int want_grad,
                         DLTensor* out_scalars, DLTensor* dtheta_out, DLTensor* dY_out,
                         int* info_host);
int gps_gpr_predict(gps_handle* h, const gps_kernel_desc* desc, const DLTensor* theta,
                    const DLTensor* X, const DLTensor* Yc, double noise, const DLTensor* Xnew,
                    int full_cov, DLTensor* mean_out, DLTensor* var_out, int* info_host);

#ifdef __cplusplus
}
#endif
#endif /* GPSLIM_B200_H_ */
