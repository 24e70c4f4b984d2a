// Handle, workspace, error reporting, argument validation and the small utility kernels
// (transpose, fills, triangular GEMV, row / vector reductions).
#include <stdarg.h>
#include <string.h>

#include "internal.cuh"

int gps_fail(gps_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

static bool per_stream_slot(int slot) {
  return slot == WS_TINV || slot == WS_INFO || slot == WS_LOGDET || slot == WS_ROWLO;
}

void* gps_ws(gps_handle* h, int slot, size_t bytes) {
  if (per_stream_slot(slot) && h->stream != 0) {
    // the handle's first (default / legacy) stream keeps the plain slot; side streams get their own
    auto& e = h->stream_ws[std::make_pair(slot, h->stream)];
    if (e.first && bytes <= e.second) return e.first;
    cudaStreamSynchronize(h->stream);
    if (e.first) cudaFree(e.first);
    e = {nullptr, 0};
    size_t want = bytes + bytes / 8 + 256;
    void* p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) {
      cudaGetLastError();
      gps_fail(h, -102, "workspace allocation of %zu bytes failed", bytes);
      return nullptr;
    }
    e = {p, want};
    return p;
  }
  if (bytes <= h->ws_bytes[slot] && h->ws_ptr[slot]) return h->ws_ptr[slot];
  // in-flight work may still use the old buffer
  cudaStreamSynchronize(h->stream);
  if (h->ws_ptr[slot]) cudaFree(h->ws_ptr[slot]);
  h->ws_ptr[slot] = nullptr;
  h->ws_bytes[slot] = 0;
  size_t want = bytes + bytes / 8 + 256;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&p, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    gps_fail(h, -102, "workspace allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  h->ws_ptr[slot] = p;
  h->ws_bytes[slot] = want;
  return p;
}

void* gps_ws_splitk(gps_handle* h, size_t bytes) {
  auto& slot = h->splitk_ws[h->stream];
  if (slot.first && bytes <= slot.second) return slot.first;
  cudaStreamSynchronize(h->stream);
  if (slot.first) cudaFree(slot.first);
  slot = {nullptr, 0};
  size_t want = bytes + bytes / 4 + 256;
  void* p = nullptr;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    cudaGetLastError();
    gps_fail(h, -102, "split-K workspace allocation of %zu bytes failed", bytes);
    return nullptr;
  }
  slot = {p, want};
  return p;
}

int gps_as_mat(gps_handle* h, const DLTensor* t, int argidx, const char* name, Mat* out,
               bool allow_vec) {
  if (!t || !t->data)
    return gps_fail(h, -argidx, "argument %d (%s): null tensor", argidx, name);
  if (t->device.device_type != 2 /*kDLCUDA*/)
    return gps_fail(h, -argidx, "argument %d (%s): not a CUDA tensor (device_type %d)", argidx,
                    name, t->device.device_type);
  if (t->device.device_id != h->device)
    return gps_fail(h, -argidx, "argument %d (%s): on device %d, handle is on device %d", argidx,
                    name, t->device.device_id, h->device);
  if (t->dtype.code != 2 || t->dtype.bits != 64 || t->dtype.lanes != 1)
    return gps_fail(h, -argidx, "argument %d (%s): dtype must be float64", argidx, name);
  double* base = reinterpret_cast<double*>(reinterpret_cast<char*>(t->data) + t->byte_offset);
  if (t->ndim == 2) {
    int64_t r = t->shape[0], c = t->shape[1];
    int64_t s0 = t->strides ? t->strides[0] : c, s1 = t->strides ? t->strides[1] : 1;
    if (c > 1 && s1 != 1)
      return gps_fail(h, -argidx, "argument %d (%s): innermost stride must be 1 (got %lld)", argidx,
                      name, (long long)s1);
    if (r > 1 && s0 < c)
      return gps_fail(h, -argidx, "argument %d (%s): row stride %lld < cols %lld", argidx, name,
                      (long long)s0, (long long)c);
    if (r <= 1 && s0 < c) s0 = c;
    *out = Mat(base, r, c, s0);
    return 0;
  }
  if (t->ndim == 1 && allow_vec) {
    int64_t n = t->shape[0];
    int64_t s = t->strides ? t->strides[0] : 1;
    if (n > 1 && s != 1)
      return gps_fail(h, -argidx, "argument %d (%s): vector stride must be 1", argidx, name);
    *out = Mat(base, 1, n, n);
    return 0;
  }
  return gps_fail(h, -argidx, "argument %d (%s): ndim must be 1 or 2 (got %d)", argidx, name,
                  t->ndim);
}

int gps_as_i64(gps_handle* h, const DLTensor* t, int argidx, const char* name, int64_t n,
               const int64_t** out) {
  if (!t || !t->data) return gps_fail(h, -argidx, "argument %d (%s): null tensor", argidx, name);
  if (t->device.device_type != 2 || t->device.device_id != h->device)
    return gps_fail(h, -argidx, "argument %d (%s): must be a CUDA tensor on device %d", argidx, name,
                    h->device);
  if (t->dtype.code != 0 /*kDLInt*/ || t->dtype.bits != 64 || t->dtype.lanes != 1)
    return gps_fail(h, -argidx, "argument %d (%s): dtype must be int64", argidx, name);
  if (t->ndim != 1 || t->shape[0] != n || (t->strides && n > 1 && t->strides[0] != 1))
    return gps_fail(h, -argidx, "argument %d (%s): need a contiguous vector of %lld entries", argidx,
                    name, (long long)n);
  *out = reinterpret_cast<const int64_t*>(reinterpret_cast<char*>(t->data) + t->byte_offset);
  return 0;
}

extern "C" {

int gps_version(void) { return 100; }

int gps_create(int device, gps_handle** out) {
  if (!out) return -2;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    return -1;
  }
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
  if (prop.major != 10) {
    fprintf(stderr, "gpslim_b200: device %d is sm_%d%d; this library is built for sm_100a only\n",
            device, prop.major, prop.minor);
    return -1;
  }
  gps_handle* h = new gps_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  *out = h;
  return 0;
}

int gps_destroy(gps_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (int i = 0; i < WS_COUNT; ++i)
    if (h->ws_ptr[i]) cudaFree(h->ws_ptr[i]);
  for (auto& kv : h->splitk_ws)
    if (kv.second.first) cudaFree(kv.second.first);
  for (auto& kv : h->stream_ws)
    if (kv.second.first) cudaFree(kv.second.first);
  for (auto& e : h->events) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  delete h;
  return 0;
}

int gps_set_stream(gps_handle* h, void* s) {
  if (!h) return -1;
  h->stream = reinterpret_cast<cudaStream_t>(s);
  return 0;
}

const char* gps_last_error(gps_handle* h) { return h ? h->err.c_str() : "null handle"; }

int gps_set_option(gps_handle* h, const char* name, int64_t value) {
  if (!h || !name) return -1;
  if (!strcmp(name, "gemm_impl")) {
    h->gemm_impl = (int)value;
    return 0;
  }
  if (!strcmp(name, "gram_impl")) {
    h->gram_impl = (int)value;
    return 0;
  }
  if (!strcmp(name, "leaf_impl")) {
    h->leaf_impl = (int)value;
    return 0;
  }
  if (!strcmp(name, "trsm_leaf")) {
    if (value < GPS_NB || value > 4096 || (value & (value - 1))) return gps_fail(h, -3, "trsm_leaf must be a power of two in [128, 4096]");
    h->trsm_leaf = (int)value;
    return 0;
  }
  if (!strcmp(name, "gemm_splitk")) {
    h->gemm_splitk = (int)value;
    return 0;
  }
  if (!strcmp(name, "profile")) {
    h->profile = (int)value;
    return 0;
  }
  return gps_fail(h, -2, "unknown option '%s'", name);
}

int gps_profile_read(gps_handle* h, double* gemm_ms, double* gemm_flops, int64_t* launches,
                     int reset) {
  if (!h) return -1;
  GPS_CUDA(h, cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < h->events_used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->events[i].a, h->events[i].b) == cudaSuccess) {
      h->gemm_ms_acc += ms;
      h->gemm_flops_acc += h->events[i].flops;
    }
  }
  h->events_used = 0;
  if (gemm_ms) *gemm_ms = h->gemm_ms_acc;
  if (gemm_flops) *gemm_flops = h->gemm_flops_acc;
  if (launches) *launches = h->launches;
  if (reset) {
    h->gemm_ms_acc = 0;
    h->gemm_flops_acc = 0;
    h->launches = 0;
  }
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------ kernels
__global__ void transpose_kernel(const double* __restrict__ A, int64_t lda, double* __restrict__ B,
                                 int64_t ldb, int64_t rows, int64_t cols) {
  __shared__ double tile[32][33];
  int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int64_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = A[r * lda + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int64_t r = c0 + i, c = r0 + threadIdx.x;  // B is cols x rows
    if (r < cols && c < rows) B[r * ldb + c] = tile[threadIdx.x][i];
  }
}

int gps_transpose_launch(gps_handle* h, Mat A, Mat At) {
  if (A.rows == 0 || A.cols == 0) return 0;
  dim3 grid((unsigned)((A.cols + 31) / 32), (unsigned)((A.rows + 31) / 32));
  if (grid.y > 65535) return gps_fail(h, -103, "transpose: too many rows");
  transpose_kernel<<<grid, dim3(32, 8), 0, h->stream>>>(A.p, A.ld, At.p, At.ld, A.rows, A.cols);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

__global__ void zero_upper_kernel(double* A, int64_t ld, int64_t n, int64_t m) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t r = blockIdx.y;
  for (; r < n; r += gridDim.y)
    if (c < m && c > r) A[r * ld + c] = 0.0;
}

int gps_zero_upper_launch(gps_handle* h, Mat A) {
  if (A.rows == 0) return 0;
  dim3 grid((unsigned)((A.cols + 255) / 256), (unsigned)(A.rows < 65535 ? A.rows : 65535));
  zero_upper_kernel<<<grid, 256, 0, h->stream>>>(A.p, A.ld, A.rows, A.cols);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

__global__ void fill_kernel(double* p, int64_t n, double v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

int gps_fill_launch(gps_handle* h, double* p, int64_t n, double v) {
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  fill_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(p, n, v);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

// one warp per row; rows are contiguous -> coalesced 8-byte loads, HBM bound
__global__ void gemv_kernel(const double* __restrict__ A, int64_t lda, int64_t rows, int64_t cols,
                            const double* __restrict__ x, double* __restrict__ y, int tri) {
  int lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (; row < rows; row += nwarps) {
    int64_t k0 = 0, k1 = cols;
    if (tri == TRI_UPPER) k0 = row;
    if (tri == TRI_LOWER) k1 = (row + 1 < cols) ? row + 1 : cols;
    const double* a = A + row * lda;
    double s0 = 0, s1 = 0;
    int64_t k = k0 + lane;
    for (; k + 32 < k1; k += 64) {
      s0 = fma(a[k], x[k], s0);
      s1 = fma(a[k + 32], x[k + 32], s1);
    }
    if (k < k1) s0 = fma(a[k], x[k], s0);
    double s = warp_sum(s0 + s1);
    if (lane == 0) y[row] = s;
  }
}

int gps_gemv_launch(gps_handle* h, Mat A, const double* x, double* y, int tri) {
  if (A.rows == 0) return 0;
  int64_t blocks = (A.rows + 7) / 8;
  if (blocks > h->sm_count * 8) blocks = h->sm_count * 8;
  gemv_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(A.p, A.ld, A.rows, A.cols, x, y, tri);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

__global__ void row_sumsq_kernel(const double* __restrict__ A, int64_t lda, int64_t rows,
                                 int64_t cols, double alpha, double beta, double* __restrict__ out) {
  int lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (; row < rows; row += nwarps) {
    const double* a = A + row * lda;
    double s0 = 0, s1 = 0;
    int64_t k = lane;
    for (; k + 32 < cols; k += 64) {
      double u = a[k], v = a[k + 32];
      s0 = fma(u, u, s0);
      s1 = fma(v, v, s1);
    }
    if (k < cols) {
      double u = a[k];
      s0 = fma(u, u, s0);
    }
    double s = warp_sum(s0 + s1);
    if (lane == 0) out[row] = (beta == 0.0 ? 0.0 : beta * out[row]) + alpha * s;
  }
}

int gps_row_sumsq_launch(gps_handle* h, double alpha, Mat A, double beta, double* out) {
  if (A.rows == 0) return 0;
  int64_t blocks = (A.rows + 7) / 8;
  if (blocks > h->sm_count * 8) blocks = h->sm_count * 8;
  row_sumsq_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(A.p, A.ld, A.rows, A.cols, alpha, beta,
                                                            out);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

// deterministic single-CTA reductions (inputs are short: <= a few thousand partials, or an
// N-vector)
template <bool SQUARE>
__global__ void reduce_kernel(const double* __restrict__ x, int64_t n, double scale,
                              double* __restrict__ out, int accumulate) {
  __shared__ double sm[32];
  double s = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    double v = x[i];
    s += SQUARE ? v * v : v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.0) + scale * t;
  }
}

int gps_sum_partials(gps_handle* h, const double* parts, int64_t n, double scale, double* out,
                     int accumulate) {
  reduce_kernel<false><<<1, 1024, 0, h->stream>>>(parts, n, scale, out, accumulate);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

int gps_sumsq_launch(gps_handle* h, const double* x, int64_t n, double scale, double* out,
                     int accumulate) {
  reduce_kernel<true><<<1, 1024, 0, h->stream>>>(x, n, scale, out, accumulate);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

__global__ void log_diag_kernel(const double* __restrict__ L, int64_t ld, int64_t n,
                                double* __restrict__ out) {
  __shared__ double sm[32];
  double s = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += log(L[i * ld + i]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = t;
  }
}

extern "C" {

int gps_transpose(gps_handle* h, const DLTensor* A, DLTensor* At) {
  if (!h) return -1;
  Mat a, b;
  int rc;
  if ((rc = gps_as_mat(h, A, 2, "A", &a, false))) return rc;
  if ((rc = gps_as_mat(h, At, 3, "At_out", &b, false))) return rc;
  if (b.rows != a.cols || b.cols != a.rows) return gps_fail(h, -3, "At_out: shape mismatch");
  GPS_CUDA(h, cudaSetDevice(h->device));
  return gps_transpose_launch(h, a, b);
}

int gps_sum_log_diag(gps_handle* h, const DLTensor* L, DLTensor* out) {
  if (!h) return -1;
  Mat l, o;
  int rc;
  if ((rc = gps_as_mat(h, L, 2, "L", &l, false))) return rc;
  if ((rc = gps_as_mat(h, out, 3, "out", &o))) return rc;
  if (l.rows != l.cols) return gps_fail(h, -2, "L must be square");
  GPS_CUDA(h, cudaSetDevice(h->device));
  log_diag_kernel<<<1, 1024, 0, h->stream>>>(l.p, l.ld, l.rows, o.p);
  GPS_LAUNCH_CHECK(h);
  return 0;
}

int gps_row_sumsq(gps_handle* h, double alpha, const DLTensor* A, double beta, DLTensor* out) {
  if (!h) return -1;
  Mat a, o;
  int rc;
  if ((rc = gps_as_mat(h, A, 3, "A", &a, false))) return rc;
  if ((rc = gps_as_mat(h, out, 5, "out", &o))) return rc;
  if (o.rows * o.cols != a.rows) return gps_fail(h, -5, "out: need %lld entries", (long long)a.rows);
  GPS_CUDA(h, cudaSetDevice(h->device));
  return gps_row_sumsq_launch(h, alpha, a, beta, o.p);
}

}  // extern "C"
