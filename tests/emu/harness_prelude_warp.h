// TEST INFRASTRUCTURE ONLY: warp-level extension of harness_prelude.h for kernels whose warps
// diverge (the Cholesky leaf: warp 0 factors the next diagonal block while the other warps run
// DMMA updates).  Warp collectives -- __syncwarp, __shfl_sync, __shfl_xor_sync and the FP64
// tensor-core instruction mma.sync.m8n8k4 (dmma884) -- synchronise the 32 threads of ONE warp
// through the warp barrier of harness_prelude.h (fiber switches inside the warp's OS thread) and an exchange buffer.  Fragment layout of m8n8k4.f64 (PTX ISA):
//   A (8x4, row):  lane holds A[lane >> 2][lane & 3]
//   B (4x8, col):  lane holds B[lane & 3][lane >> 2]
//   C/D (8x8):     lane holds C[lane >> 2][2 (lane & 3) + {0, 1}]
#pragma once
#include "harness_prelude.h"

#include <algorithm>
#include <mutex>

using std::max;
using std::min;

static std::mutex emu_atomic_mutex;

#define __syncwarp() emu_warp_barrier()

static inline double emu_shfl(double v, int src_lane) {
  const unsigned t = threadIdx.x;
  emu_wx[t] = v;
  emu_warp_barrier();
  double r = emu_wx[(t & ~31u) + ((unsigned)src_lane & 31u)];
  emu_warp_barrier();
  return r;
}
#define __shfl_sync(mask, v, src) emu_shfl((v), (src))
static inline double emu_shfl_xor_w(double v, int lanemask) {
  return emu_shfl(v, (int)((threadIdx.x & 31u) ^ (unsigned)lanemask));
}

static inline void dmma884(double& c0, double& c1, double a, double b) {
  const unsigned t = threadIdx.x, w0 = t & ~31u, lane = t & 31u;
  emu_wx[t] = a;
  emu_wx[emu_nthreads + t] = b;
  emu_warp_barrier();
  const unsigned row = lane >> 2, col0 = 2 * (lane & 3);
  double s0 = c0, s1 = c1;
  for (unsigned k = 0; k < 4; ++k) {
    const double av = emu_wx[w0 + row * 4 + k];
    s0 = fma(av, emu_wx[emu_nthreads + w0 + col0 * 4 + k], s0);
    s1 = fma(av, emu_wx[emu_nthreads + w0 + (col0 + 1) * 4 + k], s1);
  }
  emu_warp_barrier();
  c0 = s0;
  c1 = s1;
}

static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline int atomicCAS(int* addr, int cmp, int val) {
  std::lock_guard<std::mutex> g(emu_atomic_mutex);
  int old = *addr;
  if (old == cmp) *addr = val;
  return old;
}
static inline int atomicMin(int* addr, int val) {
  std::lock_guard<std::mutex> g(emu_atomic_mutex);
  int old = *addr;
  if (val < old) *addr = val;
  return old;
}
// cp.async: the copy completes before the (emulated) wait; bytes beyond src_bytes are zero-filled
static inline void cp_async16(void* smem, const void* gmem, int src_bytes) {
  memset(smem, 0, 16);
  if (src_bytes > 0) memcpy(smem, gmem, (size_t)src_bytes);
}
static inline void cp_async8(void* smem, const void* gmem, int src_bytes) {
  memset(smem, 0, 8);
  if (src_bytes > 0) memcpy(smem, gmem, (size_t)src_bytes);
}
static inline void cp_async_commit() {}
template <int N>
static inline void cp_async_wait() {}

#define GPS_NB 128
