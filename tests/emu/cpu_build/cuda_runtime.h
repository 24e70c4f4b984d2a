// TEST INFRASTRUCTURE ONLY -- a stand-in for <cuda_runtime.h> (found first on the include path of
// the emulation build, tests/test_library_on_cpu.py) that lets the WHOLE of libgpslim_b200 --
// host orchestration and kernel source -- compile with g++ and run on the CPU:
//   * device memory is host memory; streams do not exist: every launch runs to completion at its
//     call site, which is one legal serialisation of the stream order;
//   * a kernel launch `k<<<grid, block, smem, stream>>>(args)` is rewritten textually into
//     EMU_LAUNCH(grid, block, smem, k(args)) by the build script and executed by ../emu_fibers.h:
//     one OS thread per WARP of a block, the 32 lanes of a warp as cooperative fibers of that
//     thread, blocks one after the other, __syncthreads = warp gather + pthread barrier across
//     the warps, warp collectives = warp barriers + exchange buffer, mma.sync.m8n8k4.f64 emulated
//     per the PTX fragment layout, cp.async = copy with zero fill;
//   * the TMA / mbarrier GEMM kernel keeps its own main loop; only its six PTX helper functions
//     (mbarrier init / arrive / wait, the tensor copy with the 128-byte swizzle, ld.shared) are
//     replaced by the functional stand-ins of tma_emulation.inc, and the tensor-map encoder by a
//     plain description of the operand view.
// Nothing here is shipped or used by the product.
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "emu_fibers.h"

using std::max;
using std::min;

// ---------------------------------------------------------------- keywords
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

// ---------------------------------------------------------------- runtime API
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
struct EmuEvent { double t; };
typedef EmuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
struct cudaDeviceProp { int major, minor, multiProcessorCount; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->major = 10; p->minor = 0; p->multiProcessorCount = 4; return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = aligned_alloc(256, (n + 255) / 256 * 256);
  if (*p) memset(*p, 0xFF, (n + 255) / 256 * 256);     // NaN-poison: unwritten workspace reads show up
  return *p ? cudaSuccess : 2;
}
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset2DAsync(void* p, size_t pitch, int v, size_t w, size_t h, cudaStream_t) {
  for (size_t r = 0; r < h; ++r) memset((char*)p + r * pitch, v, w);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < h; ++r) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
  return cudaSuccess;
}
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent{0.0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

// ---------------------------------------------------------------- execution model: ../emu_fibers.h
inline dim3 gridDim, blockDim;
inline double* emu_smem = nullptr;
inline double* emu_wx = nullptr;                          // [2][nthreads] exchange slots
inline std::mutex emu_atomic_mutex;
#define __syncthreads() emu_barrier()
#define __syncwarp() emu_warp_barrier()

inline double emu_shfl(double v, int src_lane) {
  const unsigned t = emu_ltid;
  emu_wx[t] = v;
  emu_warp_barrier();
  double r = emu_wx[(t & ~31u) + ((unsigned)src_lane & 31u)];
  emu_warp_barrier();
  return r;
}
#define __shfl_sync(mask, v, src) emu_shfl((v), (src))
#define __shfl_xor_sync(mask, v, lm) emu_shfl((v), (int)((emu_ltid & 31u) ^ (unsigned)(lm)))
inline double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
inline void dmma884(double& c0, double& c1, double a, double b) {
  const unsigned t = emu_ltid, w0 = t & ~31u, lane = t & 31u;
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
inline uint32_t smem_u32(const void* p) { return (uint32_t)((const char*)p - (const char*)emu_smem); }   // byte offset
inline void __trap() { abort(); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline int atomicCAS(int* addr, int cmp, int val) {
  std::lock_guard<std::mutex> g(emu_atomic_mutex);
  int old = *addr;
  if (old == cmp) *addr = val;
  return old;
}
inline int atomicMin(int* addr, int val) {
  std::lock_guard<std::mutex> g(emu_atomic_mutex);
  int old = *addr;
  if (val < old) *addr = val;
  return old;
}
inline void cp_async16(void* smem, const void* gmem, int n) { memset(smem, 0, 16); if (n > 0) memcpy(smem, gmem, (size_t)n); }
inline void cp_async8(void* smem, const void* gmem, int n) { memset(smem, 0, 8); if (n > 0) memcpy(smem, gmem, (size_t)n); }
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
template <class T>
inline T __ldcg(const T* p) { return *p; }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }

inline int64_t emu_launch_count = 0;

template <class Body>
void emu_launch(dim3 grid, dim3 block, size_t smem_bytes, Body body) {
  ++emu_launch_count;
  gridDim = grid;
  blockDim = block;
  const unsigned nthreads = block.x * block.y * block.z;
  std::vector<double> smem(smem_bytes / sizeof(double) + 8), wx(2 * (size_t)nthreads + 64);
  emu_smem = smem.data();
  emu_wx = wx.data();
  EmuIdx g, b;
  g.x = grid.x; g.y = grid.y; g.z = grid.z;
  b.x = block.x; b.y = block.y; b.z = block.z;
  emu_run_grid(g, b, &smem, body);
}
#define EMU_LAUNCH(grid, block, smem, call) emu_launch((grid), (block), (size_t)(smem), [&]() { call; })
