// TEST INFRASTRUCTURE ONLY: a host emulation layer just wide enough to execute the SOURCE TEXT
// of the interpreter Gram kernels of gpflow-slim_b200/csrc/gram.cu on the CPU (one std::thread
// per CUDA thread of a block, blocks run one after the other, __syncthreads / warp shuffles
// emulated with barriers).  tests/test_gram_kernel_emulation_cpu.py generates a translation unit
// = this prelude + the kernel region of gram.cu (textually, with four mechanical substitutions
// listed there) + harness_driver.inc, compiles it with g++ and compares the kernels with the
// oracle.  Nothing here is shipped or used by the product.
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "gpslim_b200.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static              /* static __shared__ arrays: one per block == one per process here */
#define __align__(n) __attribute__((aligned(n)))

struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local EmuDim3 threadIdx, blockIdx;
static EmuDim3 gridDim, blockDim;
static double* emu_smem = nullptr;          // dynamic shared memory of the running block
static double* emu_xchg = nullptr;          // one slot per thread for the shuffle emulation
static pthread_barrier_t emu_bar;

static inline void emu_barrier() { pthread_barrier_wait(&emu_bar); }
// valid because every thread of the block executes the same sequence of shuffles in these kernels
static inline double emu_shfl_xor(double v, int lanemask) {
  emu_xchg[threadIdx.x] = v;
  emu_barrier();
  double r = emu_xchg[threadIdx.x ^ (unsigned)lanemask];
  emu_barrier();
  return r;
}
static inline double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += emu_shfl_xor(v, o);
  return v;
}

// the few host-side declarations of csrc/internal.cuh the kernel region refers to
struct gps_handle { std::string err; int gram_impl = 0; int sm_count = 4; };
static int gps_fail(gps_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}
enum { W_DENSE = 0, W_GPR = 1 };
