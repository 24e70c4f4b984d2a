// TEST INFRASTRUCTURE ONLY: a host emulation layer just wide enough to execute the SOURCE TEXT
// of the interpreter Gram kernels of gpflow-slim_b200/csrc/gram.cu on the CPU (one OS thread
// per warp of a block with its 32 lanes as cooperative fibers, blocks run one after the other,
// __syncthreads / warp shuffles emulated with barriers).  tests/test_gram_kernel_emulation_cpu.py generates a translation unit
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
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "emu_fibers.h"
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

// ---- execution model: emu_fibers.h (one OS thread per warp, its lanes as cooperative fibers)
typedef EmuIdx EmuDim3;
static EmuDim3 gridDim, blockDim;
static double* emu_smem = nullptr;          // dynamic shared memory of the running block
static double* emu_xchg = nullptr;          // one slot per thread for the shuffle emulation
static double* emu_wx = nullptr;            // [2][nthreads] exchange slots for warp collectives

// all blocks of a (gx, gy) grid of `nthreads`-thread (1-D) blocks, one after the other
template <class Body>
static void emu_run(unsigned gx, unsigned gy, unsigned nthreads, size_t smem_doubles, Body body) {
  gridDim.x = gx; gridDim.y = gy; gridDim.z = 1;
  blockDim.x = nthreads; blockDim.y = blockDim.z = 1;
  std::vector<double> smem(smem_doubles + 8), xchg(nthreads), wx(2 * (size_t)nthreads);
  emu_smem = smem.data();
  emu_xchg = xchg.data();
  emu_wx = wx.data();
  emu_run_grid(gridDim, blockDim, &smem, body);
}

// shuffles exchange inside one warp (lane masks below 32)
static inline double emu_shfl_xor(double v, int lanemask) {
  emu_xchg[threadIdx.x] = v;
  emu_warp_barrier();
  double r = emu_xchg[threadIdx.x ^ (unsigned)lanemask];
  emu_warp_barrier();
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
