// TEST INFRASTRUCTURE ONLY -- stand-in for <cuda.h> in the CPU build of the library: the tensor map
// of the TMA GEMM kernel as a plain description of the row-major operand view (tma_emulation.inc reads it).
#pragma once
#include <stdint.h>
struct CUtensorMap {
  const double* p;
  uint64_t cols, rows, ld;        // dims (K, rows), row stride in elements
};
