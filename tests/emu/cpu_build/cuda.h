// TEST INFRASTRUCTURE ONLY -- stand-in for <cuda.h> in the CPU build of the library: the TMA
// kernel and its tensor-map encoder are cut out of gemm.cu there (see cuda_runtime.h).
#pragma once
struct CUtensorMap { char opaque[128]; };
