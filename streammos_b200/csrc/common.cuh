// Shared helpers for the sm_100a hot-path kernels. Internal — not part of the C-ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/streammos_b200.h"

#define SMOS_SM_COUNT 148  // B200: 2 dies x 74 SMs

static inline cudaStream_t smos_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Launch-error check that does not synchronise.
static inline int smos_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SMOS_OK : static_cast<int>(e);
}

static inline int64_t smos_align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static inline int smos_ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

__device__ __forceinline__ unsigned smos_lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Streaming (evict-first) 128-bit store for write-once outputs larger than L2.
__device__ __forceinline__ void smos_st_cs_f4(float* p, float4 v) {
  __stcs(reinterpret_cast<float4*>(p), v);
}
