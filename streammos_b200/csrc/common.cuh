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

// ---- mbarrier + TMA 1-D bulk copy (cp.async.bulk, SASS: UBLKCP) ---------------------------------
__device__ __forceinline__ uint32_t smos_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void smos_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smos_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void smos_fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void smos_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void smos_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smos_smem_u32(bar)), "r"(bytes)
               : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void smos_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smos_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smos_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void smos_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smos_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
