// Shared helpers for the sm_100a hot-path kernels. Internal — not part of the C-ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/streammos_b200.h"

#define SMOS_SM_COUNT 148  // B200: 2 dies x 74 SMs

static inline cudaStream_t smos_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Launch-error check that does not synchronise.
static inline int smos_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SMOS_OK : static_cast<int>(e);
}

// experiment knobs (environment overrides, read per call: host-side only)
static inline int smos_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
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
  uint32_t polls = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smos_smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && ++polls > (1u << 26)) __trap();  // a copy that never lands must fail the launch, not hang the GPU
  } while (!done);
}

// ---- pooling plan layout (built by voxel_maxpool.cu, also walked by the ordered gather) ------------
struct PoolLayout {
  int64_t hw, cells;   // H*W, B*H*W
  int64_t off_cell, off_rank, off_count, off_start, off_sorted, off_multi, bytes;
};

static inline PoolLayout smos_pool_layout(int64_t B, int64_t N, int32_t H, int32_t W) {
  PoolLayout L;
  L.hw = static_cast<int64_t>(H) * W;
  L.cells = B * L.hw;
  const int64_t bn = B * N;
  int64_t off = 0;
  L.off_cell = off;   off += smos_align_up(bn * 4, 256);
  L.off_rank = off;   off += smos_align_up(bn * 4, 256);
  L.off_count = off;  off += smos_align_up((L.cells + 4) * 4, 256);  // + cursors: points, multi-piece cells, out-of-grid
  L.off_start = off;  off += smos_align_up(L.cells * 4, 256);
  L.off_sorted = off; off += smos_align_up(bn * 8, 256);
  // cells whose segment crosses a multiple of 32 (>= 2 pieces): at most one per 32 sorted points
  L.off_multi = off;  off += smos_align_up((bn / 32 + 2) * 8, 256);
  L.bytes = off;
  return L;
}

// ---- zero fill ---------------------------------------------------------------------------------------
// Our own fill kernel instead of cudaMemsetAsync: measured on B200, a 63 MB memset node costs ~80 us inside
// the voting call (it is not a kernel ncu lists); 128-bit stores from a grid sized to the SM count run at the
// HBM write rate.
static __global__ void __launch_bounds__(256) smos_zero_kernel(uint4* __restrict__ p, int64_t n16) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n16;
       i += static_cast<int64_t>(gridDim.x) * 256)
    p[i] = z;
}

static inline cudaError_t smos_zero_async(void* p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return cudaSuccess;
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0 || (bytes & 15) != 0) return cudaMemsetAsync(p, 0, bytes, st);
  const int64_t n16 = static_cast<int64_t>(bytes >> 4);
  int64_t blocks = (n16 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(SMOS_SM_COUNT) * 16;
  if (blocks > cap) blocks = cap;
  smos_zero_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<uint4*>(p), n16);
  return cudaGetLastError();
}
