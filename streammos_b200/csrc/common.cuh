// Shared helpers for the sm_100a hot-path kernels. Internal — not part of the C-ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/streammos_b200.h"

#include <atomic>
#include <utility>

// SM count of the current device (B200: 148 = 2 dies x 74), queried once per device; thread safe.
static inline int smos_sm_count() {
  static std::atomic<int> cached[64];
  int dev = 0;
  cudaGetDevice(&dev);
  const bool slot = dev >= 0 && dev < 64;
  if (slot) {
    const int v = cached[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (slot) cached[dev].store(n, std::memory_order_relaxed);
  return n;
}
#define SMOS_SM_COUNT smos_sm_count()

// One-time opt-in of `kernel` to more than 48 KB of dynamic shared memory on the current device. `done` is one atomic
// bit mask per kernel (bit = device index); concurrent host threads may both make the (idempotent) call, and devices
// with an index >= 64 simply opt in on every launch.
template <typename K>
static inline cudaError_t smos_smem_opt_in(K kernel, std::atomic<unsigned long long>& done, int bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  const bool slot = dev >= 0 && dev < 64;
  if (slot && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && slot) done.fetch_or(1ull << dev, std::memory_order_release);
  return e;
}

static inline cudaStream_t smos_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Launch-error check that does not synchronise. smos_launch_pdl records what cudaLaunchKernelEx returned (per host
// thread and translation unit: a call launches and checks inside one .cu file).
static thread_local cudaError_t smos_tls_launch_error = cudaSuccess;
static inline int smos_launch_status() {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = smos_tls_launch_error;
  smos_tls_launch_error = cudaSuccess;
  return e == cudaSuccess ? SMOS_OK : static_cast<int>(e);
}

// experiment knobs (environment overrides, read per call: host-side only)
static inline int smos_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// A scan is ~45 short dependent kernels on one stream; with plain launches each of them starts only after its
// predecessor has drained AND the launch latency has passed. Every kernel of this library is launched with the
// programmatic-stream-serialization attribute and begins with SMOS_PDL_PROLOGUE():
//   griddepcontrol.launch_dependents  lets the NEXT kernel's CTAs be scheduled as soon as all of this kernel's CTAs
//                                     have started (they fill the SMs this kernel's tail leaves idle);
//   griddepcontrol.wait               blocks until the PREVIOUS kernel has completed and its writes are visible.
// The wait comes before any global-memory access and before any early return, so a kernel never completes before
// its predecessor: completion stays transitive along the stream and read-after-write as well as write-after-read
// hazards are ordered exactly as with plain launches. What overlaps is launch latency, CTA scheduling and each
// kernel's prologue. Works under stream capture (programmatic edges in the CUDA graph). SMOS_PDL=0 (read once)
// falls back to plain launches.
__device__ __forceinline__ void smos_pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#define SMOS_PDL_PROLOGUE() smos_pdl_prologue()

static inline bool smos_pdl_enabled() {
  static const bool on = smos_env_int("SMOS_PDL", 1) != 0;
  return on;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t smos_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = smos_pdl_enabled() ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
  if (e != cudaSuccess) smos_tls_launch_error = e;
  return e;
}
// SMOS_LAUNCH((kernel<T...>), grid, block, dynamic_smem_bytes, stream, args...) — errors surface through
// smos_launch_status() (cudaGetLastError) as with <<<>>>
#define SMOS_LAUNCH(kern, grid, block, smem, st, ...) \
  ((void)smos_launch_pdl(kern, dim3(grid), dim3(block), static_cast<size_t>(smem), st, __VA_ARGS__))

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
// The same wait for barriers that other WARPS of the CTA complete a little later (pipeline skew, not a copy in flight):
// the try_wait carries a suspend-time hint, so a warp that has to wait sleeps in the barrier unit instead of spending
// issue slots on polls (ncu, PointNet stem: 683 k polls = 13 % of the kernel's instructions with the plain loop).
__device__ __forceinline__ void smos_mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  uint32_t polls = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smos_smem_u32(bar)), "r"(parity), "r"(2000u)
        : "memory");
    if (!done && ++polls > (1u << 22)) __trap();
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
  SMOS_PDL_PROLOGUE();
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
  SMOS_LAUNCH((smos_zero_kernel), static_cast<unsigned>(blocks), 256, 0, st, static_cast<uint4*>(p), n16);
  return cudaGetLastError();
}
