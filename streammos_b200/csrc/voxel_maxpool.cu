// VoxelMaxPool for sm_100a: sort-by-cell, balanced segmented max, dense one-pass writer.
//
// The reference (deep_point/src/point_deep_cuda_kernel.cu:24-99) zero-fills the dense grid,
// stores every point feature once (racy init), then issues one CAS-loop atomic per
// (point, channel) into global memory. LiDAR density is extremely skewed (hundreds of points
// in one BEV cell next to the sensor, >90 % of the cells empty), so neither per-cell atomics
// nor per-tile ownership balance. Here:
//
//   plan  (4 small kernels, coordinates only, shared by forward, backward and cell-order gathers; up to 8
//         plans per launch) zero the counters -> cell index + per-cell ranks (one atomic per run of equal
//         cells) -> segment allocation (one cursor atomic per CTA) -> scatter: `sorted` lists the valid points
//         grouped by cell, out-of-grid points at its tail; count[cell] / start[cell] locate each group.
//   permute (channel-major input only) 64-point x C tiles through shared memory: one row of C run-maxima per
//         run of equal cells, written at the run's sorted position (LDG tiles by default, tiled-TMA pipeline
//         behind SMOS_PERM_LDG=0).
//   reduce one warp per 32 consecutive sorted positions, whatever cells they fall in: perfectly balanced.
//         Inside its window a warp reduces each run of equal cells (a "piece") and writes ONE row of C maxima
//         per piece into the row buffer, at the position of the piece's first point. No atomics.
//   combine one warp per cell whose segment crosses a multiple of 32: folds its pieces into its first row.
//   write one thread per 4 adjacent output cells x all channel groups: empty cells store zeros, occupied cells
//         read their one row. Every output element is written exactly once with 128-bit stores: the 201 MB
//         zero-fill of the reference is fused away.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "gather_taps.cuh"

namespace {

constexpr int kPlanThreads = 256;
constexpr int kReduceWarps = 4;   // warps per CTA in phase A
constexpr int kWriteThreads = 256;
constexpr int kCG = 8;            // channels per thread in phase B
// A point that continues a run of equal cells begun by the lane before it (same 16-point segment, hence
// consecutive sorted positions) is "merged": on the channel-major path only the run's first point gets a
// row (the run maximum), which removes ~60 % of the permute's writes and of the reduction's reads in scan
// order. Bit 30 of a plan position / bit 31 of sorted.x carry the mark.
constexpr int32_t kMergedPos = 0x40000000;
constexpr int32_t kPosMask = 0x3fffffff;
constexpr uint32_t kMergedN = 0x80000000u;

// ---- plan kernels ------------------------------------------------------------------------------
// Up to kMaxPlans plans are built by ONE launch of each of the three kernels (a scan needs five: BEV at
// three scales, range view at two), so the per-scan plan cost is 3 launches instead of 15.
constexpr int kMaxPlans = 8;

struct PlanDev {
  const float* ind;
  int64_t ind_sb, ind_sn, ind_sd;
  int64_t* vmi;
  int64_t vmi_stride;
  int32_t* cell;
  int32_t* rank;
  int32_t* count;   // [cells] + cursor[0] (points) + cursor[1] (multi-piece cells)
  int32_t* start;
  int2* sorted;
  int2* multi;
  TapsS* taps;       // optional: BilinearSample sampling records, one per slot of `sorted`
  int64_t pt_begin;    // first global point index of this plan
  int64_t quad_begin;  // first global cell-quad index (warp aligned)
  int32_t N, H, W, hw, cells, bn;
  float sh, sw;
  const float* scale_dev;  // optional: {scale_h, scale_w} on the device (overrides sh / sw)
};

struct PlanBatch {
  PlanDev p[kMaxPlans];
  int64_t pt_total, quad_total;
  int32_t n;
};

// ---- plan 0: clear count[cells] and the four cursors of every plan (thread = 4 cells) -----------
__global__ void __launch_bounds__(kPlanThreads)
pool_zero_counts_kernel(const __grid_constant__ PlanBatch pb) {
  SMOS_PDL_PROLOGUE();
  const int64_t gq = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gq >= pb.quad_total) return;
  int j = 0;
  while (j + 1 < pb.n && gq >= pb.p[j + 1].quad_begin) ++j;
  const PlanDev& P = pb.p[j];
  const int32_t c0 = static_cast<int32_t>(gq - P.quad_begin) * 4;
  if (c0 == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) P.count[P.cells + q] = 0;
  }
  if (c0 + 3 < P.cells && (reinterpret_cast<uintptr_t>(P.count) & 15) == 0) {
    *reinterpret_cast<int4*>(P.count + c0) = make_int4(0, 0, 0, 0);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (c0 + q < P.cells) P.count[c0 + q] = 0;
  }
}

// ---- plan 1: cell index + warp-aggregated per-cell histogram -------------------------------
__global__ void __launch_bounds__(kPlanThreads)
pool_cell_index_kernel(const __grid_constant__ PlanBatch pb) {
  SMOS_PDL_PROLOGUE();
  const int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int32_t* target = nullptr;  // &count[gcell] of my plan; the out-of-grid cursor for invalid points; null past the end
  int32_t* rank_out = nullptr;
  bool tile_head = false, valid = false;
  if (gi < pb.pt_total) {
    int j = 0;
    while (j + 1 < pb.n && gi >= pb.p[j + 1].pt_begin) ++j;
    const PlanDev& P = pb.p[j];
    const int64_t i = gi - P.pt_begin;
    // (B * N < 2^31 is checked at the entry point: a 32-bit quotient, and none at all for the first batch — the 64-bit
    // division was 14 % of this kernel's instructions and sat in front of every thread's coordinate load)
    const uint32_t iu = static_cast<uint32_t>(i);
    const int32_t b = iu < static_cast<uint32_t>(P.N) ? 0 : static_cast<int32_t>(iu / static_cast<uint32_t>(P.N));
    const int32_t n = static_cast<int32_t>(i - static_cast<int64_t>(b) * P.N);
    const float* q = P.ind + b * P.ind_sb + n * P.ind_sn;
    // fp32 multiply then C-cast truncation toward zero (reference .cu:40)
    const float sh = P.scale_dev ? __ldg(P.scale_dev) : P.sh, sw = P.scale_dev ? __ldg(P.scale_dev + 1) : P.sw;
    const float fh = __fmul_rn(q[0], sh);
    const float fw = __fmul_rn(q[P.ind_sd], sw);
    const long long ih = static_cast<long long>(fh);
    const long long iw = static_cast<long long>(fw);
    int32_t cell = -1;
    target = P.count + P.cells + 2;  // cursor[2]: out-of-grid points, listed at the tail of `sorted`
    if (ih >= 0 && ih < P.H && iw >= 0 && iw < P.W) {
      cell = static_cast<int32_t>(ih) * P.W + static_cast<int32_t>(iw);
      target = P.count + (b * P.hw + cell);
      valid = true;
    }
    P.cell[i] = cell;
    rank_out = P.rank + i;
    tile_head = (n & 15) == 0;  // runs never cross the 16-point segments the permute kernels sweep
    if (P.vmi != nullptr) P.vmi[i] = cell >= 0 ? static_cast<int64_t>(b) * P.vmi_stride + cell : -1;
  }
  // one atomic per RUN of equal cells among adjacent lanes (in scan order neighbouring points share
  // cells): a shuffle + ballot finds the runs, the first lane of a run claims ranks for all of it.
  // (Equal cells in non-adjacent lanes simply form separate runs; __match_any_sync would merge them
  // but costs far more issue slots than it saves atomics.)
  const unsigned long long key = reinterpret_cast<unsigned long long>(target);
  const int lane = threadIdx.x & 31;
  const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev || tile_head);
  const int leader = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));       // last head at or before me
  const unsigned after = lane == 31 ? 0u : (heads & (0xfffffffeu << lane));  // heads strictly after me
  const int run_end = after ? __ffs(after) - 1 : 32;                        // first lane of the next run
  int32_t base = 0;
  if (lane == leader && target != nullptr) base = atomicAdd(target, run_end - leader);
  base = __shfl_sync(0xffffffffu, base, leader);
  // valid: rank inside the cell (| kMergedPos for run followers); invalid: -2 - rank among the out-of-grid points
  if (rank_out != nullptr)
    *rank_out = valid ? (base + (lane - leader)) | (lane != leader ? kMergedPos : 0) : -2 - (base + (lane - leader));
}

// ---- plan 2: give every occupied cell a segment of the sorted list ---------------------------
// Order between warps is irrelevant (max is order independent), so a warp scan plus one atomic on a
// global cursor replaces a device-wide prefix scan. thread = 4 cells; a warp never straddles two plans.
__global__ void __launch_bounds__(kPlanThreads)
pool_cell_alloc_kernel(const __grid_constant__ PlanBatch pb) {
  SMOS_PDL_PROLOGUE();
  __shared__ int32_t s_tot[kPlanThreads / 32], s_plan[kPlanThreads / 32], s_base;
  const int64_t gq = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = gq < pb.quad_total;  // whole warps only (quad ranges are warp aligned)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int j = 0;
  while (j + 1 < pb.n && gq >= pb.p[j + 1].quad_begin) ++j;
  const PlanDev& P = pb.p[j];
  const int32_t c0 = active ? static_cast<int32_t>(gq - P.quad_begin) * 4 : 0;
  int32_t c[4] = {0, 0, 0, 0};
  const bool vec = active && (c0 + 3 < P.cells) && ((reinterpret_cast<uintptr_t>(P.count) & 15) == 0);
  if (vec) {
    const int4 v = *reinterpret_cast<const int4*>(P.count + c0);
    c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
  } else if (active) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (c0 + q < P.cells) c[q] = P.count[c0 + q];
  }
  const int32_t mine = c[0] + c[1] + c[2] + c[3];
  int32_t x = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  const int32_t warp_total = __shfl_sync(0xffffffffu, x, 31);
  int32_t* cursor = P.count + P.cells;
  // one atomic per CTA on the plan's cursor (thousands of same-address returning atomics, one per warp, were the
  // kernel's critical path); a CTA whose warps belong to different plans falls back to one atomic per warp
  if (lane == 31) { s_tot[wid] = warp_total; s_plan[wid] = active ? j : -1; }
  __syncthreads();
  bool uniform = true;
  int32_t before = 0, cta_total = 0;
#pragma unroll
  for (int w = 0; w < kPlanThreads / 32; ++w) {
    const int32_t pw = s_plan[w];
    if (pw >= 0 && pw != s_plan[0]) uniform = false;
    if (w < wid) before += s_tot[w];
    cta_total += s_tot[w];
  }
  if (s_plan[0] < 0) uniform = false;
  int32_t base = 0;
  if (uniform) {
    if (threadIdx.x == 0) s_base = cta_total > 0 ? atomicAdd(cursor, cta_total) : 0;
    __syncthreads();
    base = s_base + before;
  } else {
    if (lane == 31 && warp_total > 0) base = atomicAdd(cursor, warp_total);
    base = __shfl_sync(0xffffffffu, base, 31);
  }
  if (!active) return;
  int32_t s[4];
  s[0] = base + x - mine;
  s[1] = s[0] + c[0];
  s[2] = s[1] + c[1];
  s[3] = s[2] + c[2];
  if (vec && ((reinterpret_cast<uintptr_t>(P.start) & 15) == 0)) {
    *reinterpret_cast<int4*>(P.start + c0) = make_int4(s[0], s[1], s[2], s[3]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (c0 + q < P.cells) P.start[c0 + q] = s[q];
  }
  // a segment [s, s+c) that crosses a multiple of 32 is reduced as several pieces by phase A; list it
  // so the combine kernel can fold them into row s (cursor[1] counts the list)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const bool is_multi = c[q] > 0 && ((s[q] & 31) + c[q] > 32);
    const unsigned mm = __ballot_sync(0xffffffffu, is_multi);
    if (mm) {
      int32_t mbase = 0;
      if (lane == 0) mbase = atomicAdd(cursor + 1, __popc(mm));
      mbase = __shfl_sync(0xffffffffu, mbase, 0);
      if (is_multi) P.multi[mbase + __popc(mm & smos_lanemask_lt())] = make_int2(s[q], c[q]);
    }
  }
}

// ---- plan 3: place every valid point in its cell's segment -----------------------------------
__global__ void __launch_bounds__(kPlanThreads)
pool_cell_scatter_kernel(const __grid_constant__ PlanBatch pb) {
  SMOS_PDL_PROLOGUE();
  const int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gi >= pb.pt_total) return;
  int j = 0;
  while (j + 1 < pb.n && gi >= pb.p[j + 1].pt_begin) ++j;
  const PlanDev& P = pb.p[j];
  const int64_t i = gi - P.pt_begin;
  const int32_t cell = P.cell[i];
  const uint32_t iu = static_cast<uint32_t>(i);  // 32-bit quotient (30 % of this kernel's instructions as a 64-bit one)
  const int32_t b = iu < static_cast<uint32_t>(P.N) ? 0 : static_cast<int32_t>(iu / static_cast<uint32_t>(P.N));
  const int32_t n = static_cast<int32_t>(i - static_cast<int64_t>(b) * P.N);
  const int32_t r = P.rank[i];
  if (cell < 0) {
    // out-of-grid points fill `sorted` from the end (their .y = -1 - b); pooling never reads them (it stops at
    // cursor[0]) but a gather that walks the list in cell order must still produce their (zero-padded) samples
    const int32_t slot = P.bn - 1 - (-2 - r);
    P.sorted[slot] = make_int2(n, -1 - b);
    if (P.taps != nullptr) {
      const float* q = P.ind + b * P.ind_sb + n * P.ind_sn;
      P.taps[slot] = make_taps_record(__ldg(q), __ldg(q + P.ind_sd), P.sh, P.sw, P.H, P.W, n, b);
    }
    return;
  }
  const int32_t gcell = b * P.hw + cell;
  const int32_t merged = r & kMergedPos;
  const int32_t pos = __ldg(P.start + gcell) + (r & kPosMask);
  P.rank[i] = pos | merged;  // from here on: sorted position (| kMergedPos), negative if invalid
  P.sorted[pos] = make_int2(static_cast<int32_t>(static_cast<uint32_t>(n) | (merged ? kMergedN : 0u)), gcell);
  if (P.taps != nullptr) {
    const float* q = P.ind + b * P.ind_sb + n * P.ind_sn;
    P.taps[pos] = make_taps_record(__ldg(q), __ldg(q + P.ind_sd), P.sh, P.sw, P.H, P.W, n, b);
  }
}

// ---- phase A0 (channel-major input only): permute into sorted point-major rows -----------------
// A (B,C,N,1)-contiguous tensor (the PointNet output) keeps a point's channels N*4 bytes apart, so
// gathering it in sorted order costs one 32-byte sector per 4 useful bytes. Instead read it in its own
// order (fully coalesced), transpose 64 points x C channels through shared memory and write each
// point's C-float row to its sorted position: every sector moved is fully used, and the reduction
// below then streams rows in position order.
constexpr int kPermPts = 64;
constexpr int kPermThreads = 256;

__global__ void __launch_bounds__(kPermThreads)
pool_permute_kernel(const float* __restrict__ feat, int32_t C, int32_t N, int64_t f_sb, int64_t f_sc, int64_t f_sn,
                    const int32_t* __restrict__ pos, float* __restrict__ rows) {
  SMOS_PDL_PROLOGUE();
  extern __shared__ float tile[];  // [kPermPts][C + 1]
  __shared__ int32_t s_pos[kPermPts];
  const int32_t b = blockIdx.y;
  const int32_t n0 = blockIdx.x * kPermPts;
  const int32_t np = min(kPermPts, N - n0);
  const int32_t ld = C + 1;
  if (threadIdx.x < kPermPts) s_pos[threadIdx.x] = threadIdx.x < np ? pos[static_cast<int64_t>(b) * N + n0 + threadIdx.x] : -1;
  const float* fb = feat + b * f_sb + static_cast<int64_t>(n0) * f_sn;
  // loads: consecutive threads -> consecutive points of one channel (coalesced when f_sn == 1)
  // (eight unconditional loads per thread in flight; out-of-range slots re-read a clamped address)
  const int32_t total = C * kPermPts;
  for (int32_t i0 = threadIdx.x; i0 < total; i0 += kPermThreads * 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int32_t i = min(i0 + u * kPermThreads, total - 1);
      const int32_t c = i / kPermPts, p = min(i - c * kPermPts, np - 1);
      v[u] = __ldg(fb + static_cast<int64_t>(c) * f_sc + static_cast<int64_t>(p) * f_sn);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int32_t i = i0 + u * kPermThreads;
      if (i < total) {
        const int32_t c = i / kPermPts, p = i - c * kPermPts;
        tile[p * ld + c] = v[u];
      }
    }
  }
  __syncthreads();
  // stores: one warp per point row, lanes over channels (C*4 contiguous bytes per row)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int32_t p = wid; p < np; p += kPermThreads / 32) {
    const int32_t q = s_pos[p];
    if (q < 0 || (q & kMergedPos)) continue;  // invalid point, or merged into the row of its run's first point
    int32_t len = 1;
    while (p + len < np && s_pos[p + len] >= 0 && (s_pos[p + len] & kMergedPos)) ++len;
    float* dst = rows + static_cast<int64_t>(q) * C;
    for (int32_t c = lane; c < C; c += 32) {
      float v = tile[p * ld + c];
      for (int32_t t = 1; t < len; ++t) v = fmaxf(v, tile[(p + t) * ld + c]);
      dst[c] = v;
    }
  }
}

// Sweep of one 16-point segment of a staged tile for CPL channels per lane (lane, lane + 32, ...): runs never
// cross a multiple of 16 points (pool_cell_index_kernel breaks them there), so the segment is self-contained and
// its run structure (hm: heads, mm: merged followers, 16 bits each) is the same for every channel: the code below
// is fully unrolled with warp-uniform predicates. ld(j, k) returns the four values of points 4k..4k+3 of the
// lane's j-th channel; q_of(p) the sorted position of point p of the segment.
template <int CPL, typename Ld, typename Qof>
__device__ __forceinline__ void perm_sweep16(unsigned hm, unsigned mm, int32_t C, int32_t c0, bool (&ok)[CPL], Ld ld,
                                             Qof q_of, float* __restrict__ rows) {
  float v[CPL][16];
#pragma unroll
  for (int j = 0; j < CPL; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 t = ld(j, k);
      v[j][4 * k] = t.x; v[j][4 * k + 1] = t.y; v[j][4 * k + 2] = t.z; v[j][4 * k + 3] = t.w;
    }
  float acc[CPL];
  int32_t q = -1;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    if ((mm >> p) & 1u) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] = fmaxf(acc[j], v[j][p]);
    } else {
      if (q >= 0) {
#pragma unroll
        for (int j = 0; j < CPL; ++j)
          if (ok[j]) rows[static_cast<int64_t>(q) * C + c0 + 32 * j] = acc[j];
      }
      q = -1;
      if ((hm >> p) & 1u) {
        q = q_of(p);
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc[j] = v[j][p];
      }
    }
  }
  if (q >= 0) {
#pragma unroll
    for (int j = 0; j < CPL; ++j)
      if (ok[j]) rows[static_cast<int64_t>(q) * C + c0 + 32 * j] = acc[j];
  }
}

// LDG variant of the permute for the same layouts as the TMA kernel: one 64-point x C tile per CTA, every thread
// issues all of its 128-bit loads (4 points of one channel each) back to back, parks them in shared memory
// ([channel][68]: conflict-free 128-bit column reads) and sweeps 16-point segments. Many small CTAs per SM instead
// of a software pipeline: the LSU path keeps far more requests in flight per SM than the TMA unit.
template <int CPL>
__global__ void __launch_bounds__(kPermThreads)
pool_permute_ldg_kernel(const float* __restrict__ feat, int32_t C, int32_t N, int64_t f_sb, int64_t f_sc,
                        const int32_t* __restrict__ pos, float* __restrict__ rows) {
  SMOS_PDL_PROLOGUE();
  constexpr int PTS = kPermPts;
  constexpr int kPitch = PTS + 4;
  extern __shared__ __align__(16) float ltile[];  // [C][kPitch]
  __shared__ int32_t s_pos[PTS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int32_t b = blockIdx.y;
  const int32_t n0 = blockIdx.x * PTS;
  const int32_t np = min(PTS, N - n0);  // multiple of 4
  if (threadIdx.x < PTS) s_pos[threadIdx.x] = threadIdx.x < np ? __ldg(pos + static_cast<int64_t>(b) * N + n0 + threadIdx.x) : -1;
  const float* fb = feat + b * f_sb + n0;
  const int32_t nvec = C * (PTS / 4);
  for (int32_t i0 = threadIdx.x; i0 < nvec; i0 += kPermThreads * 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int32_t i = min(i0 + u * kPermThreads, nvec - 1);
      const int32_t c = i >> 4, p4 = min((i & 15) << 2, np - 4);  // unconditional loads on clamped addresses
      v[u] = __ldg(reinterpret_cast<const float4*>(fb + static_cast<int64_t>(c) * f_sc + p4));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int32_t i = i0 + u * kPermThreads;
      if (i < nvec) *reinterpret_cast<float4*>(ltile + (i >> 4) * kPitch + ((i & 15) << 2)) = v[u];
    }
  }
  __syncthreads();
  const int32_t q0 = s_pos[lane], q1 = s_pos[32 + lane];
  const unsigned long long hmask =
      static_cast<unsigned long long>(__ballot_sync(0xffffffffu, q0 >= 0 && !(q0 & kMergedPos))) |
      (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, q1 >= 0 && !(q1 & kMergedPos))) << 32);
  const unsigned long long mmask =
      static_cast<unsigned long long>(__ballot_sync(0xffffffffu, q0 >= 0 && (q0 & kMergedPos))) |
      (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, q1 >= 0 && (q1 & kMergedPos))) << 32);
  const int32_t ncg = (C + 32 * CPL - 1) / (32 * CPL);  // channel groups of 32 * CPL
  constexpr int32_t nseg = PTS / 16;
  for (int32_t item = wid; item < ncg * nseg; item += kPermThreads / 32) {
    const int32_t cg = item / nseg, seg = item - cg * nseg;
    const int32_t c0 = cg * 32 * CPL + lane;
    bool ok[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) ok[j] = c0 + 32 * j < C;
    const unsigned hm = static_cast<unsigned>(hmask >> (seg * 16)) & 0xffffu;
    const unsigned mm = static_cast<unsigned>(mmask >> (seg * 16)) & 0xffffu;
    if (hm == 0u) continue;  // no row owner in this segment
    perm_sweep16<CPL>(
        hm, mm, C, c0, ok,
        [&](int j, int k) { return *reinterpret_cast<const float4*>(ltile + (ok[j] ? c0 + 32 * j : 0) * kPitch + seg * 16 + 4 * k); },
        [&](int p) { return s_pos[seg * 16 + p]; }, rows);
  }
}

// TMA variant of the permute (used when the channel rows are contiguous and 16-byte aligned, i.e. the
// (B,C,N,1)-contiguous tensor the reference produces, N % 4 == 0, C % 32 == 0). Persistent CTAs walk tiles of
// 32 channels x 128 points. A tile arrives with four tiled-TMA operations (cp.async.bulk.tensor over a rank-3
// tensor map {N, C, B}, box {32 points, 32 channels, 1}, 128-byte swizzle, SASS UTMALDG) plus one linear bulk
// copy for its sorted positions; STAGES tiles are in flight per CTA and one __syncthreads per tile releases a
// stage. (Earlier versions: one 256-byte cp.async.bulk per channel row of a 64-point x C tile, a proxy fence
// (MEMBAR.ALL.CTA) and two barriers per tile — 43 us; the sweep below waited on every one of those latencies.)
constexpr int kTmBox = 32;  // points per TMA box: 32 floats = one 128-byte swizzle span

__device__ __forceinline__ void smos_tma_load_3d(void* dst_smem, const CUtensorMap* tm, int32_t x, int32_t y, int32_t z,
                                                 uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smos_smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(z), "r"(smos_smem_u32(bar))
      : "memory");
}

constexpr int kTmCh = 32;    // channels per tile (one lane per channel)
constexpr int kTmPts = 128;  // points per tile: 512 contiguous bytes per channel row and DRAM page visit

template <int STAGES>
__global__ void __launch_bounds__(kPermThreads)
pool_permute_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ feat, int32_t C, int32_t N,
                        int32_t B, int64_t f_sb, int64_t f_sc, const int32_t* __restrict__ pos,
                        float* __restrict__ rows) {
  SMOS_PDL_PROLOGUE();
  constexpr int PTS = kTmPts;
  constexpr int kBoxes = PTS / kTmBox;            // TMA boxes per tile (32 points x 32 channels, 4 KB each)
  constexpr int kBoxFloats = kTmCh * kTmBox;      // 1024 floats: boxes stay 1024-byte aligned
  constexpr int kSeg = PTS / (kPermThreads / 32); // points per warp segment (16)
  static_assert(kSeg % 4 == 0 && PTS == 128, "masks are kept as two 64-bit words");
  extern __shared__ float ptile_raw[];  // [STAGES][kBoxes][32 channels][32 points], 128-byte swizzle
  __shared__ __align__(8) uint64_t bar[STAGES];
  __shared__ __align__(16) int32_t s_pos[STAGES][PTS];
  float* ptile = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ptile_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int32_t tiles_per_row = (N + PTS - 1) / PTS;
  const int32_t ncg = C / kTmCh;                       // channel groups (C % 32 == 0 on this path)
  const int32_t ntiles = tiles_per_row * ncg * B;      // tile id = (b * ncg + cg) * tiles_per_row + point tile
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) smos_mbar_init(&bar[i], 1);
    smos_fence_mbar_init();
  }
  __syncthreads();

  // one thread arms the stage's mbarrier and issues every copy of a tile: the tile's sorted positions (one linear
  // bulk copy) and, for whole tiles, its features (four tiled-TMA boxes). Partial tiles at the end of a scan do
  // not rely on out-of-bounds fill: their features are read straight from global memory by the sweep.
  auto issue = [&](int32_t t, int buf) {
    const int32_t bc = t / tiles_per_row;
    const int32_t b = bc / ncg, cg = bc - b * ncg;
    const int32_t n0 = (t - bc * tiles_per_row) * PTS;
    const int32_t np = min(PTS, N - n0);  // multiple of 4
    const bool whole = np == PTS;
    smos_mbar_expect_tx(&bar[buf], static_cast<uint32_t>(np) * 4u + (whole ? static_cast<uint32_t>(PTS) * kTmCh * 4u : 0u));
    smos_bulk_g2s(&s_pos[buf][0], pos + static_cast<int64_t>(b) * N + n0, static_cast<uint32_t>(np) * 4u, &bar[buf]);
    if (whole) {
#pragma unroll
      for (int k = 0; k < kBoxes; ++k)
        smos_tma_load_3d(ptile + (buf * kBoxes + k) * kBoxFloats, &tmap, n0 + k * kTmBox, cg * kTmCh, b, &bar[buf]);
    }
  };

  const int32_t stride = gridDim.x;
  int32_t t = blockIdx.x;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i)
      if (t + i * stride < ntiles) issue(t + i * stride, i);
  }
  uint32_t phases = 0u;  // bit i: parity to wait for on bar[i]
  int buf = 0;
  for (; t < ntiles; t += stride) {
    const int32_t bc = t / tiles_per_row;
    const int32_t b = bc / ncg, cg = bc - b * ncg;
    const int32_t n0 = (t - bc * tiles_per_row) * PTS;
    const int32_t np = min(PTS, N - n0);
    const bool whole = np == PTS;
    smos_mbar_wait(&bar[buf], (phases >> buf) & 1u);  // every thread observes the completion: the bytes are visible
    phases ^= 1u << buf;
    const float* tile = ptile + buf * kBoxes * kBoxFloats;
    // run structure of the tile, the same for every channel: heads (points that own a row) and merged followers
    unsigned long long hmask[2], mmask[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int32_t pa = h * 64 + lane, pb = pa + 32;
      const int32_t qa = pa < np ? s_pos[buf][pa] : -1, qb = pb < np ? s_pos[buf][pb] : -1;
      hmask[h] = static_cast<unsigned long long>(__ballot_sync(0xffffffffu, qa >= 0 && !(qa & kMergedPos))) |
                 (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, qb >= 0 && !(qb & kMergedPos))) << 32);
      mmask[h] = static_cast<unsigned long long>(__ballot_sync(0xffffffffu, qa >= 0 && (qa & kMergedPos))) |
                 (static_cast<unsigned long long>(__ballot_sync(0xffffffffu, qb >= 0 && (qb & kMergedPos))) << 32);
    }
    // warp = 16-point segment of the tile, lane = channel; all control flow below is warp uniform. A run that starts
    // in my segment is followed to its end (runs never cross a multiple of 64 points); leading merged points belong
    // to the previous segment's sweep. Four points per 128-bit shared-memory load: with the 128-byte swizzle the
    // 16-byte chunk j of channel row c sits at chunk j ^ (c & 7) — eight consecutive lanes, eight bank groups.
    {
      const int32_t c = cg * kTmCh + lane;
      const float* gsrc = feat + b * f_sb + static_cast<int64_t>(c) * f_sc + n0;  // partial tiles only
      const int32_t p_end = (wid + 1) * kSeg;
      float acc = 0.f;
      int32_t q = -1;  // sorted position of the open run's row, -1: none
      bool done = false;
      for (int32_t g = wid * kSeg; g < PTS && !done; g += 4) {
        const unsigned hm = static_cast<unsigned>((g & 64 ? hmask[1] : hmask[0]) >> (g & 63)) & 0xfu;
        const unsigned mm = static_cast<unsigned>((g & 64 ? mmask[1] : mmask[0]) >> (g & 63)) & 0xfu;
        if (g >= p_end && !(q >= 0 && (mm & 1u))) break;  // beyond my segment and no run of mine continues
        float4 v4;
        if (whole) {
          v4 = *reinterpret_cast<const float4*>(tile + (g >> 5) * kBoxFloats + lane * kTmBox + ((((g & 31) >> 2) ^ (lane & 7)) << 2));
        } else {  // N % 4 == 0: a group of four points is inside the scan or entirely past its end
          v4 = g < np ? __ldg(reinterpret_cast<const float4*>(gsrc + g)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if ((mm >> k) & 1u) {
            if (q >= 0) acc = fmaxf(acc, v[k]);
          } else {  // a head or an invalid point closes the open run
            if (q >= 0) rows[static_cast<int64_t>(q) * C + c] = acc;
            q = -1;
            if (g + k >= p_end) { done = true; break; }
            if ((hm >> k) & 1u) { q = s_pos[buf][g + k]; acc = v[k]; }
          }
        }
      }
      if (q >= 0) rows[static_cast<int64_t>(q) * C + c] = acc;
    }
    // everyone is done reading `buf` (generic-proxy reads ordered by the barrier); refill it right away with the
    // tile STAGES strides ahead, so that STAGES - 1 tiles are in flight during the next sweep
    __syncthreads();
    if (threadIdx.x == 0 && t + STAGES * stride < ntiles) issue(t + STAGES * stride, buf);
    buf = buf + 1 == STAGES ? 0 : buf + 1;
  }
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query (no link-time dependency on libcuda)
typedef CUresult (*smos_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static smos_encode_tiled_fn encode_tiled() {
  // function-local static: initialised once, thread safe (C++11)
  static const smos_encode_tiled_fn fn = []() -> smos_encode_tiled_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<smos_encode_tiled_fn>(p);
    return nullptr;
  }();
  return fn;
}

template <int STAGES>
static int launch_permute_tma(const float* feat, int32_t C, int64_t N, int64_t B, int64_t f_sb, int64_t f_sc,
                              const int32_t* pos, float* rows, int device, cudaStream_t st) {
  smos_encode_tiled_fn enc = encode_tiled();
  if (enc == nullptr) return SMOS_EUNSUPPORTED;
  CUtensorMap tm;
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(B)};
  const cuuint64_t gstr[2] = {static_cast<cuuint64_t>(f_sc) * 4u, static_cast<cuuint64_t>(f_sb) * 4u};
  const cuuint32_t box[3] = {kTmBox, kTmCh, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(feat), gdim, gstr, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return SMOS_EUNSUPPORTED;
  const size_t smem = static_cast<size_t>(STAGES) * kTmPts * kTmCh * 4 + 1024;
  static std::atomic<unsigned long long> opted{0};
  (void)device;
  if (cudaError_t e = smos_smem_opt_in(pool_permute_tma_kernel<STAGES>, opted, 200 * 1024); e != cudaSuccess)
    return static_cast<int>(e);
  const int64_t ntiles = static_cast<int64_t>(smos_ceil_div(N, kTmPts)) * (C / kTmCh) * B;
  int ctas_per_sm = static_cast<int>((224 * 1024) / (smem + 2048));
  if (ctas_per_sm > 6) ctas_per_sm = 6;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const int want = env_int("SMOS_PERM_CTAS", 0);
  if (want >= 1 && want < ctas_per_sm) ctas_per_sm = want;
  int64_t pgrid = static_cast<int64_t>(SMOS_SM_COUNT) * ctas_per_sm;
  if (pgrid > ntiles) pgrid = ntiles;
  SMOS_LAUNCH((pool_permute_tma_kernel<STAGES>), static_cast<unsigned>(pgrid), kPermThreads, smem, st, 
      tm, feat, C, static_cast<int32_t>(N), static_cast<int32_t>(B), f_sb, f_sc, pos, rows);
  return SMOS_OK;
}

// pipeline depth of the TMA permute; SMOS_PERM_STAGES (2 / 3 / 4) overrides the default for experiments
static int perm_stages() {
  const int x = env_int("SMOS_PERM_STAGES", 0);
  return (x >= 2 && x <= 4) ? x : 3;
}

// ---- phase A: piece maxima ------------------------------------------------------------------
// Warp w owns sorted positions [32w, 32w+32). A piece = maximal run of equal cells inside that
// range; its C maxima go to rows[first position of the piece][0..C).
// Lanes run over channels (VEC consecutive channels per lane, 32*VEC per pass); the 32 point rows of
// the warp are loaded kBatch at a time so that enough independent loads are in flight to cover the
// HBM latency (one wave of warps covers the whole input).
// SORTED_ROWS: the point rows already sit at their sorted positions inside `rows` (after the permute
//   kernel) and are reduced in place; otherwise they are read from `feat` (point-major, f_sc == 1)
//   through the sorted list.
template <int VEC> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int VEC> __device__ __forceinline__ typename VecT<VEC>::type vmax(typename VecT<VEC>::type a, typename VecT<VEC>::type b);
template <> __device__ __forceinline__ float vmax<1>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ float2 vmax<2>(float2 a, float2 b) { return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
template <> __device__ __forceinline__ float4 vmax<4>(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

template <int VEC> __device__ __forceinline__ typename VecT<VEC>::type vneg_inf();
template <> __device__ __forceinline__ float vneg_inf<1>() { return -INFINITY; }
template <> __device__ __forceinline__ float2 vneg_inf<2>() { return make_float2(-INFINITY, -INFINITY); }
template <> __device__ __forceinline__ float4 vneg_inf<4>() { return make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); }

template <int VEC, bool SORTED_ROWS>
__global__ void __launch_bounds__(kReduceWarps * 32)
pool_reduce_kernel(const float* __restrict__ feat, int32_t C, int64_t f_sb, int64_t f_sn, int32_t hw,
                   const int2* __restrict__ sorted, int32_t bn, float* rows) {
  SMOS_PDL_PROLOGUE();
  using V = typename VecT<VEC>::type;
  constexpr int kBatch = VEC == 4 ? 8 : 16;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // The valid entries of `sorted` are its first cursor[0] slots; every slot behind them holds an out-of-grid point
  // (cell < 0, written by the plan's scatter kernel). The window therefore finds its own length from the entries
  // it loads anyway instead of reading the cursor first: one dependent round trip less in a kernel that is a
  // chain of four.
  const int32_t p0 = (blockIdx.x * kReduceWarps + wib) * 32;
  if (p0 >= bn) return;
  int2 e = make_int2(0, -1);
  if (p0 + lane < bn) e = sorted[p0 + lane];
  // Point-major input (no SORTED_ROWS): a cell whose segment runs over the end of this window for fewer than 32
  // positions is FINISHED here — the warp also reads the next window's entries and folds that short tail into its
  // last piece — and the next window's warp skips the same leading run (both sides decide from the entries alone:
  // a leading run that continues the previous window's last cell and is shorter than the window belongs to the
  // previous warp). Only cells that cover at least one whole further window are still reduced as several pieces
  // ("rule B": pieces at `start` and at every window the segment covers completely, the last one including the short
  // tail). In the small grids that makes multi-piece cells rare instead of the rule (avg 4-15 points per cell).
  int2 e2 = make_int2(0, -1);
  int32_t prev_cell = -1;
  if (!SORTED_ROWS) {
    if (p0 + 32 + lane < bn) e2 = sorted[p0 + 32 + lane];
    if (p0 > 0) prev_cell = sorted[p0 - 1].y;
  }
  const int32_t cnt = __popc(__ballot_sync(0xffffffffu, e.y >= 0));  // valid entries are the leading ones
  if (cnt == 0) return;
  if (lane >= cnt) e = make_int2(0, -1 - lane);  // distinct negative cells for lanes past the end
  const int32_t prev = __shfl_up_sync(0xffffffffu, e.y, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || e.y != prev) &
                         (cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u));
  int32_t lead = 0, tail_k = 0;
  int64_t tail_off = 0;
  if (!SORTED_ROWS) {
    const unsigned same_prev = __ballot_sync(0xffffffffu, prev_cell >= 0 && e.y == prev_cell);
    const int32_t j = same_prev == 0xffffffffu ? 32 : __ffs(~same_prev) - 1;  // leading run that continues the previous window
    if (j < 32) lead = j;            // shorter than a window: the previous warp has folded it into its last piece
    if (lead >= cnt) return;         // (a run that fills the whole window is a piece of its own: lead stays 0)
    if (cnt == 32) {
      const int32_t last_cell = __shfl_sync(0xffffffffu, e.y, 31);
      const unsigned same_last = __ballot_sync(0xffffffffu, e2.y == last_cell);
      const int32_t k = same_last == 0xffffffffu ? 32 : __ffs(~same_last) - 1;
      if (k < 32) tail_k = k;        // my last cell ends inside the next window: finish it here
    }
    tail_off = (e2.y >= hw ? e2.y / hw : 0) * f_sb + static_cast<int64_t>(static_cast<uint32_t>(e2.x) & ~kMergedN) * f_sn;
  }
  // Entries to reduce, compacted to the low lanes: lane j holds the row offset (c = 0) of entry j and the position
  // of the piece it belongs to (the first position of its cell inside this window). Without SORTED_ROWS every
  // position is an entry. With SORTED_ROWS only the row owners are: a merged follower has no row (its value is
  // already in its run head's row), so ~55 % of the positions cost nothing here.
  int32_t m = cnt;
  int64_t ent_off;
  int32_t ent_idx = lane;  // window index of my entry
  if (SORTED_ROWS) {
    const bool absent = lane < cnt && (static_cast<uint32_t>(e.x) & kMergedN) != 0u;
    const unsigned owners = __ballot_sync(0xffffffffu, lane < cnt && !absent);
    m = __popc(owners);
    ent_idx = lane < m ? static_cast<int32_t>(__fns(owners, 0, lane + 1)) : 0;
    ent_off = static_cast<int64_t>(p0 + ent_idx) * C;
  } else {
    // (batch of the entry: cells of the first batch — all of them in the streaming case — need no quotient)
    ent_off = (e.y >= hw ? e.y / hw : 0) * f_sb + static_cast<int64_t>(static_cast<uint32_t>(e.x) & ~kMergedN) * f_sn;
  }
  const int32_t ent_pp = p0 + (31 - __clz(heads & (0xffffffffu >> (31 - ent_idx))));  // head at or before my entry
  // SORTED_ROWS: a window can begin with followers of a run whose head sits in the previous window; if their cell
  // has no owner in this window the piece is empty, but the combine stage reads a row at its position: -inf
  const bool lead_empty = SORTED_ROWS && (m == 0 || __shfl_sync(0xffffffffu, ent_pp, 0) != p0);
  const float* src = SORTED_ROWS ? rows : feat;

  for (int32_t c0 = 0; c0 < C; c0 += 32 * VEC) {
    const int32_t c = c0 + lane * VEC;
    const bool c_ok = c < C;  // C % VEC == 0 is guaranteed by the dispatcher
    const int32_t c_ld = c_ok ? c : 0;  // lanes past C load a valid address and discard the value
    if (lead_empty && c_ok) *reinterpret_cast<V*>(rows + static_cast<int64_t>(p0) * C + c) = vneg_inf<VEC>();
    V acc;
    int32_t piece_pos = -1;
    for (int32_t j0 = lead; j0 < m; j0 += kBatch) {
      V v[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        // UNCONDITIONAL loads (entries past the end re-read the last one): a predicated load makes ptxas funnel
        // every value through one temporary register, which serialises the batch
        const int32_t jj = min(j0 + u, m - 1);
        const int64_t off = __shfl_sync(0xffffffffu, ent_off, jj);
        v[u] = __ldg(reinterpret_cast<const V*>(src + off + c_ld));
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int32_t jj = j0 + u;
        const int32_t pp = __shfl_sync(0xffffffffu, ent_pp, min(jj, m - 1));
        if (jj < m && c_ok) {
          if (pp != piece_pos) {  // warp uniform
            if (piece_pos >= 0) *reinterpret_cast<V*>(rows + static_cast<int64_t>(piece_pos) * C + c) = acc;
            piece_pos = pp;
            acc = v[u];
          } else {
            acc = vmax<VEC>(acc, v[u]);
          }
        }
      }
    }
    if (!SORTED_ROWS && tail_k > 0) {  // warp uniform: the short tail of my last cell in the next window
      for (int32_t t0 = 0; t0 < tail_k; t0 += kBatch) {
        V v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int64_t off = __shfl_sync(0xffffffffu, tail_off, min(t0 + u, tail_k - 1));
          v[u] = __ldg(reinterpret_cast<const V*>(src + off + c_ld));
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) acc = vmax<VEC>(acc, v[u]);  // (re-read rows repeat: max is idempotent)
      }
    }
    if (c_ok && piece_pos >= 0) *reinterpret_cast<V*>(rows + static_cast<int64_t>(piece_pos) * C + c) = acc;
  }
}

// Number of piece rows a cell's segment [s, s + k) leaves behind the row at s. Rule A (channel-major path): one per
// further window the segment touches. Rule B (point-major path, see pool_reduce_kernel): one per further window it
// covers completely. A cell is "multi-piece" when this is > 0; its pieces sit at s and at first + 32 i.
__device__ __forceinline__ int32_t pool_extra_pieces(int32_t s, int32_t k, bool rule_b) {
  const int32_t first = (s & ~31) + 32, end = s + k;
  if (k <= 0 || end <= first) return 0;
  return rule_b ? (end - first) >> 5 : (end - first + 31) >> 5;
}

// ---- phase A2: fold the pieces of multi-piece cells into their first row ------------------------
// One warp per listed cell, lanes over channels, piece rows loaded kBatch at a time. After this every
// occupied cell's maxima sit in rows[start[cell]].
template <int VEC>
__global__ void __launch_bounds__(kReduceWarps * 32)
pool_combine_kernel(int32_t C, const int2* __restrict__ multi, const int32_t* __restrict__ cursor, float* rows,
                    int rule_b) {
  SMOS_PDL_PROLOGUE();
  using V = typename VecT<VEC>::type;
  constexpr int kBatch = VEC == 4 ? 8 : 16;
  const int lane = threadIdx.x & 31;
  const int32_t w = blockIdx.x * kReduceWarps + (threadIdx.x >> 5);
  if (w >= cursor[1]) return;
  const int2 sk = multi[w];
  const int32_t s = sk.x;
  const int32_t first_aligned = (s & ~31) + 32;
  const int32_t npieces = 1 + pool_extra_pieces(s, sk.y, rule_b != 0);
  if (npieces == 1) return;  // listed under rule A, single piece under rule B
  for (int32_t c0 = 0; c0 < C; c0 += 32 * VEC) {
    const int32_t c = c0 + lane * VEC;
    if (c >= C) continue;
    V acc = *reinterpret_cast<const V*>(rows + static_cast<int64_t>(s) * C + c);
    for (int32_t p0 = 1; p0 < npieces; p0 += kBatch) {
      V v[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int32_t pi = min(p0 + u, npieces - 1);  // unconditional loads, see pool_reduce_kernel
        v[u] = *reinterpret_cast<const V*>(rows + static_cast<int64_t>(first_aligned + (pi - 1) * 32) * C + c);
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u)
        if (p0 + u < npieces) acc = vmax<VEC>(acc, v[u]);
    }
    *reinterpret_cast<V*>(rows + static_cast<int64_t>(s) * C + c) = acc;
  }
}

// Fold of the multi-piece cells by the first CTAs of the writer launch (see pool_write_kernel).
__device__ __forceinline__ void pool_fold_multi(const float* __restrict__ rows, int32_t C, int32_t hw,
                                             const int32_t* __restrict__ count, float* __restrict__ out,
                                             const int2* __restrict__ multi, const int2* __restrict__ sorted,
                                             int32_t fold_ctas, int32_t multi_cap, int32_t bx, bool rule_b) {
  constexpr int kBatch = 16;
  const int lane = threadIdx.x & 31;
  const int32_t nwarps = fold_ctas * (kWriteThreads / 32);
  int32_t w = bx * (kWriteThreads / 32) + (threadIdx.x >> 5);
  // the list entry is fetched together with the list length (the list has room for multi_cap entries; what
  // lies behind the length is never used)
  int2 sk = w < multi_cap ? __ldg(multi + w) : make_int2(0, 0);
  const int32_t nmulti = __ldg(count + (static_cast<int64_t>(hw) * gridDim.z) + 1);  // cursor[1]
  for (; w < nmulti; w += nwarps, sk = w < nmulti ? __ldg(multi + w) : sk) {
    const int32_t s0 = sk.x;
    const int32_t first = (s0 & ~31) + 32;
    const int32_t npieces = 1 + pool_extra_pieces(s0, sk.y, rule_b);
    if (npieces == 1) continue;  // listed under rule A, single piece under rule B: the writer thread has it
    const int32_t gcell = __ldg(&sorted[s0].y);
    const int32_t bb = gcell / hw, cell = gcell - bb * hw;
    for (int32_t c0 = 0; c0 < C; c0 += 32) {
      const int32_t c = min(c0 + lane, C - 1);
      float acc = __ldg(rows + static_cast<int64_t>(s0) * C + c);
      for (int32_t p0 = 1; p0 < npieces; p0 += kBatch) {
        float v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u)  // unconditional loads: the last piece is re-read past the end
          v[u] = __ldg(rows + static_cast<int64_t>(first + (min(p0 + u, npieces - 1) - 1) * 32) * C + c);
#pragma unroll
        for (int u = 0; u < kBatch; ++u) acc = fmaxf(acc, v[u]);
      }
      if (c0 + lane < C) out[(static_cast<int64_t>(bb) * C + c) * hw + cell] = acc;
    }
  }
}

// ---- phase B: dense writer --------------------------------------------------------------------
// thread = 4 adjacent cells (along W); it walks the channel groups (kCG channels each) of its CTA's
// channel range, so one read of count/start feeds up to C output stores. An occupied cell reads its
// one row (rows[start]); all row loads of a channel group are issued together. Empty cells store zeros.
// Every output element is written exactly once with 128-bit stores.
// FOLD: the first `fold_ctas` CTAs of the launch (blockIdx.z == 0) are not writers: their
// warps walk the plan's list of multi-piece cells (segments that cross a multiple of 32, reduced as several pieces),
// fold each cell's piece rows (lanes over channels, 16 rows in flight) and store the C maxima straight into the
// output; the writer threads skip exactly those cells. This used to be a separate launch between the reduction and
// the writer (pool_combine_kernel: 5-9 us of launch + a chain of dependent round trips per pooling call); inside
// the writer launch it runs beside the dense stores and nothing waits for it. (Folding inside the writer THREADS was
// measured and rejected: a thread owns 4 cells x 8 channels and a hot cell has up to 47 pieces — the small writers
// went from 4 to 9-21 us, the big one from 42 to 76 us.)
template <bool VEC4, bool FOLD>
__global__ void __launch_bounds__(kWriteThreads)
pool_write_kernel(const float* __restrict__ rows, int32_t C, int32_t hw, int32_t groups_per_cta,
                  const int32_t* __restrict__ count, const int32_t* __restrict__ start,
                  float* __restrict__ out, int stream_out, const int2* __restrict__ multi,
                  const int2* __restrict__ sorted, int32_t fold_ctas, int32_t multi_cap, int rule_b, int32_t tiles_x) {
  SMOS_PDL_PROLOGUE();
  int32_t bx = blockIdx.x;
  if (FOLD) {
    if (bx < fold_ctas) {
      if (blockIdx.z != 0) return;
      pool_fold_multi(rows, C, hw, count, out, multi, sorted, fold_ctas, multi_cap, bx, rule_b != 0);
      return;
    }
    bx -= fold_ctas;
  }
  // writer CTAs: (cell tile, channel-group slice) flattened into blockIdx.x so that the fold CTAs exist once per launch
  const int32_t by = bx / tiles_x;
  bx -= by * tiles_x;
  const int32_t b = blockIdx.z;
  const bool vec_rows = ((C & 7) == 0);
  constexpr int CPT = VEC4 ? 4 : 1;  // cells per thread
  const int32_t cell0 = (bx * kWriteThreads + threadIdx.x) * CPT;
  if (cell0 >= hw) return;
  const int32_t g0 = b * hw + cell0;
  int32_t k[CPT], s[CPT];
#pragma unroll
  for (int q = 0; q < CPT; ++q) { k[q] = 0; s[q] = 0; }
  if (VEC4) {
    const int4 kk = __ldg(reinterpret_cast<const int4*>(count + g0));
    k[0] = kk.x; k[1 % CPT] = kk.y; k[2 % CPT] = kk.z; k[3 % CPT] = kk.w;
  } else {
    k[0] = __ldg(count + g0);
  }
  bool any = false;
#pragma unroll
  for (int q = 0; q < CPT; ++q) any |= k[q] > 0;
  // outputs that fit L2 are latency bound: fetch `start` together with `count` (the HBM-bound big writer keeps the
  // conditional second load — unconditional it was 39.5 -> 45.3 us)
  if (any || !stream_out) {
    if (VEC4) {
      const int4 ss = __ldg(reinterpret_cast<const int4*>(start + g0));
      s[0] = ss.x; s[1 % CPT] = ss.y; s[2 % CPT] = ss.z; s[3 % CPT] = ss.w;
    } else {
      s[0] = __ldg(start + g0);
    }
  }
  unsigned multi_mask = 0u;  // cells of mine whose segment crosses a multiple of 32: written by the fold warps
  if (FOLD) {
#pragma unroll
    for (int q = 0; q < CPT; ++q)
      if (pool_extra_pieces(s[q], k[q], rule_b != 0) > 0) multi_mask |= 1u << q;
  }
  const bool any_multi = multi_mask != 0u;
  const int32_t ngroups = (C + kCG - 1) / kCG;
  const int32_t cg_begin = by * groups_per_cta;
  const int32_t cg_end = min(ngroups, cg_begin + groups_per_cta);
  for (int32_t cg = cg_begin; cg < cg_end; ++cg) {
    const int32_t c0 = cg * kCG;
    const int32_t nch = min(kCG, C - c0);
    float v[CPT][kCG];
#pragma unroll
    for (int q = 0; q < CPT; ++q)
#pragma unroll
      for (int j = 0; j < kCG; ++j) v[q][j] = 0.f;
    if (any) {
      // all row loads of the thread are unconditional (empty cells re-read row s = 0, which is valid
      // whenever any cell is occupied) and selected afterwards, so they are in flight together
      if (vec_rows) {
        float4 a[CPT], bb[CPT];
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
          const float4* rp = reinterpret_cast<const float4*>(rows + static_cast<int64_t>(s[q]) * C + c0);
          a[q] = __ldg(rp);
          bb[q] = __ldg(rp + 1);
        }
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
          const bool occ = k[q] > 0;
          v[q][0] = occ ? a[q].x : 0.f; v[q][1] = occ ? a[q].y : 0.f; v[q][2] = occ ? a[q].z : 0.f;
          v[q][3] = occ ? a[q].w : 0.f; v[q][4] = occ ? bb[q].x : 0.f; v[q][5] = occ ? bb[q].y : 0.f;
          v[q][6] = occ ? bb[q].z : 0.f; v[q][7] = occ ? bb[q].w : 0.f;
        }
      } else {
#pragma unroll
        for (int q = 0; q < CPT; ++q)
#pragma unroll
          for (int j = 0; j < kCG; ++j) {
            const float t = __ldg(rows + static_cast<int64_t>(s[q]) * C + c0 + (j < nch ? j : 0));
            v[q][j] = (k[q] > 0 && j < nch) ? t : 0.f;
          }
      }
    }
    float* ob = out + (static_cast<int64_t>(b) * C + c0) * hw + cell0;
    if (FOLD && any_multi) {  // rare: cells owned by the fold warps are left alone, the others leave one by one
#pragma unroll
      for (int j = 0; j < kCG; ++j)
        if (j < nch) {
#pragma unroll
          for (int q = 0; q < CPT; ++q)
            if (!((multi_mask >> q) & 1u)) ob[static_cast<int64_t>(j) * hw + q] = v[q][j];
        }
      continue;
    }
#pragma unroll
    for (int j = 0; j < kCG; ++j) {
      if (j < nch) {
        if (VEC4) {
          const float4 o = make_float4(v[0][j], v[1 % CPT][j], v[2 % CPT][j], v[3 % CPT][j]);
          float* dst = ob + static_cast<int64_t>(j) * hw;
          if (stream_out) smos_st_cs_f4(dst, o);
          else *reinterpret_cast<float4*>(dst) = o;
        } else {
          ob[static_cast<int64_t>(j) * hw] = v[0][j];
        }
      }
    }
  }
}

// Backward: equality-mask gather (reference .cu:109-132). `fast_n` selects which index runs
// fastest across threads so that the feature/grad accesses coalesce for either layout.
__global__ void __launch_bounds__(256)
pool_backward_kernel(const float* __restrict__ feat, int32_t C, int32_t N, int64_t total,
                     int64_t f_sb, int64_t f_sc, int64_t f_sn, int64_t hw,
                     const int32_t* __restrict__ cell_in, const float* __restrict__ vout,
                     const float* __restrict__ gout, float* __restrict__ gfeat,
                     int64_t g_sb, int64_t g_sc, int64_t g_sn, int fast_n) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int32_t b, c, n;
  const int64_t cn = static_cast<int64_t>(C) * N;
  b = static_cast<int32_t>(i / cn);
  const int64_t r = i - b * cn;
  if (fast_n) { c = static_cast<int32_t>(r / N); n = static_cast<int32_t>(r - static_cast<int64_t>(c) * N); }
  else        { n = static_cast<int32_t>(r / C); c = static_cast<int32_t>(r - static_cast<int64_t>(n) * C); }
  const int32_t cell = cell_in[static_cast<int64_t>(b) * N + n];
  float g = 0.f;
  if (cell >= 0) {
    const int64_t o = (static_cast<int64_t>(b) * C + c) * hw + cell;
    const float f = feat[b * f_sb + c * f_sc + n * f_sn];
    if (__ldg(vout + o) == f) g = __ldg(gout + o);
  }
  gfeat[b * g_sb + c * g_sc + n * g_sn] = g;
}

}  // namespace

extern "C" {

int64_t smos_pool_plan_bytes(int64_t B, int64_t N, int32_t H, int32_t W) {
  if (B <= 0 || N < 0 || H <= 0 || W <= 0) return SMOS_EINVAL;
  return smos_pool_layout(B, N, H, W).bytes;
}

int64_t smos_pool_workspace_bytes(int64_t B, int64_t C, int64_t N) {
  if (B <= 0 || C <= 0 || N < 0) return SMOS_EINVAL;
  return smos_align_up(B * N * C * 4 + 256, 256);
}

int smos_pool_plan_build_multi(const smos_pool_plan_desc* descs_host, int32_t n, void* stream) {
  if (descs_host == nullptr || n <= 0 || n > kMaxPlans) return SMOS_EINVAL;
  PlanBatch pb;
  pb.n = n;
  int64_t pt = 0, quad = 0;
  for (int32_t j = 0; j < n; ++j) {
    const smos_pool_plan_desc& d = descs_host[j];
    if (d.B <= 0 || d.N < 0 || d.H <= 0 || d.W <= 0 || d.plan == nullptr) return SMOS_EINVAL;
    if (d.B * d.N >= (int64_t(1) << 31) || d.B * static_cast<int64_t>(d.H) * d.W >= (int64_t(1) << 31))
      return SMOS_EUNSUPPORTED;
    if (d.N > 0 && d.pcds_ind == nullptr) return SMOS_EINVAL;
    const PoolLayout L = smos_pool_layout(d.B, d.N, d.H, d.W);
    char* base = static_cast<char*>(d.plan);
    PlanDev& P = pb.p[j];
    P.ind = d.pcds_ind; P.ind_sb = d.ind_sb; P.ind_sn = d.ind_sn; P.ind_sd = d.ind_sd;
    P.vmi = d.voxel_max_idx; P.vmi_stride = d.idx_batch_stride;
    P.cell = reinterpret_cast<int32_t*>(base + L.off_cell);
    P.rank = reinterpret_cast<int32_t*>(base + L.off_rank);
    P.count = reinterpret_cast<int32_t*>(base + L.off_count);
    P.start = reinterpret_cast<int32_t*>(base + L.off_start);
    P.sorted = reinterpret_cast<int2*>(base + L.off_sorted);
    P.multi = reinterpret_cast<int2*>(base + L.off_multi);
    if (d.gather_taps != nullptr && (reinterpret_cast<uintptr_t>(d.gather_taps) & 15) != 0) return SMOS_EINVAL;
    P.taps = static_cast<TapsS*>(d.gather_taps);
    P.N = static_cast<int32_t>(d.N); P.H = d.H; P.W = d.W;
    P.hw = static_cast<int32_t>(L.hw); P.cells = static_cast<int32_t>(L.cells);
    P.bn = static_cast<int32_t>(d.B * d.N);
    P.sh = d.scale_h; P.sw = d.scale_w;
    P.scale_dev = d.scale_dev;
    if (d.scale_dev != nullptr && d.gather_taps != nullptr) return SMOS_EUNSUPPORTED;  // the records need host scales
    P.pt_begin = pt; P.quad_begin = quad;
    pt += d.B * d.N;
    quad += smos_align_up((L.cells + 3) / 4, 32);
  }
  pb.pt_total = pt;
  pb.quad_total = quad;
  cudaStream_t st = smos_stream(stream);
  // zero the per-cell counters and the cursors of every plan: one launch (a cudaMemsetAsync node per plan, or one
  // over the whole span of the plans, cost more than the three plan kernels together)
  SMOS_LAUNCH((pool_zero_counts_kernel), smos_ceil_div(quad, kPlanThreads), kPlanThreads, 0, st, pb);
  if (pt > 0) SMOS_LAUNCH((pool_cell_index_kernel), smos_ceil_div(pt, kPlanThreads), kPlanThreads, 0, st, pb);
  SMOS_LAUNCH((pool_cell_alloc_kernel), smos_ceil_div(quad, kPlanThreads), kPlanThreads, 0, st, pb);
  if (pt > 0) SMOS_LAUNCH((pool_cell_scatter_kernel), smos_ceil_div(pt, kPlanThreads), kPlanThreads, 0, st, pb);
  return smos_launch_status();
}

int smos_pool_plan_build(const float* pcds_ind, int64_t B, int64_t N, int64_t ind_sb, int64_t ind_sn,
                         int64_t ind_sd, int32_t H, int32_t W, float scale_h, float scale_w,
                         int64_t* voxel_max_idx, int64_t idx_batch_stride, void* plan, void* stream) {
  smos_pool_plan_desc d;
  d.pcds_ind = pcds_ind; d.B = B; d.N = N; d.ind_sb = ind_sb; d.ind_sn = ind_sn; d.ind_sd = ind_sd;
  d.H = H; d.W = W; d.scale_h = scale_h; d.scale_w = scale_w;
  d.voxel_max_idx = voxel_max_idx; d.idx_batch_stride = idx_batch_stride; d.plan = plan;
  d.gather_taps = nullptr;
  d.scale_dev = nullptr;
  return smos_pool_plan_build_multi(&d, 1, stream);
}

int smos_voxel_maxpool_forward(const float* pcds_feat, int64_t B, int64_t C, int64_t N, int64_t f_sb,
                               int64_t f_sc, int64_t f_sn, int32_t H, int32_t W, const void* plan,
                               void* workspace, float* voxel_out, void* stream) {
  return smos_voxel_maxpool_forward_stages(pcds_feat, B, C, N, f_sb, f_sc, f_sn, H, W, plan, workspace, voxel_out,
                                           SMOS_POOL_STAGE_ALL, stream);
}

int smos_voxel_maxpool_forward_stages(const float* pcds_feat, int64_t B, int64_t C, int64_t N, int64_t f_sb,
                                      int64_t f_sc, int64_t f_sn, int32_t H, int32_t W, const void* plan,
                                      void* workspace, float* voxel_out, int32_t stages, void* stream) {
  if (B <= 0 || C <= 0 || N < 0 || H <= 0 || W <= 0 || plan == nullptr || voxel_out == nullptr) return SMOS_EINVAL;
  if (N > 0 && (pcds_feat == nullptr || workspace == nullptr)) return SMOS_EINVAL;
  if (B * N >= (int64_t(1) << 31) || C >= (1 << 20) || B > 65535) return SMOS_EUNSUPPORTED;
  const PoolLayout L = smos_pool_layout(B, N, H, W);
  const char* base = static_cast<const char*>(plan);
  const int32_t* count = reinterpret_cast<const int32_t*>(base + L.off_count);
  const int32_t* start = reinterpret_cast<const int32_t*>(base + L.off_start);
  const int2* sorted = reinterpret_cast<const int2*>(base + L.off_sorted);
  const int32_t* cursor = count + L.cells;
  float* rows = static_cast<float*>(workspace);
  const int32_t* pos = reinterpret_cast<const int32_t*>(base + L.off_rank);
  cudaStream_t st = smos_stream(stream);
  const int32_t hw = static_cast<int32_t>(L.hw);
  const int32_t Ci = static_cast<int32_t>(C);
  const int64_t total = B * N;
  const bool point_major = (f_sc == 1 && C > 1);
  const int rule_b = point_major ? 1 : 0;  // which reduction produced the piece rows (pool_extra_pieces)
  if (total > 0 && (stages & SMOS_POOL_STAGE_REDUCE)) {
    const int grid = smos_ceil_div(total, kReduceWarps * 32);
    const bool aligned = (reinterpret_cast<uintptr_t>(pcds_feat) & 15) == 0 && (f_sb & 3) == 0 && (f_sn & 3) == 0;
    if (!point_major) {
      // channel-major: permute into sorted rows, then reduce those rows in place
      static std::atomic<unsigned long long> opted_generic{0}, opted_ldg1{0}, opted_ldg2{0};
      int device = 0;
      cudaGetDevice(&device);
      const size_t smem = static_cast<size_t>(kPermPts) * (C + 1) * 4;
      if (smem > 200 * 1024) return SMOS_EUNSUPPORTED;
      if (cudaError_t e = smos_smem_opt_in(pool_permute_kernel, opted_generic, 200 * 1024); e != cudaSuccess)
        return static_cast<int>(e);
      const int depth = perm_stages();
      const bool tma_ok = f_sn == 1 && (N & 3) == 0 && (f_sb & 3) == 0 && (f_sc & 3) == 0 && (C % kTmCh) == 0 &&
                          (reinterpret_cast<uintptr_t>(pcds_feat) & 15) == 0 && N >= kTmPts && encode_tiled() != nullptr;
      const bool vec_ok = f_sn == 1 && (N & 3) == 0 && (f_sb & 3) == 0 && (f_sc & 3) == 0 && N >= 4 &&
                          (reinterpret_cast<uintptr_t>(pcds_feat) & 15) == 0 &&
                          static_cast<size_t>(C) * (kPermPts + 4) * 4 <= 200 * 1024;
      if (vec_ok && env_int("SMOS_PERM_LDG", 1)) {  // default; SMOS_PERM_LDG=0 selects the TMA pipeline below
        if (cudaError_t e = smos_smem_opt_in(pool_permute_ldg_kernel<1>, opted_ldg1, 200 * 1024); e != cudaSuccess)
          return static_cast<int>(e);
        if (cudaError_t e = smos_smem_opt_in(pool_permute_ldg_kernel<2>, opted_ldg2, 200 * 1024); e != cudaSuccess)
          return static_cast<int>(e);
        dim3 pg(smos_ceil_div(N, kPermPts), static_cast<unsigned>(B));
        const size_t lsmem = static_cast<size_t>(C) * (kPermPts + 4) * 4;
        if (C > 32)
          SMOS_LAUNCH((pool_permute_ldg_kernel<2>), pg, kPermThreads, lsmem, st, pcds_feat, Ci, static_cast<int32_t>(N), f_sb, f_sc, pos, rows);
        else
          SMOS_LAUNCH((pool_permute_ldg_kernel<1>), pg, kPermThreads, lsmem, st, pcds_feat, Ci, static_cast<int32_t>(N), f_sb, f_sc, pos, rows);
      } else if (tma_ok) {
        int rc;
        if (depth == 2) rc = launch_permute_tma<2>(pcds_feat, Ci, N, B, f_sb, f_sc, pos, rows, device, st);
        else if (depth == 4) rc = launch_permute_tma<4>(pcds_feat, Ci, N, B, f_sb, f_sc, pos, rows, device, st);
        else rc = launch_permute_tma<3>(pcds_feat, Ci, N, B, f_sb, f_sc, pos, rows, device, st);
        if (rc != SMOS_OK) return rc;
      } else {
        dim3 pg(smos_ceil_div(N, kPermPts), static_cast<unsigned>(B));
        SMOS_LAUNCH((pool_permute_kernel), pg, kPermThreads, smem, st, pcds_feat, Ci, static_cast<int32_t>(N), f_sb, f_sc, f_sn, pos, rows);
      }
      if (stages & SMOS_POOL_STAGE_NO_REDUCE) {}
      else if ((C & 127) == 0) SMOS_LAUNCH((pool_reduce_kernel<4, true>), grid, kReduceWarps * 32, 0, st, rows, Ci, 0, 0, hw, sorted, static_cast<int32_t>(total), rows);
      else if ((C & 63) == 0) SMOS_LAUNCH((pool_reduce_kernel<2, true>), grid, kReduceWarps * 32, 0, st, rows, Ci, 0, 0, hw, sorted, static_cast<int32_t>(total), rows);
      else SMOS_LAUNCH((pool_reduce_kernel<1, true>), grid, kReduceWarps * 32, 0, st, rows, Ci, 0, 0, hw, sorted, static_cast<int32_t>(total), rows);
    } else {
      if ((C & 127) == 0 && aligned) SMOS_LAUNCH((pool_reduce_kernel<4, false>), grid, kReduceWarps * 32, 0, st, pcds_feat, Ci, f_sb, f_sn, hw, sorted, static_cast<int32_t>(total), rows);
      else if ((C & 63) == 0 && aligned) SMOS_LAUNCH((pool_reduce_kernel<2, false>), grid, kReduceWarps * 32, 0, st, pcds_feat, Ci, f_sb, f_sn, hw, sorted, static_cast<int32_t>(total), rows);
      else SMOS_LAUNCH((pool_reduce_kernel<1, false>), grid, kReduceWarps * 32, 0, st, pcds_feat, Ci, f_sb, f_sn, hw, sorted, static_cast<int32_t>(total), rows);
    }
  }
  // multi-piece cells: folded by dedicated CTAs of the writer launch (default) or by the separate combine kernel
  // (SMOS_POOL_FOLD=0, kept for A/B runs)
  // Default (2): inside the writer launch for outputs that fit L2 (the latency-bound small grids, -0.5 us per call and
  // one launch less); the HBM-bound 201 MB writer keeps the lean variant + the combine kernel — with the fold path
  // compiled in, ptxas allocates 63 instead of 84 registers, 4 CTAs per SM become resident and the writer slows
  // from 42 to 45 us (more CTAs per SM were measured slower before). 1: always, 0: never.
  const int fold_mode = env_int("SMOS_POOL_FOLD", 2);
  const bool big_out = B * C * L.hw * 4 > (int64_t(96) << 20);
  const bool fold = (fold_mode == 1 || (fold_mode == 2 && !big_out)) && total >= 32;
  if (!fold && total >= 32 && (stages & SMOS_POOL_STAGE_COMBINE)) {
    // fold multi-piece cells (<= total/32 of them; the exact number is only known on the device)
    const int2* multi = reinterpret_cast<const int2*>(base + L.off_multi);
    const int cgrid = smos_ceil_div(total / 32 + 1, kReduceWarps);
    if ((C & 127) == 0) SMOS_LAUNCH((pool_combine_kernel<4>), cgrid, kReduceWarps * 32, 0, st, Ci, multi, cursor, rows, rule_b);
    else if ((C & 63) == 0) SMOS_LAUNCH((pool_combine_kernel<2>), cgrid, kReduceWarps * 32, 0, st, Ci, multi, cursor, rows, rule_b);
    else SMOS_LAUNCH((pool_combine_kernel<1>), cgrid, kReduceWarps * 32, 0, st, Ci, multi, cursor, rows, rule_b);
  }
  // outputs beyond L2 capacity are written with evict-first stores
  const int stream_out = big_out ? 1 : 0;
  const bool vec4 = ((hw & 3) == 0) && ((reinterpret_cast<uintptr_t>(voxel_out) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(count) & 15) == 0);
  const int cpt = vec4 ? 4 : 1;
  const int gx = smos_ceil_div(smos_ceil_div(hw, cpt), kWriteThreads);
  // split the channel groups over blockIdx.y only as far as needed to cover the SMs ~3x
  const int32_t ngroups = (Ci + kCG - 1) / kCG;
  int32_t groups_per_cta = ngroups;
  while (groups_per_cta > 1 && static_cast<int64_t>(gx) * B * ((ngroups + groups_per_cta - 1) / groups_per_cta) < 3 * SMOS_SM_COUNT)
    groups_per_cta = (groups_per_cta + 1) / 2;
  // fold CTAs come first in the launch (their chains are the longest): one warp per possible list entry (a warp that
  // walks several entries pays the whole chain of dependent round trips once per entry), at most four CTAs per SM
  const int32_t multi_cap = static_cast<int32_t>(total / 32 + 1);
  int32_t fold_ctas = 0;
  if (fold) {
    fold_ctas = smos_ceil_div(multi_cap, (kWriteThreads / 32) * env_int("SMOS_FOLD_PER_WARP", 1));
    if (fold_ctas > 4 * SMOS_SM_COUNT) fold_ctas = 4 * SMOS_SM_COUNT;
    if (fold_ctas < 1) fold_ctas = 1;
  }
  const int2* multi_list = reinterpret_cast<const int2*>(base + L.off_multi);
  const int32_t gy = (ngroups + groups_per_cta - 1) / groups_per_cta;
  dim3 grid(static_cast<unsigned>(gx) * gy + fold_ctas, 1, static_cast<unsigned>(B));
  if (stages & SMOS_POOL_STAGE_WRITE) {
    // experiment knob: dynamic shared memory nobody uses caps the resident CTAs per SM of the HBM-bound big writer
    const size_t wsmem = stream_out ? static_cast<size_t>(env_int("SMOS_WRITE_SMEM_KB", 0)) * 1024 : 0;
    if (wsmem > 48 * 1024) {
      static std::atomic<unsigned long long> o1{0}, o2{0};
      if (cudaError_t e = smos_smem_opt_in(pool_write_kernel<true, true>, o1, 200 * 1024); e != cudaSuccess) return static_cast<int>(e);
      if (cudaError_t e = smos_smem_opt_in(pool_write_kernel<true, false>, o2, 200 * 1024); e != cudaSuccess) return static_cast<int>(e);
    }
#define SMOS_LAUNCH_WRITE(V, F)                                                                                   \
    SMOS_LAUNCH((pool_write_kernel<V, F>), grid, kWriteThreads, wsmem, st, rows, Ci, hw, groups_per_cta, count, start, \
                voxel_out, stream_out, multi_list, sorted, fold_ctas, multi_cap, rule_b, gx)
    if (vec4 && fold) SMOS_LAUNCH_WRITE(true, true);
    else if (vec4) SMOS_LAUNCH_WRITE(true, false);
    else if (fold) SMOS_LAUNCH_WRITE(false, true);
    else SMOS_LAUNCH_WRITE(false, false);
#undef SMOS_LAUNCH_WRITE
  }
  return smos_launch_status();
}

int smos_voxel_maxpool_backward(const float* pcds_feat, int64_t B, int64_t C, int64_t N, int64_t f_sb,
                                int64_t f_sc, int64_t f_sn, int32_t H, int32_t W, const void* plan,
                                const float* voxel_out, const float* grad_voxel_out, float* grad_feat,
                                int64_t g_sb, int64_t g_sc, int64_t g_sn, void* stream) {
  if (B <= 0 || C <= 0 || N < 0 || H <= 0 || W <= 0 || plan == nullptr) return SMOS_EINVAL;
  if (N == 0) return SMOS_OK;
  if (!pcds_feat || !voxel_out || !grad_voxel_out || !grad_feat) return SMOS_EINVAL;
  const PoolLayout L = smos_pool_layout(B, N, H, W);
  const int32_t* cell = reinterpret_cast<const int32_t*>(static_cast<const char*>(plan) + L.off_cell);
  const int64_t total = B * C * N;
  const int fast_n = (f_sc == 1 && C > 1) ? 0 : 1;
  SMOS_LAUNCH((pool_backward_kernel), smos_ceil_div(total, 256), 256, 0, smos_stream(stream), 
      pcds_feat, static_cast<int32_t>(C), static_cast<int32_t>(N), total, f_sb, f_sc, f_sn, L.hw, cell, voxel_out,
      grad_voxel_out, grad_feat, g_sb, g_sc, g_sn, fast_n);
  return smos_launch_status();
}

}  // extern "C"
