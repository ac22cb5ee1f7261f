// VoxelMaxPool for sm_100a: sort-by-cell, balanced segmented max, dense one-pass writer.
//
// The reference (deep_point/src/point_deep_cuda_kernel.cu:24-99) zero-fills the dense grid,
// stores every point feature once (racy init), then issues one CAS-loop atomic per
// (point, channel) into global memory. LiDAR density is extremely skewed (hundreds of points
// in one BEV cell next to the sensor, >90 % of the cells empty), so neither per-cell atomics
// nor per-tile ownership balance. Here:
//
//   plan  (3 small kernels, coordinates only, shared by forward and backward)
//         cell index -> per-cell counting sort: `sorted` lists the valid points grouped by
//         cell; count[cell] / start[cell] locate each group.
//   reduce (phase A) one warp per 32 consecutive sorted points, whatever cells they fall in:
//         perfectly balanced. Inside its 32 points a warp reduces each run of equal cells (a
//         "piece") and writes ONE row of C maxima per piece into a scratch row buffer, indexed
//         by the position of the piece's first point. No atomics, no initialisation.
//   write (phase B) one thread per 4 adjacent output cells x 8 channels: empty cells store
//         zeros, occupied cells combine their <= 1 + count/32 piece rows. Every output element
//         is written exactly once with 128-bit stores: the 201 MB zero-fill of the reference
//         is fused away and HBM traffic stays at the algorithmic minimum.
#include "common.cuh"

namespace {

constexpr int kPlanThreads = 256;
constexpr int kReduceWarps = 8;   // warps per CTA in phase A
constexpr int kWriteThreads = 256;
constexpr int kCG = 8;            // channels per thread in phase B
// A point that continues a run of equal cells begun by the lane before it (same 64-point tile, hence
// consecutive sorted positions) is "merged": on the channel-major path only the run's first point gets a
// row (the run maximum), which removes ~60 % of the permute's writes and of the reduction's reads in scan
// order. Bit 30 of a plan position / bit 31 of sorted.x carry the mark.
constexpr int32_t kMergedPos = 0x40000000;
constexpr int32_t kPosMask = 0x3fffffff;
constexpr uint32_t kMergedN = 0x80000000u;

struct PoolLayout {
  int64_t hw, cells;   // H*W, B*H*W
  int64_t off_cell, off_rank, off_count, off_start, off_sorted, off_multi, bytes;
};

PoolLayout pool_layout(int64_t B, int64_t N, int32_t H, int32_t W) {
  PoolLayout L;
  L.hw = static_cast<int64_t>(H) * W;
  L.cells = B * L.hw;
  const int64_t bn = B * N;
  int64_t off = 0;
  L.off_cell = off;   off += smos_align_up(bn * 4, 256);
  L.off_rank = off;   off += smos_align_up(bn * 4, 256);
  L.off_count = off;  off += smos_align_up((L.cells + 4) * 4, 256);  // + point cursor, multi-cell counter
  L.off_start = off;  off += smos_align_up(L.cells * 4, 256);
  L.off_sorted = off; off += smos_align_up(bn * 8, 256);
  // cells whose segment crosses a multiple of 32 (>= 2 pieces): at most one per 32 sorted points
  L.off_multi = off;  off += smos_align_up((bn / 32 + 2) * 8, 256);
  L.bytes = off;
  return L;
}

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
  int64_t pt_begin;    // first global point index of this plan
  int64_t quad_begin;  // first global cell-quad index (warp aligned)
  int32_t N, H, W, hw, cells;
  float sh, sw;
};

struct PlanBatch {
  PlanDev p[kMaxPlans];
  int64_t pt_total, quad_total;
  int32_t n;
};

// ---- plan 1: cell index + warp-aggregated per-cell histogram -------------------------------
__global__ void __launch_bounds__(kPlanThreads)
pool_cell_index_kernel(const __grid_constant__ PlanBatch pb) {
  const int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int32_t* target = nullptr;  // &count[gcell] of my plan, or null if invalid
  int32_t* rank_out = nullptr;
  bool tile_head = false;
  if (gi < pb.pt_total) {
    int j = 0;
    while (j + 1 < pb.n && gi >= pb.p[j + 1].pt_begin) ++j;
    const PlanDev& P = pb.p[j];
    const int64_t i = gi - P.pt_begin;
    const int32_t b = static_cast<int32_t>(i / P.N);
    const int32_t n = static_cast<int32_t>(i - static_cast<int64_t>(b) * P.N);
    const float* q = P.ind + b * P.ind_sb + n * P.ind_sn;
    // fp32 multiply then C-cast truncation toward zero (reference .cu:40)
    const float fh = __fmul_rn(q[0], P.sh);
    const float fw = __fmul_rn(q[P.ind_sd], P.sw);
    const long long ih = static_cast<long long>(fh);
    const long long iw = static_cast<long long>(fw);
    int32_t cell = -1;
    if (ih >= 0 && ih < P.H && iw >= 0 && iw < P.W) {
      cell = static_cast<int32_t>(ih) * P.W + static_cast<int32_t>(iw);
      target = P.count + (b * P.hw + cell);
    }
    P.cell[i] = cell;
    rank_out = P.rank + i;
    tile_head = (n & 63) == 0;  // runs never cross the 64-point tiles of the permute kernel
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
  if (rank_out != nullptr)
    *rank_out = target != nullptr ? (base + (lane - leader)) | (lane != leader ? kMergedPos : 0) : -1;
}

// ---- plan 2: give every occupied cell a segment of the sorted list ---------------------------
// Order between warps is irrelevant (max is order independent), so a warp scan plus one atomic on a
// global cursor replaces a device-wide prefix scan. thread = 4 cells; a warp never straddles two plans.
__global__ void __launch_bounds__(kPlanThreads)
pool_cell_alloc_kernel(const __grid_constant__ PlanBatch pb) {
  const int64_t gq = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gq >= pb.quad_total) return;  // whole warps only (quad ranges are warp aligned)
  int j = 0;
  while (j + 1 < pb.n && gq >= pb.p[j + 1].quad_begin) ++j;
  const PlanDev& P = pb.p[j];
  const int32_t c0 = static_cast<int32_t>(gq - P.quad_begin) * 4;
  const int lane = threadIdx.x & 31;
  int32_t c[4] = {0, 0, 0, 0};
  const bool vec = (c0 + 3 < P.cells) && ((reinterpret_cast<uintptr_t>(P.count) & 15) == 0);
  if (vec) {
    const int4 v = *reinterpret_cast<const int4*>(P.count + c0);
    c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
  } else {
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
  int32_t base = 0;
  if (lane == 31 && warp_total > 0) base = atomicAdd(cursor, warp_total);
  base = __shfl_sync(0xffffffffu, base, 31);
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
  const int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gi >= pb.pt_total) return;
  int j = 0;
  while (j + 1 < pb.n && gi >= pb.p[j + 1].pt_begin) ++j;
  const PlanDev& P = pb.p[j];
  const int64_t i = gi - P.pt_begin;
  const int32_t cell = P.cell[i];
  if (cell < 0) return;
  const int32_t b = static_cast<int32_t>(i / P.N);
  const int32_t n = static_cast<int32_t>(i - static_cast<int64_t>(b) * P.N);
  const int32_t gcell = b * P.hw + cell;
  const int32_t r = P.rank[i];
  const int32_t merged = r & kMergedPos;
  const int32_t pos = __ldg(P.start + gcell) + (r & kPosMask);
  P.rank[i] = pos | merged;  // from here on: sorted position (| kMergedPos), -1 if invalid
  P.sorted[pos] = make_int2(static_cast<int32_t>(static_cast<uint32_t>(n) | (merged ? kMergedN : 0u)), gcell);
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

// TMA variant of the permute (used when the channel rows are contiguous and 16-byte aligned, i.e.
// the (B,C,N,1)-contiguous tensor the reference produces, N % 4 == 0). Persistent CTAs walk the
// 64-point tiles; warp 0 streams the C channel rows of the NEXT tile into the other shared-memory
// buffer with cp.async.bulk (one 256-byte bulk copy per channel, completion counted on an mbarrier)
// while all warps write the rows of the current tile: no register staging, loads of tile t+1 overlap
// stores of tile t.
constexpr int kPermPitch = kPermPts + 4;  // floats; keeps every row 16-byte aligned, 4-way bank conflicts on read

__global__ void __launch_bounds__(kPermThreads)
pool_permute_tma_kernel(const float* __restrict__ feat, int32_t C, int32_t N, int32_t B, int64_t f_sb, int64_t f_sc,
                        const int32_t* __restrict__ pos, float* __restrict__ rows) {
  extern __shared__ __align__(128) float ptile[];  // [2][C][kPermPitch]
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int32_t s_pos[2][kPermPts];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int32_t tiles_per_b = (N + kPermPts - 1) / kPermPts;
  const int32_t ntiles = tiles_per_b * B;
  const int32_t buf_floats = C * kPermPitch;
  if (threadIdx.x == 0) {
    smos_mbar_init(&bar[0], 1);
    smos_mbar_init(&bar[1], 1);
    smos_fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](int32_t t, int buf) {
    const int32_t b = t / tiles_per_b;
    const int32_t n0 = (t - b * tiles_per_b) * kPermPts;
    const int32_t np = min(kPermPts, N - n0);
    if (wid == 0) {
      if (lane == 0) smos_mbar_expect_tx(&bar[buf], static_cast<uint32_t>(C) * np * 4u);
      __syncwarp();
      const float* src = feat + b * f_sb + n0;
      float* dst = ptile + buf * buf_floats;
      for (int32_t c = lane; c < C; c += 32)
        smos_bulk_g2s(dst + c * kPermPitch, src + static_cast<int64_t>(c) * f_sc, static_cast<uint32_t>(np) * 4u, &bar[buf]);
    } else if (wid == 1 || wid == 2) {
      const int32_t p = (wid - 1) * 32 + lane;
      s_pos[buf][p] = p < np ? __ldg(pos + static_cast<int64_t>(b) * N + n0 + p) : -1;
    }
  };

  int32_t t = blockIdx.x;
  if (t < ntiles) issue(t, 0);
  uint32_t phase[2] = {0u, 0u};
  int buf = 0;
  for (; t < ntiles; t += gridDim.x, buf ^= 1) {
    const int32_t tn = t + gridDim.x;
    if (tn < ntiles) issue(tn, buf ^ 1);  // the other buffer was released by the barrier below
    smos_mbar_wait(&bar[buf], phase[buf]);
    phase[buf] ^= 1u;
    __syncthreads();  // s_pos[buf] (written one iteration ago) is visible; tile bytes have landed
    const float* tile = ptile + buf * buf_floats;
    // thread = (channel, segment of the tile's points): it sweeps its points once, keeping a running max
    // over a run (first point + merged followers) and storing one row element per run. A run that starts in
    // my segment is followed to its end; leading merged points belong to the previous segment's sweep.
    {
      const int32_t nseg = max(1, kPermThreads / C);            // segments per tile (C <= 256 here)
      const int32_t seg_len = (kPermPts + nseg - 1) / nseg;
      for (int32_t item = threadIdx.x; item < C * nseg; item += kPermThreads) {
        const int32_t seg = item / C, c = item - seg * C;
        const float* tc = tile + c * kPermPitch;
        int32_t p = seg * seg_len;
        const int32_t seg_end = min(kPermPts, p + seg_len);
        while (p < seg_end && s_pos[buf][p] >= 0 && (s_pos[buf][p] & kMergedPos)) ++p;  // someone else's run
        while (p < seg_end) {
          const int32_t q = s_pos[buf][p];
          if (q < 0) { ++p; continue; }  // invalid point or past the end of the scan
          float v = tc[p];
          ++p;
          while (p < kPermPts && s_pos[buf][p] >= 0 && (s_pos[buf][p] & kMergedPos)) { v = fmaxf(v, tc[p]); ++p; }
          rows[static_cast<int64_t>(q) * C + c] = v;
        }
      }
    }
    smos_fence_proxy_async();  // order my generic-proxy reads before the async-proxy refill
    __syncthreads();           // everyone is done with `buf`: it may be refilled next iteration
  }
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
                   const int2* __restrict__ sorted, const int32_t* __restrict__ cursor, float* rows) {
  using V = typename VecT<VEC>::type;
  constexpr int kBatch = VEC == 4 ? 8 : 16;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int32_t total = *cursor;
  const int32_t p0 = (blockIdx.x * kReduceWarps + wib) * 32;
  if (p0 >= total) return;
  const int32_t cnt = min(32, total - p0);
  int2 e = make_int2(0, -1 - lane);  // distinct negative cells for lanes past the end
  if (lane < cnt) e = sorted[p0 + lane];
  const int32_t prev = __shfl_up_sync(0xffffffffu, e.y, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || e.y != prev) &
                         (cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u));
  // element offset of my point's row (c = 0)
  int64_t base_lane;
  bool absent = false;  // SORTED_ROWS: this position has no row of its own (merged into its run's first row)
  if (SORTED_ROWS) {
    absent = lane < cnt && (static_cast<uint32_t>(e.x) & kMergedN) != 0u;
    // read a row that is needed anyway instead of the unwritten one: the nearest row-owning lane at or before
    // me, else the first one after me (a 32-position window always contains one: runs are <= 32 long)
    const unsigned owners = __ballot_sync(0xffffffffu, lane < cnt && !absent);
    const unsigned before = owners & (0xffffffffu >> (31 - lane));
    const int src_lane = before ? 31 - __clz(before) : (owners ? __ffs(owners) - 1 : lane);
    base_lane = static_cast<int64_t>(p0 + src_lane) * C;
  } else {
    base_lane = (e.y >= 0 ? e.y / hw : 0) * f_sb +
                static_cast<int64_t>(static_cast<uint32_t>(e.x) & ~kMergedN) * f_sn;
  }
  const unsigned absent_mask = __ballot_sync(0xffffffffu, absent);
  const float* src = SORTED_ROWS ? rows : feat;

  for (int32_t c0 = 0; c0 < C; c0 += 32 * VEC) {
    const int32_t c = c0 + lane * VEC;
    const bool c_ok = c < C;  // C % VEC == 0 is guaranteed by the dispatcher
    const int32_t c_ld = c_ok ? c : 0;  // lanes past C load a valid address and discard the value
    V acc;
    int32_t piece_pos = p0;
    for (int32_t i0 = 0; i0 < cnt; i0 += kBatch) {
      V v[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        // UNCONDITIONAL loads (rows past the end re-read the last valid row): a predicated load makes
        // ptxas funnel every value through one temporary register, which serialises the batch
        const int32_t i = min(i0 + u, cnt - 1);
        const int64_t off = __shfl_sync(0xffffffffu, base_lane, i);
        v[u] = __ldg(reinterpret_cast<const V*>(src + off + c_ld));
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int32_t i = i0 + u;
        if (i < cnt && c_ok) {
          if (SORTED_ROWS && ((absent_mask >> i) & 1u)) v[u] = vneg_inf<VEC>();  // no row of its own
          if ((heads >> i) & 1u) {
            if (i > 0) *reinterpret_cast<V*>(rows + static_cast<int64_t>(piece_pos) * C + c) = acc;
            piece_pos = p0 + i;
            acc = v[u];
          } else {
            acc = vmax<VEC>(acc, v[u]);
          }
        }
      }
    }
    if (c_ok) *reinterpret_cast<V*>(rows + static_cast<int64_t>(piece_pos) * C + c) = acc;
  }
}

// ---- phase A2: fold the pieces of multi-piece cells into their first row ------------------------
// One warp per listed cell, lanes over channels, piece rows loaded kBatch at a time. After this every
// occupied cell's maxima sit in rows[start[cell]].
template <int VEC>
__global__ void __launch_bounds__(kReduceWarps * 32)
pool_combine_kernel(int32_t C, const int2* __restrict__ multi, const int32_t* __restrict__ cursor, float* rows) {
  using V = typename VecT<VEC>::type;
  constexpr int kBatch = VEC == 4 ? 8 : 16;
  const int lane = threadIdx.x & 31;
  const int32_t w = blockIdx.x * kReduceWarps + (threadIdx.x >> 5);
  if (w >= cursor[1]) return;
  const int2 sk = multi[w];
  const int32_t s = sk.x, end = sk.x + sk.y;
  const int32_t first_aligned = (s & ~31) + 32;
  const int32_t npieces = 1 + (end - first_aligned + 31) / 32;  // end > first_aligned by construction
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

// ---- phase B: dense writer --------------------------------------------------------------------
// thread = 4 adjacent cells (along W); it walks the channel groups (kCG channels each) of its CTA's
// channel range, so one read of count/start feeds up to C output stores. An occupied cell reads its
// one row (rows[start]); all row loads of a channel group are issued together. Empty cells store zeros.
// Every output element is written exactly once with 128-bit stores.
template <bool VEC4>
__global__ void __launch_bounds__(kWriteThreads)
pool_write_kernel(const float* __restrict__ rows, int32_t C, int32_t hw, int32_t groups_per_cta,
                  const int32_t* __restrict__ count, const int32_t* __restrict__ start,
                  float* __restrict__ out, int stream_out) {
  const int32_t b = blockIdx.z;
  const bool vec_rows = ((C & 7) == 0);
  constexpr int CPT = VEC4 ? 4 : 1;  // cells per thread
  const int32_t cell0 = (blockIdx.x * kWriteThreads + threadIdx.x) * CPT;
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
  if (any) {
    if (VEC4) {
      const int4 ss = __ldg(reinterpret_cast<const int4*>(start + g0));
      s[0] = ss.x; s[1 % CPT] = ss.y; s[2 % CPT] = ss.z; s[3 % CPT] = ss.w;
    } else {
      s[0] = __ldg(start + g0);
    }
  }
  const int32_t ngroups = (C + kCG - 1) / kCG;
  const int32_t cg_begin = blockIdx.y * groups_per_cta;
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
  return pool_layout(B, N, H, W).bytes;
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
  uintptr_t lo = ~uintptr_t(0), hi = 0;
  for (int32_t j = 0; j < n; ++j) {
    const smos_pool_plan_desc& d = descs_host[j];
    if (d.B <= 0 || d.N < 0 || d.H <= 0 || d.W <= 0 || d.plan == nullptr) return SMOS_EINVAL;
    if (d.B * d.N >= (int64_t(1) << 31) || d.B * static_cast<int64_t>(d.H) * d.W >= (int64_t(1) << 31))
      return SMOS_EUNSUPPORTED;
    if (d.N > 0 && d.pcds_ind == nullptr) return SMOS_EINVAL;
    const PoolLayout L = pool_layout(d.B, d.N, d.H, d.W);
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
    P.N = static_cast<int32_t>(d.N); P.H = d.H; P.W = d.W;
    P.hw = static_cast<int32_t>(L.hw); P.cells = static_cast<int32_t>(L.cells);
    P.sh = d.scale_h; P.sw = d.scale_w;
    P.pt_begin = pt; P.quad_begin = quad;
    pt += d.B * d.N;
    quad += smos_align_up((L.cells + 3) / 4, 32);
    lo = lo < reinterpret_cast<uintptr_t>(base) ? lo : reinterpret_cast<uintptr_t>(base);
    const uintptr_t end = reinterpret_cast<uintptr_t>(base) + static_cast<uintptr_t>(L.bytes);
    hi = hi > end ? hi : end;
  }
  pb.pt_total = pt;
  pb.quad_total = quad;
  cudaStream_t st = smos_stream(stream);
  // zero the per-cell counters (+ cursors). Plans carved out of one buffer are cleared with a single
  // memset over the whole span (a few MB) instead of one memset node per plan.
  int64_t sum_bytes = 0;
  for (int32_t j = 0; j < n; ++j) sum_bytes += pool_layout(descs_host[j].B, descs_host[j].N, descs_host[j].H, descs_host[j].W).bytes;
  if (n > 1 && static_cast<int64_t>(hi - lo) <= sum_bytes + 4096 * n) {
    cudaError_t e = cudaMemsetAsync(reinterpret_cast<void*>(lo), 0, hi - lo, st);
    if (e != cudaSuccess) return static_cast<int>(e);
  } else {
    for (int32_t j = 0; j < n; ++j) {
      cudaError_t e = cudaMemsetAsync(pb.p[j].count, 0, static_cast<size_t>(pb.p[j].cells + 4) * 4, st);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
  }
  if (pt > 0) pool_cell_index_kernel<<<smos_ceil_div(pt, kPlanThreads), kPlanThreads, 0, st>>>(pb);
  pool_cell_alloc_kernel<<<smos_ceil_div(quad, kPlanThreads), kPlanThreads, 0, st>>>(pb);
  if (pt > 0) pool_cell_scatter_kernel<<<smos_ceil_div(pt, kPlanThreads), kPlanThreads, 0, st>>>(pb);
  return smos_launch_status();
}

int smos_pool_plan_build(const float* pcds_ind, int64_t B, int64_t N, int64_t ind_sb, int64_t ind_sn,
                         int64_t ind_sd, int32_t H, int32_t W, float scale_h, float scale_w,
                         int64_t* voxel_max_idx, int64_t idx_batch_stride, void* plan, void* stream) {
  smos_pool_plan_desc d;
  d.pcds_ind = pcds_ind; d.B = B; d.N = N; d.ind_sb = ind_sb; d.ind_sn = ind_sn; d.ind_sd = ind_sd;
  d.H = H; d.W = W; d.scale_h = scale_h; d.scale_w = scale_w;
  d.voxel_max_idx = voxel_max_idx; d.idx_batch_stride = idx_batch_stride; d.plan = plan;
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
  const PoolLayout L = pool_layout(B, N, H, W);
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
  if (total > 0 && (stages & SMOS_POOL_STAGE_REDUCE)) {
    const int grid = smos_ceil_div(total, kReduceWarps * 32);
    const bool aligned = (reinterpret_cast<uintptr_t>(pcds_feat) & 15) == 0 && (f_sb & 3) == 0 && (f_sn & 3) == 0;
    const bool point_major = (f_sc == 1 && C > 1);
    if (!point_major) {
      // channel-major: permute into sorted rows, then reduce those rows in place
      static bool smem_opt_in[64] = {};
      int device = 0;
      cudaGetDevice(&device);
      const size_t smem = static_cast<size_t>(kPermPts) * (C + 1) * 4;
      if (smem > 200 * 1024) return SMOS_EUNSUPPORTED;
      if (device >= 0 && device < 64 && !smem_opt_in[device]) {
        cudaError_t e = cudaFuncSetAttribute(pool_permute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return static_cast<int>(e);
        smem_opt_in[device] = true;
      }
      const size_t smem_tma = static_cast<size_t>(2) * C * kPermPitch * 4;
      const bool tma_ok = f_sn == 1 && (N & 3) == 0 && (f_sb & 3) == 0 && (f_sc & 3) == 0 &&
                          (reinterpret_cast<uintptr_t>(pcds_feat) & 15) == 0 && smem_tma <= 96 * 1024;
      if (tma_ok) {
        static bool tma_opt_in[64] = {};
        if (device >= 0 && device < 64 && !tma_opt_in[device]) {
          cudaError_t e = cudaFuncSetAttribute(pool_permute_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
          if (e != cudaSuccess) return static_cast<int>(e);
          tma_opt_in[device] = true;
        }
        const int64_t ntiles = static_cast<int64_t>(smos_ceil_div(N, kPermPts)) * B;
        int ctas_per_sm = static_cast<int>((200 * 1024) / (smem_tma + 1024));
        if (ctas_per_sm > 6) ctas_per_sm = 6;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        int64_t pgrid = static_cast<int64_t>(SMOS_SM_COUNT) * ctas_per_sm;
        if (pgrid > ntiles) pgrid = ntiles;
        pool_permute_tma_kernel<<<static_cast<unsigned>(pgrid), kPermThreads, smem_tma, st>>>(
            pcds_feat, Ci, static_cast<int32_t>(N), static_cast<int32_t>(B), f_sb, f_sc, pos, rows);
      } else {
        dim3 pg(smos_ceil_div(N, kPermPts), static_cast<unsigned>(B));
        pool_permute_kernel<<<pg, kPermThreads, smem, st>>>(pcds_feat, Ci, static_cast<int32_t>(N), f_sb, f_sc, f_sn, pos, rows);
      }
      if ((C & 127) == 0) pool_reduce_kernel<4, true><<<grid, kReduceWarps * 32, 0, st>>>(rows, Ci, 0, 0, hw, sorted, cursor, rows);
      else if ((C & 63) == 0) pool_reduce_kernel<2, true><<<grid, kReduceWarps * 32, 0, st>>>(rows, Ci, 0, 0, hw, sorted, cursor, rows);
      else pool_reduce_kernel<1, true><<<grid, kReduceWarps * 32, 0, st>>>(rows, Ci, 0, 0, hw, sorted, cursor, rows);
    } else {
      if ((C & 127) == 0 && aligned) pool_reduce_kernel<4, false><<<grid, kReduceWarps * 32, 0, st>>>(pcds_feat, Ci, f_sb, f_sn, hw, sorted, cursor, rows);
      else if ((C & 63) == 0 && aligned) pool_reduce_kernel<2, false><<<grid, kReduceWarps * 32, 0, st>>>(pcds_feat, Ci, f_sb, f_sn, hw, sorted, cursor, rows);
      else pool_reduce_kernel<1, false><<<grid, kReduceWarps * 32, 0, st>>>(pcds_feat, Ci, f_sb, f_sn, hw, sorted, cursor, rows);
    }
  }
  if (total >= 32 && (stages & SMOS_POOL_STAGE_COMBINE)) {
    // fold multi-piece cells (<= total/32 of them; the exact number is only known on the device)
    const int2* multi = reinterpret_cast<const int2*>(base + L.off_multi);
    const int cgrid = smos_ceil_div(total / 32 + 1, kReduceWarps);
    if ((C & 127) == 0) pool_combine_kernel<4><<<cgrid, kReduceWarps * 32, 0, st>>>(Ci, multi, cursor, rows);
    else if ((C & 63) == 0) pool_combine_kernel<2><<<cgrid, kReduceWarps * 32, 0, st>>>(Ci, multi, cursor, rows);
    else pool_combine_kernel<1><<<cgrid, kReduceWarps * 32, 0, st>>>(Ci, multi, cursor, rows);
  }
  // outputs beyond L2 capacity are written with evict-first stores
  const int stream_out = (B * C * L.hw * 4 > (int64_t(96) << 20)) ? 1 : 0;
  const bool vec4 = ((hw & 3) == 0) && ((reinterpret_cast<uintptr_t>(voxel_out) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(count) & 15) == 0);
  const int cpt = vec4 ? 4 : 1;
  const int gx = smos_ceil_div(smos_ceil_div(hw, cpt), kWriteThreads);
  // split the channel groups over blockIdx.y only as far as needed to cover the SMs ~3x
  const int32_t ngroups = (Ci + kCG - 1) / kCG;
  int32_t groups_per_cta = ngroups;
  while (groups_per_cta > 1 && static_cast<int64_t>(gx) * B * ((ngroups + groups_per_cta - 1) / groups_per_cta) < 3 * SMOS_SM_COUNT)
    groups_per_cta = (groups_per_cta + 1) / 2;
  dim3 grid(gx, (ngroups + groups_per_cta - 1) / groups_per_cta, static_cast<unsigned>(B));
  if (stages & SMOS_POOL_STAGE_WRITE) {
    if (vec4)
      pool_write_kernel<true><<<grid, kWriteThreads, 0, st>>>(rows, Ci, hw, groups_per_cta, count, start, voxel_out, stream_out);
    else
      pool_write_kernel<false><<<grid, kWriteThreads, 0, st>>>(rows, Ci, hw, groups_per_cta, count, start, voxel_out, stream_out);
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
  const PoolLayout L = pool_layout(B, N, H, W);
  const int32_t* cell = reinterpret_cast<const int32_t*>(static_cast<const char*>(plan) + L.off_cell);
  const int64_t total = B * C * N;
  const int fast_n = (f_sc == 1 && C > 1) ? 0 : 1;
  pool_backward_kernel<<<smos_ceil_div(total, 256), 256, 0, smos_stream(stream)>>>(
      pcds_feat, static_cast<int32_t>(C), static_cast<int32_t>(N), total, f_sb, f_sc, f_sn, L.hw, cell, voxel_out,
      grad_voxel_out, grad_feat, g_sb, g_sc, g_sn, fast_n);
  return smos_launch_status();
}

}  // extern "C"
