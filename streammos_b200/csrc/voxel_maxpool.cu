// VoxelMaxPool for sm_100a: output-stationary scatter-max.
//
// The reference (deep_point/src/point_deep_cuda_kernel.cu:24-99) zero-fills the
// dense grid, stores every point feature once (racy init), then issues one
// CAS-loop atomic per (point, channel) into global memory. Here the points are
// bucketed by OUTPUT TILE first (the "plan"); one CTA then owns a tile of the
// grid for a chunk of channels in shared memory, reduces its points with native
// shared-memory integer atomics, and writes every output element exactly once
// with 128-bit stores — the zero-fill, the init pass and all global atomics
// disappear, so HBM traffic is the algorithmic minimum (features in, grid out).
#include "common.cuh"

namespace {

constexpr int kPlanThreads = 256;
constexpr int kPoolThreads = 256;
constexpr uint32_t kNegInfBits = 0xff800000u;  // "empty cell" marker in the smem tile

struct PoolLayout {
  int32_t th, tw, ntx, nty, nt;
  int64_t off_cell, off_rank, off_count, off_start, off_sorted, bytes;
};

// Tile shape depends on the grid only, so plan and pooling always agree.
// Large grids: 1024-cell tiles (4 KB / channel); small grids: 256-cell tiles so
// that B x tiles x channel-chunks still covers the 148 SMs several times.
void pick_tile(int32_t H, int32_t W, int32_t* th, int32_t* tw) {
  const int64_t cells = static_cast<int64_t>(H) * W;
  const int32_t target = cells > 65536 ? 1024 : 256;
  int32_t w = 64;
  if (target == 256) w = 32;
  while (w > W && w > 4) w >>= 1;
  int32_t h = target / w;
  while (h > H && h > 1) h >>= 1;
  *th = h;
  *tw = w;
}

PoolLayout pool_layout(int64_t B, int64_t N, int32_t H, int32_t W) {
  PoolLayout L;
  pick_tile(H, W, &L.th, &L.tw);
  L.ntx = (W + L.tw - 1) / L.tw;
  L.nty = (H + L.th - 1) / L.th;
  L.nt = L.ntx * L.nty;
  const int64_t bn = B * N;
  const int64_t btiles = B * L.nt;
  int64_t off = 0;
  L.off_cell = off;   off += smos_align_up(bn * 4, 256);
  L.off_rank = off;   off += smos_align_up(bn * 4, 256);
  L.off_count = off;  off += smos_align_up(btiles * 4, 256);
  L.off_start = off;  off += smos_align_up((btiles + 1) * 4, 256);
  L.off_sorted = off; off += smos_align_up(bn * 8, 256);
  L.bytes = off;
  return L;
}

// ---- plan kernel 1: cell index + warp-aggregated tile histogram ------------------------
__global__ void __launch_bounds__(kPlanThreads)
pool_cell_index_kernel(const float* __restrict__ ind, int64_t total, int32_t N,
                       int64_t ind_sb, int64_t ind_sn, int64_t ind_sd,
                       int32_t H, int32_t W, float scale_h, float scale_w,
                       int32_t th, int32_t tw, int32_t ntx, int32_t nt,
                       int64_t* __restrict__ voxel_max_idx, int64_t idx_batch_stride,
                       int32_t* __restrict__ cell_out, int32_t* __restrict__ rank_out,
                       int32_t* __restrict__ tile_count) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int32_t gtile = -1;
  int32_t cell = -1;
  int32_t b = 0;
  if (i < total) {
    b = static_cast<int32_t>(i / N);
    const int32_t n = static_cast<int32_t>(i - static_cast<int64_t>(b) * N);
    const float* p = ind + b * ind_sb + n * ind_sn;
    // fp32 multiply then C-cast truncation (reference .cu:40)
    const float fh = __fmul_rn(p[0], scale_h);
    const float fw = __fmul_rn(p[ind_sd], scale_w);
    const long long ih = static_cast<long long>(fh);
    const long long iw = static_cast<long long>(fw);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      const int32_t h = static_cast<int32_t>(ih), w = static_cast<int32_t>(iw);
      cell = h * W + w;
      gtile = b * nt + (h / th) * ntx + (w / tw);
    }
    cell_out[i] = cell;
    if (voxel_max_idx != nullptr)
      voxel_max_idx[i] = cell >= 0 ? static_cast<int64_t>(b) * idx_batch_stride + cell : -1;
  }
  // one atomic per (warp, tile): lanes that hit the same tile elect a leader
  const unsigned peers = __match_any_sync(0xffffffffu, gtile);
  const int leader = __ffs(peers) - 1;
  const int lane = threadIdx.x & 31;
  int32_t base = 0;
  if (lane == leader && gtile >= 0) base = atomicAdd(&tile_count[gtile], __popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (i < total) rank_out[i] = gtile >= 0 ? base + __popc(peers & smos_lanemask_lt()) : -1;
}

// ---- plan kernel 2: exclusive scan of the tile counts (single CTA) -----------------------
__global__ void __launch_bounds__(1024)
pool_tile_scan_kernel(const int32_t* __restrict__ count, int32_t* __restrict__ start, int32_t total) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int32_t base = 0; base < total; base += 1024) {
    const int32_t i = base + threadIdx.x;
    const int32_t v = i < total ? count[i] : 0;
    int32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int32_t s = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_sums[lane] = s;  // inclusive
    }
    __syncthreads();
    const int32_t warp_off = wid > 0 ? warp_sums[wid - 1] : 0;
    const int32_t c = carry;
    if (i < total) start[i] = c + warp_off + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + warp_off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[total] = carry;
}

// ---- plan kernel 3: place each valid point in its tile's segment ---------------------------
__global__ void __launch_bounds__(kPlanThreads)
pool_tile_scatter_kernel(const int32_t* __restrict__ cell_in, const int32_t* __restrict__ rank_in,
                         const int32_t* __restrict__ start, int64_t total, int32_t N, int32_t W,
                         int32_t th, int32_t tw, int32_t ntx, int32_t nt,
                         int2* __restrict__ sorted) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int32_t cell = cell_in[i];
  if (cell < 0) return;
  const int32_t b = static_cast<int32_t>(i / N);
  const int32_t n = static_cast<int32_t>(i - static_cast<int64_t>(b) * N);
  const int32_t h = cell / W, w = cell - h * W;
  const int32_t ty = h / th, tx = w / tw;
  const int32_t lc = (h - ty * th) * tw + (w - tx * tw);
  const int32_t pos = start[b * nt + ty * ntx + tx] + rank_in[i];
  sorted[pos] = make_int2(n, lc);
}

// ---- float max through native integer shared-memory atomics -------------------------------
// Tile initialised to -inf bits. For v >= +0 a signed max orders correctly and beats any
// negative pattern; for v < 0 an unsigned min orders negatives and never displaces a
// non-negative value.
__device__ __forceinline__ void smem_fmax(uint32_t* addr, float v) {
  const uint32_t bits = __float_as_uint(v);
  if (static_cast<int32_t>(bits) >= 0)
    atomicMax(reinterpret_cast<int32_t*>(addr), static_cast<int32_t>(bits));
  else
    atomicMin(addr, bits);
}

__device__ __forceinline__ float decode_cell(uint32_t bits) {
  return bits == kNegInfBits ? 0.0f : __uint_as_float(bits);
}

// One CTA = (batch b, output tile, chunk of CC channels).
// POINT_MAJOR: feature rows are contiguous per point (f_sc == 1) -> lanes run over channels;
// otherwise (channel-major, f_sn == 1 typical) lanes run over the tile's points.
template <bool POINT_MAJOR>
__global__ void __launch_bounds__(kPoolThreads)
pool_forward_kernel(const float* __restrict__ feat, int32_t C, int64_t f_sb, int64_t f_sc, int64_t f_sn,
                    int32_t H, int32_t W, int32_t th, int32_t tw, int32_t ntx, int32_t nt,
                    int32_t CC, int32_t nchunk, const int32_t* __restrict__ tile_start,
                    const int2* __restrict__ sorted, float* __restrict__ out, int stream_out) {
  extern __shared__ uint32_t tile[];
  const int32_t tcells = th * tw;
  const int32_t chunk = blockIdx.x % nchunk;
  const int32_t gt = blockIdx.x / nchunk;  // b * nt + tile
  const int32_t b = gt / nt;
  const int32_t t = gt - b * nt;
  const int32_t ty = t / ntx, tx = t - ty * ntx;
  const int32_t c0 = chunk * CC;
  const int32_t cc = min(CC, C - c0);
  const int32_t p0 = tile_start[gt], p1 = tile_start[gt + 1];
  // smem addressing: channel-major [c][lc]   or point-major [lc][CC+1] (conflict-free both ways)
  const int32_t s_c = POINT_MAJOR ? 1 : tcells;
  const int32_t s_l = POINT_MAJOR ? (CC + 1) : 1;

  if (p1 > p0) {
    const int32_t words = POINT_MAJOR ? tcells * (CC + 1) : tcells * CC;
    for (int32_t i = threadIdx.x; i < words; i += kPoolThreads) tile[i] = kNegInfBits;
    __syncthreads();
    const float* fb = feat + b * f_sb + static_cast<int64_t>(c0) * f_sc;
    if (POINT_MAJOR) {
      // CC is a power of two <= 32: a warp covers 32/CC points x CC channels per step
      const int32_t lane = threadIdx.x & 31;
      const int32_t c = lane & (CC - 1);
      const int32_t ppw = 32 / CC;
      const int32_t sub = lane / CC;
      const int32_t nsteps = (kPoolThreads / 32) * ppw;
      for (int32_t p = p0 + (threadIdx.x >> 5) * ppw + sub; p < p1; p += nsteps) {
        const int2 e = sorted[p];
        if (c < cc) {
          const float v = __ldg(fb + static_cast<int64_t>(e.x) * f_sn + c);
          smem_fmax(&tile[e.y * s_l + c], v);
        }
      }
    } else {
      const int32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
      constexpr int32_t nw = kPoolThreads / 32;
      for (int32_t p = p0 + lane; p < p1; p += 32) {
        const int2 e = sorted[p];
        const float* fp = fb + static_cast<int64_t>(e.x) * f_sn;
        for (int32_t c = wid; c < cc; c += nw) {
          const float v = __ldg(fp + static_cast<int64_t>(c) * f_sc);
          smem_fmax(&tile[c * s_c + e.y], v);
        }
      }
    }
    __syncthreads();
  }

  // write-out: every element of the tile once; rows of tw contiguous floats
  const int32_t h0 = ty * th, w0 = tx * tw;
  float* ob = out + (static_cast<int64_t>(b) * C + c0) * H * W;
  const bool vec = ((W & 3) == 0) && ((tw & 3) == 0);
  if (vec) {
    const int32_t q = tcells >> 2;
    for (int32_t i = threadIdx.x; i < cc * q; i += kPoolThreads) {
      const int32_t c = i / q;
      const int32_t lc = (i - c * q) << 2;
      const int32_t r = lc / tw, col = lc - r * tw;
      const int32_t h = h0 + r, w = w0 + col;
      if (h < H && w < W) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p1 > p0) {
          const uint32_t* s = tile + c * s_c + lc * s_l;
          v.x = decode_cell(s[0]);
          v.y = decode_cell(s[s_l]);
          v.z = decode_cell(s[2 * s_l]);
          v.w = decode_cell(s[3 * s_l]);
        }
        float* dst = ob + (static_cast<int64_t>(c) * H + h) * W + w;
        if (stream_out) smos_st_cs_f4(dst, v);
        else *reinterpret_cast<float4*>(dst) = v;
      }
    }
  } else {
    for (int32_t i = threadIdx.x; i < cc * tcells; i += kPoolThreads) {
      const int32_t c = i / tcells;
      const int32_t lc = i - c * tcells;
      const int32_t r = lc / tw, col = lc - r * tw;
      const int32_t h = h0 + r, w = w0 + col;
      if (h < H && w < W) {
        float v = 0.f;
        if (p1 > p0) v = decode_cell(tile[c * s_c + lc * s_l]);
        ob[(static_cast<int64_t>(c) * H + h) * W + w] = v;
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

int32_t pick_chan_chunk(int64_t B, int32_t C, int32_t nt, int32_t tcells, bool point_major) {
  // Keep the smem tile <= 64 KB (>= 3 CTAs / SM) and make enough CTAs to cover the SMs ~4x.
  int32_t cc = 32;
  while (cc > 1 && static_cast<int64_t>(cc + (point_major ? 1 : 0)) * tcells * 4 > 65536) cc >>= 1;
  while (cc > 8 && B * nt * ((C + cc - 1) / cc) < 4 * SMOS_SM_COUNT) cc >>= 1;
  while (cc > C && cc > 1) cc >>= 1;
  if (cc < 1) cc = 1;
  return cc;
}

}  // namespace

extern "C" {

int smos_pool_tile_shape(int32_t B, int32_t C, int32_t H, int32_t W, int32_t* tile_h_host,
                         int32_t* tile_w_host, int32_t* chan_chunk_host) {
  if (H <= 0 || W <= 0 || C <= 0 || B <= 0) return SMOS_EINVAL;
  int32_t th, tw;
  pick_tile(H, W, &th, &tw);
  if (tile_h_host) *tile_h_host = th;
  if (tile_w_host) *tile_w_host = tw;
  if (chan_chunk_host) {
    const int32_t nt = ((W + tw - 1) / tw) * ((H + th - 1) / th);
    *chan_chunk_host = pick_chan_chunk(B, C, nt, th * tw, false);
  }
  return SMOS_OK;
}

int64_t smos_pool_plan_bytes(int64_t B, int64_t N, int32_t H, int32_t W) {
  if (B <= 0 || N < 0 || H <= 0 || W <= 0) return SMOS_EINVAL;
  return pool_layout(B, N, H, W).bytes;
}

int smos_pool_plan_build(const float* pcds_ind, int64_t B, int64_t N, int64_t ind_sb, int64_t ind_sn,
                         int64_t ind_sd, int32_t H, int32_t W, float scale_h, float scale_w,
                         int64_t* voxel_max_idx, int64_t idx_batch_stride, void* plan, void* stream) {
  if (B <= 0 || N < 0 || H <= 0 || W <= 0 || plan == nullptr) return SMOS_EINVAL;
  if (B * N >= (int64_t(1) << 31) || static_cast<int64_t>(H) * W >= (int64_t(1) << 31)) return SMOS_EUNSUPPORTED;
  if (N > 0 && pcds_ind == nullptr) return SMOS_EINVAL;
  const PoolLayout L = pool_layout(B, N, H, W);
  char* base = static_cast<char*>(plan);
  int32_t* cell = reinterpret_cast<int32_t*>(base + L.off_cell);
  int32_t* rank = reinterpret_cast<int32_t*>(base + L.off_rank);
  int32_t* count = reinterpret_cast<int32_t*>(base + L.off_count);
  int32_t* start = reinterpret_cast<int32_t*>(base + L.off_start);
  int2* sorted = reinterpret_cast<int2*>(base + L.off_sorted);
  cudaStream_t st = smos_stream(stream);
  const int64_t total = B * N;
  const int32_t btiles = static_cast<int32_t>(B * L.nt);
  cudaError_t e = cudaMemsetAsync(count, 0, static_cast<size_t>(btiles) * 4, st);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (total > 0) {
    pool_cell_index_kernel<<<smos_ceil_div(total, kPlanThreads), kPlanThreads, 0, st>>>(
        pcds_ind, total, static_cast<int32_t>(N), ind_sb, ind_sn, ind_sd, H, W, scale_h, scale_w, L.th, L.tw,
        L.ntx, L.nt, voxel_max_idx, idx_batch_stride, cell, rank, count);
  }
  pool_tile_scan_kernel<<<1, 1024, 0, st>>>(count, start, btiles);
  if (total > 0) {
    pool_tile_scatter_kernel<<<smos_ceil_div(total, kPlanThreads), kPlanThreads, 0, st>>>(
        cell, rank, start, total, static_cast<int32_t>(N), W, L.th, L.tw, L.ntx, L.nt, sorted);
  }
  return smos_launch_status();
}

int smos_voxel_maxpool_forward(const float* pcds_feat, int64_t B, int64_t C, int64_t N, int64_t f_sb,
                               int64_t f_sc, int64_t f_sn, int32_t H, int32_t W, const void* plan,
                               float* voxel_out, void* stream) {
  if (B <= 0 || C <= 0 || N < 0 || H <= 0 || W <= 0 || plan == nullptr || voxel_out == nullptr) return SMOS_EINVAL;
  if (N > 0 && pcds_feat == nullptr) return SMOS_EINVAL;
  if (B * N >= (int64_t(1) << 31) || C >= (1 << 20)) return SMOS_EUNSUPPORTED;
  const PoolLayout L = pool_layout(B, N, H, W);
  const char* base = static_cast<const char*>(plan);
  const int32_t* start = reinterpret_cast<const int32_t*>(base + L.off_start);
  const int2* sorted = reinterpret_cast<const int2*>(base + L.off_sorted);
  const bool point_major = (f_sc == 1 && C > 1);
  const int32_t tcells = L.th * L.tw;
  const int32_t CC = pick_chan_chunk(B, static_cast<int32_t>(C), L.nt, tcells, point_major);
  const int32_t nchunk = (static_cast<int32_t>(C) + CC - 1) / CC;
  const int64_t grid = B * L.nt * nchunk;
  if (grid >= (int64_t(1) << 31)) return SMOS_EUNSUPPORTED;
  const size_t smem = static_cast<size_t>(CC + (point_major ? 1 : 0)) * tcells * 4;
  // outputs beyond L2 capacity are written with evict-first stores
  const int stream_out = (B * C * static_cast<int64_t>(H) * W * 4 > (int64_t(96) << 20)) ? 1 : 0;
  cudaStream_t st = smos_stream(stream);
  // opt in to > 48 KB dynamic shared memory once per device (not a stream operation; kept out of the
  // per-launch path so that launches can be captured into CUDA graphs)
  static bool smem_opt_in[64] = {};
  int device = 0;
  cudaGetDevice(&device);
  if (device >= 0 && device < 64 && !smem_opt_in[device]) {
    cudaError_t e = cudaFuncSetAttribute(pool_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(pool_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    smem_opt_in[device] = true;
  }
  if (point_major) {
    pool_forward_kernel<true><<<static_cast<unsigned>(grid), kPoolThreads, smem, st>>>(
        pcds_feat, static_cast<int32_t>(C), f_sb, f_sc, f_sn, H, W, L.th, L.tw, L.ntx, L.nt, CC, nchunk, start,
        sorted, voxel_out, stream_out);
  } else {
    pool_forward_kernel<false><<<static_cast<unsigned>(grid), kPoolThreads, smem, st>>>(
        pcds_feat, static_cast<int32_t>(C), f_sb, f_sc, f_sn, H, W, L.th, L.tw, L.ntx, L.nt, CC, nchunk, start,
        sorted, voxel_out, stream_out);
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
      pcds_feat, static_cast<int32_t>(C), static_cast<int32_t>(N), total, f_sb, f_sc, f_sn,
      static_cast<int64_t>(H) * W, cell, voxel_out, grad_voxel_out, grad_feat, g_sb, g_sc, g_sn, fast_n);
  return smos_launch_status();
}

}  // extern "C"
