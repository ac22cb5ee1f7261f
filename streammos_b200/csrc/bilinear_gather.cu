// BilinearSample for sm_100a: bilinear gather of 2-D grid features back to points.
//
// Replaces networks/backbone.py:458-475, i.e. four elementwise prep kernels + torch.stack +
// F.grid_sample(bilinear, zeros, align_corners=True). One kernel: the sampling position and
// the four tap weights are computed once per point and reused for every channel; the
// reference's float sequence (normalise then un-normalise) is replayed with explicitly
// rounded fp32 operations so results stay within 1e-5 of grid_sample.
#include "common.cuh"
#include "gather_taps.cuh"

namespace {

#ifndef SMOS_GATHER_CPT
#define SMOS_GATHER_CPT 8
#endif
#ifndef SMOS_GATHER_PLANAR_MIN_CTAS
#define SMOS_GATHER_PLANAR_MIN_CTAS 14  // <= 72 registers: room for the 32 loads of a step in flight
#endif
constexpr int kCPT = SMOS_GATHER_CPT;  // channels per thread-step in the planar kernel (4 or 8)

// Per-lane sampling state: slot `slot` of the launch (ORDERED: entry of a pooling plan's `sorted` list over all B*N
// points; else point `slot` of batch b). Each lane computes (or, with TAPS, loads) the state of ITS point: lanes are
// points in the planar kernel, so nothing needs to be shared between threads.
template <bool ORDERED, bool TAPS>
__device__ __forceinline__ TapsS lane_taps(int64_t slot, int32_t b, const float* __restrict__ coord, int32_t N,
                                           int64_t co_sb, int64_t co_sn, int64_t co_sd, float sh, float sw, int32_t H,
                                           int32_t W, const int2* __restrict__ order, int32_t order_hw, int64_t order_len,
                                           const TapsS* __restrict__ taps) {
  if (TAPS) {
    TapsS r;
    const bool live = slot < order_len;
    const uint4* src = reinterpret_cast<const uint4*>(taps + (live ? slot : order_len - 1));
    uint4* dst = reinterpret_cast<uint4*>(&r);
    dst[0] = __ldg(src); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2);
    if (!live) r.n = -1;
    return r;
  }
  int32_t n;
  bool live;
  if (ORDERED) {
    live = slot < order_len;
    const int2 e = __ldg(order + (live ? slot : order_len - 1));
    n = static_cast<int32_t>(static_cast<uint32_t>(e.x) & 0x7fffffffu);  // bit 31: the plan's run-merge mark
    // out-of-grid entries carry -1 - b; a single batch (order_len == N, the streaming case) needs no division — the
    // quotient sat on the chain order entry -> coordinates of every warp
    b = e.y >= 0 ? (order_len == static_cast<int64_t>(N) ? 0 : e.y / order_hw) : -1 - e.y;
  } else {
    live = slot < N;
    n = static_cast<int32_t>(live ? slot : N - 1);
  }
  const float* cp = coord + b * co_sb + static_cast<int64_t>(n) * co_sn;
  return make_taps_record(__ldg(cp), __ldg(cp + co_sd), sh, sw, H, W, live ? n : -1, b);
}

// Planar (NCHW-like: arbitrary channel stride, contiguous H x W planes up to gr_sh / gr_sw) grid.
// One WARP = 32 consecutive slots x all channels, independent of every other warp (no block barrier anywhere): lanes
// are points, the warp walks the channel groups of kCPT channels. A scan is then ONE wave of ~N/32 warps spread over
// the whole GPU — every warp runs its latency chain (order entry -> coordinates -> sampling state -> tap loads ->
// row stores) once, concurrently with all others, instead of 2.5 waves of CTAs each paying the chain plus two block
// barriers (the round-1 kernel: 4 warps sharing 32 points through shared memory, 17 us for 32 ch @ 256^2).
// All 4*kCPT tap loads of a step are UNCONDITIONAL (out-of-image taps read a clamped, valid pixel and are zeroed by a
// select afterwards): predicated loads get serialised through one temporary register by ptxas. For point-major
// outputs the warp parks its 32 x C results in its own slice of shared memory ([point][C + 4]: conflict-free 128-bit
// stores) and the rows leave as whole 128-byte lines.
constexpr int kGatherWarps = 2;  // independent warps per CTA

template <bool DENSE, bool ROWS_OUT, bool ORDERED, bool TAPS>
__global__ void __launch_bounds__(kGatherWarps * 32, SMOS_GATHER_PLANAR_MIN_CTAS)
gather_forward_planar_kernel(const float* __restrict__ grid, int32_t C, int32_t H, int32_t W,
                             int64_t gr_sb, int64_t gr_sc, int64_t gr_sh, int64_t gr_sw,
                             const float* __restrict__ coord, int32_t N,
                             int64_t co_sb, int64_t co_sn, int64_t co_sd, float sh, float sw,
                             float* __restrict__ out, int64_t o_sb, int64_t o_sc, int64_t o_sn,
                             const int2* __restrict__ order, int32_t order_hw, int64_t order_len,
                             const TapsS* __restrict__ taps, int32_t groups_per_warp) {
  SMOS_PDL_PROLOGUE();
  extern __shared__ __align__(16) float s_out_all[];  // ROWS_OUT: [kGatherWarps][32][channels of one warp + 4]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t slot0 = (static_cast<int64_t>(blockIdx.x) * kGatherWarps + wid) * 32;
  if (slot0 >= (ORDERED || TAPS ? order_len : static_cast<int64_t>(N))) return;  // whole warp past the end
  const TapsS t = lane_taps<ORDERED, TAPS>(slot0 + lane, blockIdx.z, coord, N, co_sb, co_sn, co_sd, sh, sw, H, W, order,
                                           order_hw, order_len, taps);
  const int32_t b = t.b;
  const int32_t n = t.n;  // < 0: slot without a point
  // channel groups (kCPT channels each) of this warp: blockIdx.y splits the channels when there are few points
  const int32_t ngroups = (C + kCPT - 1) / kCPT;
  const int32_t cg_begin = blockIdx.y * groups_per_warp;
  const int32_t cg_end = min(ngroups, cg_begin + groups_per_warp);
  const float* g = grid + b * gr_sb;
  // the staging slice holds only the channels THIS warp produces (a warp of a split launch used to reserve rows of
  // all C channels: at 64 channels that capped the SM at 13 CTAs)
  const int32_t c_lo0 = cg_begin * kCPT;
  const int32_t ld_s = groups_per_warp * kCPT + 4;
  float* s_out = s_out_all + static_cast<size_t>(wid) * 32 * ld_s;
  const bool vec_ok = (o_sc == 1) && ((o_sn & 3) == 0) && ((o_sb & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  float* const orow = out + b * o_sb + static_cast<int64_t>(n >= 0 ? n : 0) * o_sn;
  auto emit = [&](int32_t c0, int32_t nch, const float (&acc)[kCPT]) {
    if (ROWS_OUT) {
      float* so = s_out + lane * ld_s + (c0 - c_lo0);
#pragma unroll
      for (int k4 = 0; k4 < kCPT; k4 += 4)
        *reinterpret_cast<float4*>(so + k4) = make_float4(acc[k4], acc[k4 + 1], acc[k4 + 2], acc[k4 + 3]);
    } else if (n >= 0) {
      float* o = orow + static_cast<int64_t>(c0) * o_sc;
      if (vec_ok && nch == kCPT) {
#pragma unroll
        for (int k4 = 0; k4 < kCPT; k4 += 4)
          *reinterpret_cast<float4*>(o + k4) = make_float4(acc[k4], acc[k4 + 1], acc[k4 + 2], acc[k4 + 3]);
      } else {
#pragma unroll
        for (int k = 0; k < kCPT; ++k)
          if (k < nch) o[static_cast<int64_t>(k) * o_sc] = acc[k];
      }
    }
  };
  // Almost every warp has all four taps of all its points inside the image (only border pixels and the padded points
  // at the tail of the cell order do not): then NE = NW + 1 pixel and SE = SW + 1 pixel, no tap needs zeroing, and a
  // channel costs two address computations, four loads and four FMAs.
  const bool all_in = DENSE && ((C & (kCPT - 1)) == 0) && __all_sync(0xffffffffu, t.in_mask == 0xfu);
  if (all_in) {
    const float* pn = g + t.o_nw;
    const float* ps = g + t.o_sw;
#pragma unroll 1
    for (int32_t cg = cg_begin; cg < cg_end; ++cg) {
      const int32_t c0 = cg * kCPT;
      const float* an = pn + static_cast<int64_t>(c0) * gr_sc;
      const float* as = ps + static_cast<int64_t>(c0) * gr_sc;
      float v[kCPT][4];
#pragma unroll
      for (int k = 0; k < kCPT; ++k) {
        v[k][0] = __ldg(an + k * gr_sc);
        v[k][1] = __ldg(an + k * gr_sc + 1);
        v[k][2] = __ldg(as + k * gr_sc);
        v[k][3] = __ldg(as + k * gr_sc + 1);
      }
      float acc[kCPT];
#pragma unroll
      for (int k = 0; k < kCPT; ++k)
        acc[k] = fmaf(v[k][3], t.w_se, fmaf(v[k][2], t.w_sw, fmaf(v[k][1], t.w_ne, fmaf(v[k][0], t.w_nw, 0.f))));
      emit(c0, kCPT, acc);
    }
  } else {
    // general path: any strides, taps clamped into the image and zeroed by a select (value AND weight: a clamped
    // pixel may hold inf / NaN). All loads unconditional: predicated ones get serialised by ptxas.
    int64_t q_nw = t.o_nw, q_ne = t.o_ne, q_sw = t.o_sw, q_se = t.o_se;
    if (!DENSE) {
      q_nw = static_cast<int64_t>(t.o_nw / W) * gr_sh + static_cast<int64_t>(t.o_nw % W) * gr_sw;
      q_ne = static_cast<int64_t>(t.o_ne / W) * gr_sh + static_cast<int64_t>(t.o_ne % W) * gr_sw;
      q_sw = static_cast<int64_t>(t.o_sw / W) * gr_sh + static_cast<int64_t>(t.o_sw % W) * gr_sw;
      q_se = static_cast<int64_t>(t.o_se / W) * gr_sh + static_cast<int64_t>(t.o_se % W) * gr_sw;
    }
    const bool i_nw = t.in_mask & 1u, i_ne = t.in_mask & 2u, i_sw = t.in_mask & 4u, i_se = t.in_mask & 8u;
#pragma unroll 1
    for (int32_t cg = cg_begin; cg < cg_end; ++cg) {
      const int32_t c0 = cg * kCPT;
      const int32_t nch = min(kCPT, C - c0);
      float v[kCPT][4];
#pragma unroll
      for (int k = 0; k < kCPT; ++k) {
        const float* gc = g + static_cast<int64_t>(min(c0 + k, C - 1)) * gr_sc;  // a partial group re-reads its last channel
        v[k][0] = __ldg(gc + q_nw);
        v[k][1] = __ldg(gc + q_ne);
        v[k][2] = __ldg(gc + q_sw);
        v[k][3] = __ldg(gc + q_se);
      }
      float acc[kCPT];
#pragma unroll
      for (int k = 0; k < kCPT; ++k) {
        const float a = i_nw ? v[k][0] : 0.f, bq = i_ne ? v[k][1] : 0.f;
        const float c = i_sw ? v[k][2] : 0.f, d = i_se ? v[k][3] : 0.f;
        acc[k] = fmaf(d, t.w_se, fmaf(c, t.w_sw, fmaf(bq, t.w_ne, fmaf(a, t.w_nw, 0.f))));
      }
      emit(c0, nch, acc);
    }
  }
  if (ROWS_OUT) {
    __syncwarp();
    // C % 8 == 0 here. This warp's channel range [c_lo, c_hi) of every row is q float4 pieces; lane l moves piece
    // i % q of row i / q for i = l, l + 32, ...: each store instruction writes whole 32-byte sectors of 32 / q rows
    const int32_t c_lo = cg_begin * kCPT, q = ((cg_end - cg_begin) * kCPT) >> 2;
    const int64_t my_row = n >= 0 ? (b * o_sb + static_cast<int64_t>(n) * o_sn) : -1;
    const bool pow2 = (q & (q - 1)) == 0;
    const int32_t shq = 31 - __clz(q);
    if (pow2 && q <= 32) {
      // q pieces per row, 32 / q rows per step: a lane keeps its piece index, only the row advances — one shuffle, one
      // shared load and one store per step on pointers that move by constants (the generic loop below recomputed row,
      // piece and both addresses for every piece: a fifth of the kernel's instructions)
      const int32_t rpi = 32 >> shq, pr = lane >> shq, j4 = (lane & (q - 1)) << 2;
      const float* sp = s_out + pr * ld_s + j4;
      float* gp = out + c_lo + j4;
      const int32_t sstep = rpi * ld_s;
#pragma unroll 4
      for (int32_t it = 0; it < q; ++it) {
        const int64_t rb = __shfl_sync(0xffffffffu, my_row, it * rpi + pr);
        const float4 r = *reinterpret_cast<const float4*>(sp + it * sstep);
        if (rb >= 0) *reinterpret_cast<float4*>(gp + rb) = r;
      }
    } else {
      for (int32_t i = lane; i < 32 * q; i += 32) {
        const int32_t p = i / q;
        const int32_t j = i - p * q;
        const int64_t rb = __shfl_sync(0xffffffffu, my_row, p);
        if (rb < 0) continue;
        const float4 r = *reinterpret_cast<const float4*>(s_out + p * ld_s + (j << 2));
        *reinterpret_cast<float4*>(out + rb + c_lo + (j << 2)) = r;
      }
    }
  }
}

// Channels-last grid (gr_sc == 1, C % 4 == 0): a point's four taps are four contiguous C*4-byte rows, so a lane
// fetches 16 bytes per tap and 32 lanes cover the taps of 32 / q points at once (q = C / 4 channel quads per point).
// Same shape as the planar kernel — one independent warp per 32 consecutive slots, no block barrier: the lanes first
// compute the sampling state of THEIR point (lanes = points) and park it in the warp's slice of shared memory, then
// the warp walks the points 32 / q at a time (lanes = (point, channel quad)): three 128-bit shared loads for the
// state, four unconditional 16-byte tap loads, sixteen FMAs and one 16-byte store per lane — the lanes of a point
// write its row as one contiguous piece, no output staging. Per output element that is a quarter of the planar
// kernel's load instructions and less than half of its instructions: channels_last feature maps are the layout this
// operator wants. (The round-1 version of this kernel was a 4-warp CTA sharing 32 points behind two barriers.)
constexpr int kNhwcWarps = 4;  // independent warps per CTA

template <bool ORDERED, bool TAPS>
__global__ void __launch_bounds__(kNhwcWarps * 32)
gather_forward_nhwc_kernel(const float* __restrict__ grid, int32_t C, int32_t H, int32_t W,
                           int64_t gr_sb, int64_t gr_sh, int64_t gr_sw,
                           const float* __restrict__ coord, int32_t N,
                           int64_t co_sb, int64_t co_sn, int64_t co_sd, float sh, float sw,
                           float* __restrict__ out, int64_t o_sb, int64_t o_sc, int64_t o_sn,
                           const int2* __restrict__ order, int32_t order_hw, int64_t order_len,
                           const TapsS* __restrict__ taps) {
  SMOS_PDL_PROLOGUE();
  __shared__ TapsS s_taps[kNhwcWarps][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t slot0 = (static_cast<int64_t>(blockIdx.x) * kNhwcWarps + wid) * 32;
  if (slot0 >= (ORDERED || TAPS ? order_len : static_cast<int64_t>(N))) return;  // whole warp past the end
  {
    TapsS t = lane_taps<ORDERED, TAPS>(slot0 + lane, blockIdx.z, coord, N, co_sb, co_sn, co_sd, sh, sw, H, W, order,
                                       order_hw, order_len, taps);
    // pixel offsets -> element offsets inside the batch image (the host checked that they fit 32 bits)
    t.o_nw = static_cast<int32_t>((t.o_nw / W) * gr_sh + (t.o_nw % W) * gr_sw);
    t.o_ne = static_cast<int32_t>((t.o_ne / W) * gr_sh + (t.o_ne % W) * gr_sw);
    t.o_sw = static_cast<int32_t>((t.o_sw / W) * gr_sh + (t.o_sw % W) * gr_sw);
    t.o_se = static_cast<int32_t>((t.o_se / W) * gr_sh + (t.o_se % W) * gr_sw);
    s_taps[wid][lane] = t;
  }
  __syncwarp();
  const int32_t q = C >> 2;  // channel quads per point
  const bool vec_ok = (o_sc == 1) && ((o_sn & 3) == 0) && ((o_sb & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  auto one = [&](int32_t p, int32_t cq) {
    const TapsS t = s_taps[wid][p];
    const float* gq = grid + t.b * gr_sb + (cq << 2);
    // unconditional loads from clamped pixels, selected afterwards (see the planar kernel); a slot without a point
    // (n < 0) still carries valid clamped offsets and b
    const float4 l_nw = __ldg(reinterpret_cast<const float4*>(gq + t.o_nw));
    const float4 l_ne = __ldg(reinterpret_cast<const float4*>(gq + t.o_ne));
    const float4 l_sw = __ldg(reinterpret_cast<const float4*>(gq + t.o_sw));
    const float4 l_se = __ldg(reinterpret_cast<const float4*>(gq + t.o_se));
    const float4 v_nw = (t.in_mask & 1u) ? l_nw : z, v_ne = (t.in_mask & 2u) ? l_ne : z;
    const float4 v_sw = (t.in_mask & 4u) ? l_sw : z, v_se = (t.in_mask & 8u) ? l_se : z;
    float4 a;
    a.x = fmaf(v_se.x, t.w_se, fmaf(v_sw.x, t.w_sw, fmaf(v_ne.x, t.w_ne, fmaf(v_nw.x, t.w_nw, 0.f))));
    a.y = fmaf(v_se.y, t.w_se, fmaf(v_sw.y, t.w_sw, fmaf(v_ne.y, t.w_ne, fmaf(v_nw.y, t.w_nw, 0.f))));
    a.z = fmaf(v_se.z, t.w_se, fmaf(v_sw.z, t.w_sw, fmaf(v_ne.z, t.w_ne, fmaf(v_nw.z, t.w_nw, 0.f))));
    a.w = fmaf(v_se.w, t.w_se, fmaf(v_sw.w, t.w_sw, fmaf(v_ne.w, t.w_ne, fmaf(v_nw.w, t.w_nw, 0.f))));
    if (t.n < 0) return;
    float* o = out + t.b * o_sb + static_cast<int64_t>(t.n) * o_sn + static_cast<int64_t>(cq << 2) * o_sc;
    if (vec_ok) {
      *reinterpret_cast<float4*>(o) = a;
    } else {
      o[0] = a.x; o[o_sc] = a.y; o[2 * o_sc] = a.z; o[3 * o_sc] = a.w;
    }
  };
  if (q <= 32 && (q & (q - 1)) == 0) {  // 32 / q points per step: lane = (point, quad)
    const int32_t shq = 31 - __clz(q);
    const int32_t ppi = 32 >> shq;
    const int32_t p_in = lane >> shq, cq = lane & (q - 1);
#pragma unroll 2
    for (int32_t g = 0; g < q; ++g) one(g * ppi + p_in, cq);
  } else {  // any other channel count: the warp takes the points one by one, lanes over the quads
    for (int32_t p = 0; p < 32; ++p)
      for (int32_t cq = lane; cq < q; cq += 32) one(p, cq);
  }
}

// Backward wrt the grid (grid_sampler_2d backward, input gradient only: coordinates are data).
// One thread per (b, c, n). With channel-major gradients consecutive lanes are consecutive points of one channel,
// and consecutive LiDAR returns mostly fall into the same pixel (a warp of 32 scan-order points touches ~4 distinct
// BEV pixels at 128^2, ~7 at 256^2): runs of lanes with the same north-west tap are summed in registers (segmented
// shuffle reduction) and only the head of each run issues the four atomics.
__global__ void __launch_bounds__(256)
gather_backward_kernel(const float* __restrict__ gout, int32_t C, int32_t N, int64_t total,
                       int64_t go_sb, int64_t go_sc, int64_t go_sn, const float* __restrict__ coord,
                       int64_t co_sb, int64_t co_sn, int64_t co_sd, float sh, float sw,
                       int32_t H, int32_t W, float* __restrict__ ggrid, int fast_n) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < total;
  float v_nw = 0.f, v_ne = 0.f, v_sw = 0.f, v_se = 0.f;
  bool in_nw = false, in_ne = false, in_sw = false, in_se = false;
  float* g = nullptr;
  // run key: (b, c, y0, x0) spelled out — the tap ADDRESS alone is not injective once x0 is clamped to [-2, W+1]
  // ((y0, -1) aliases (y0 - 1, W - 1)), and lanes merged across such an alias would share the head's in-image flags
  unsigned long long key = 0xffffffffffffff00ull | static_cast<unsigned>(threadIdx.x & 31);
  if (valid) {
    const int64_t cn = static_cast<int64_t>(C) * N;
    const int32_t b = static_cast<int32_t>(i / cn);
    const int64_t r = i - b * cn;
    int32_t c, n;
    if (fast_n) { c = static_cast<int32_t>(r / N); n = static_cast<int32_t>(r - static_cast<int64_t>(c) * N); }
    else        { n = static_cast<int32_t>(r / C); c = static_cast<int32_t>(r - static_cast<int64_t>(n) * C); }
    const float* cp = coord + b * co_sb + static_cast<int64_t>(n) * co_sn;
    const Taps t = make_taps(__ldg(cp), __ldg(cp + co_sd), sh, sw, H, W);
    const float go = gout[b * go_sb + c * go_sc + n * go_sn];
    g = ggrid + ((static_cast<int64_t>(b) * C + c) * H + t.y0) * W + t.x0;
    v_nw = go * t.w_nw; v_ne = go * t.w_ne; v_sw = go * t.w_sw; v_se = go * t.w_se;
    in_nw = t.in_nw; in_ne = t.in_ne; in_sw = t.in_sw; in_se = t.in_se;
    key = (static_cast<unsigned long long>(static_cast<uint32_t>(b) * static_cast<uint32_t>(C) + static_cast<uint32_t>(c)) << 32) |
          static_cast<uint32_t>((t.y0 + 2) * (W + 4) + (t.x0 + 2));
  }
  if (fast_n) {  // uniform: every lane of the warp takes part in the shuffles
    const int lane = threadIdx.x & 31;
    // equal key <=> equal (b, c, y0, x0) <=> equal in-image flags; the products are each lane's own
    const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane == 0 || key != prev;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned above = lane == 31 ? 0u : heads >> (lane + 1);
    const int run_end = above ? lane + __ffs(above) : 32;  // first lane of the next run
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float a0 = __shfl_down_sync(0xffffffffu, v_nw, d), a1 = __shfl_down_sync(0xffffffffu, v_ne, d);
      const float a2 = __shfl_down_sync(0xffffffffu, v_sw, d), a3 = __shfl_down_sync(0xffffffffu, v_se, d);
      if (lane + d < run_end) { v_nw += a0; v_ne += a1; v_sw += a2; v_se += a3; }
    }
    if (!head) return;
  }
  if (!valid) return;
  if (in_nw) atomicAdd(g, v_nw);
  if (in_ne) atomicAdd(g + 1, v_ne);
  if (in_sw) atomicAdd(g + W, v_sw);
  if (in_se) atomicAdd(g + W + 1, v_se);
}

}  // namespace

extern "C" {

static int gather_forward_launch(const float* grid, int64_t B, int64_t C, int32_t H, int32_t W, int64_t gr_sb,
                                 int64_t gr_sc, int64_t gr_sh, int64_t gr_sw, const float* coord, int64_t N,
                                 int64_t co_sb, int64_t co_sn, int64_t co_sd, float scale_h, float scale_w,
                                 float* out, int64_t o_sb, int64_t o_sc, int64_t o_sn, const int2* order,
                                 int32_t order_hw, const TapsS* taps, void* stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || N < 0) return SMOS_EINVAL;
  if (N == 0) return SMOS_OK;
  if (!grid || (!coord && !taps) || !out) return SMOS_EINVAL;
  if (B > 65535 || N >= (int64_t(1) << 31) || C >= (1 << 24) || B * N >= (int64_t(1) << 31)) return SMOS_EUNSUPPORTED;
  cudaStream_t st = smos_stream(stream);
  const bool nhwc = (gr_sc == 1) && ((C & 3) == 0) && ((gr_sw & 3) == 0) && ((gr_sh & 3) == 0) &&
                    ((gr_sb & 3) == 0) && ((reinterpret_cast<uintptr_t>(grid) & 15) == 0) &&
                    (static_cast<int64_t>(H - 1) * gr_sh + static_cast<int64_t>(W - 1) * gr_sw < (int64_t(1) << 31));
  // cell order only pays when a point's channels leave as whole sectors (point-major output rows)
  const bool with_taps = taps != nullptr;  // implies slot (cell) order; the caller checked the output layout
  const bool ordered = (order != nullptr && o_sc == 1 && C > 1) || with_taps;
  const int64_t order_len = B * N;
  dim3 g(smos_ceil_div(ordered ? order_len : N, 32 * kNhwcWarps), 1, ordered ? 1u : static_cast<unsigned>(B));
  dim3 gp(smos_ceil_div(ordered ? order_len : N, 32 * kGatherWarps), 1, ordered ? 1u : static_cast<unsigned>(B));
  // one warp = 32 slots x (all channels / split). The split adds warps (latency hiding) at the price of recomputing
  // the sampling state per warp: by default aim at >= 48 resident warps per SM
  const int32_t ngroups_p = static_cast<int32_t>((C + kCPT - 1) / kCPT);
  int32_t split = smos_env_int("SMOS_GATHER_SPLIT", 0);
  if (split <= 0) {
    split = 1;
    const int64_t warps = static_cast<int64_t>(gp.x) * gp.z * kGatherWarps;
    while (split < 4 && split * 2 <= ngroups_p && warps * split < static_cast<int64_t>(48) * SMOS_SM_COUNT) split *= 2;
  }
  if (split > ngroups_p) split = ngroups_p;
  const int32_t groups_per_warp = (ngroups_p + split - 1) / split;
  gp.y = static_cast<unsigned>((ngroups_p + groups_per_warp - 1) / groups_per_warp);
  const int32_t Ni = static_cast<int32_t>(N), Ci = static_cast<int32_t>(C);
  if (nhwc) {
    if (with_taps)
      SMOS_LAUNCH((gather_forward_nhwc_kernel<true, true>), g, kNhwcWarps * 32, 0, st, grid, Ci, H, W, gr_sb, gr_sh, gr_sw, coord, Ni,
                                                                            co_sb, co_sn, co_sd, scale_h, scale_w, out,
                                                                            o_sb, o_sc, o_sn, order, order_hw, order_len, taps);
    else if (ordered)
      SMOS_LAUNCH((gather_forward_nhwc_kernel<true, false>), g, kNhwcWarps * 32, 0, st, grid, Ci, H, W, gr_sb, gr_sh, gr_sw, coord, Ni,
                                                                             co_sb, co_sn, co_sd, scale_h, scale_w, out,
                                                                             o_sb, o_sc, o_sn, order, order_hw, order_len, nullptr);
    else
      SMOS_LAUNCH((gather_forward_nhwc_kernel<false, false>), g, kNhwcWarps * 32, 0, st, grid, Ci, H, W, gr_sb, gr_sh, gr_sw, coord,
                                                                              Ni, co_sb, co_sn, co_sd, scale_h, scale_w, out,
                                                                              o_sb, o_sc, o_sn, nullptr, 1, 0, nullptr);
  } else {
    // point-major output rows (C % 8 == 0, 16-byte aligned) are assembled in shared memory and leave as whole lines
    const size_t smem_rows = static_cast<size_t>(groups_per_warp * kCPT + 4) * 32 * 4 * kGatherWarps;
    const bool rows_out = (o_sc == 1) && C > 1 && ((C & 7) == 0) && ((o_sn & 3) == 0) && ((o_sb & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (smem_rows <= 46 * 1024) &&
                          smos_env_int("SMOS_GATHER_ROWS", 1);
    const size_t smem = rows_out ? smem_rows : 0;
    const bool dense = (gr_sw == 1 && gr_sh == W);
#define SMOS_LAUNCH_PLANAR(D, R, O, T)                                                                       \
    SMOS_LAUNCH((gather_forward_planar_kernel<D, R, O, T>), gp, kGatherWarps * 32, smem, st,                                  \
        grid, Ci, H, W, gr_sb, gr_sc, gr_sh, gr_sw, coord, Ni, co_sb, co_sn, co_sd, scale_h, scale_w, out, o_sb, \
        o_sc, o_sn, order, order_hw, order_len, taps, groups_per_warp)
    if (with_taps) {
      if (dense && rows_out) SMOS_LAUNCH_PLANAR(true, true, true, true);
      else if (dense) SMOS_LAUNCH_PLANAR(true, false, true, true);
      else if (rows_out) SMOS_LAUNCH_PLANAR(false, true, true, true);
      else SMOS_LAUNCH_PLANAR(false, false, true, true);
    } else if (ordered) {
      if (dense && rows_out) SMOS_LAUNCH_PLANAR(true, true, true, false);
      else if (dense) SMOS_LAUNCH_PLANAR(true, false, true, false);
      else if (rows_out) SMOS_LAUNCH_PLANAR(false, true, true, false);
      else SMOS_LAUNCH_PLANAR(false, false, true, false);
    } else {
      if (dense && rows_out) SMOS_LAUNCH_PLANAR(true, true, false, false);
      else if (dense) SMOS_LAUNCH_PLANAR(true, false, false, false);
      else if (rows_out) SMOS_LAUNCH_PLANAR(false, true, false, false);
      else SMOS_LAUNCH_PLANAR(false, false, false, false);
    }
#undef SMOS_LAUNCH_PLANAR
  }
  return smos_launch_status();
}

int smos_bilinear_gather_forward(const float* grid, int64_t B, int64_t C, int32_t H, int32_t W, int64_t gr_sb,
                                 int64_t gr_sc, int64_t gr_sh, int64_t gr_sw, const float* coord, int64_t N,
                                 int64_t co_sb, int64_t co_sn, int64_t co_sd, float scale_h, float scale_w,
                                 float* out, int64_t o_sb, int64_t o_sc, int64_t o_sn, void* stream) {
  return gather_forward_launch(grid, B, C, H, W, gr_sb, gr_sc, gr_sh, gr_sw, coord, N, co_sb, co_sn, co_sd, scale_h,
                               scale_w, out, o_sb, o_sc, o_sn, nullptr, 1, nullptr, stream);
}

int smos_bilinear_gather_forward_ordered(const float* grid, int64_t B, int64_t C, int32_t H, int32_t W, int64_t gr_sb,
                                         int64_t gr_sc, int64_t gr_sh, int64_t gr_sw, const float* coord, int64_t N,
                                         int64_t co_sb, int64_t co_sn, int64_t co_sd, float scale_h, float scale_w,
                                         float* out, int64_t o_sb, int64_t o_sc, int64_t o_sn, const void* plan,
                                         int32_t plan_H, int32_t plan_W, void* stream) {
  if (!plan || plan_H <= 0 || plan_W <= 0 || B <= 0 || N < 0) return SMOS_EINVAL;
  const PoolLayout L = smos_pool_layout(B, N, plan_H, plan_W);
  const int2* sorted = reinterpret_cast<const int2*>(static_cast<const char*>(plan) + L.off_sorted);
  return gather_forward_launch(grid, B, C, H, W, gr_sb, gr_sc, gr_sh, gr_sw, coord, N, co_sb, co_sn, co_sd, scale_h,
                               scale_w, out, o_sb, o_sc, o_sn, sorted, static_cast<int32_t>(L.hw), nullptr, stream);
}

int64_t smos_gather_taps_bytes(int64_t B, int64_t N) {
  if (B <= 0 || N < 0) return SMOS_EINVAL;
  return B * N * static_cast<int64_t>(sizeof(TapsS));
}

int smos_bilinear_gather_forward_taps(const float* grid, int64_t B, int64_t C, int32_t H, int32_t W, int64_t gr_sb,
                                      int64_t gr_sc, int64_t gr_sh, int64_t gr_sw, const void* taps, int64_t N,
                                      float* out, int64_t o_sb, int64_t o_sc, int64_t o_sn, void* stream) {
  if (!taps || (reinterpret_cast<uintptr_t>(taps) & 15) != 0 || o_sc != 1) return SMOS_EINVAL;
  return gather_forward_launch(grid, B, C, H, W, gr_sb, gr_sc, gr_sh, gr_sw, nullptr, N, 0, 0, 0, 0.f, 0.f, out, o_sb,
                               o_sc, o_sn, nullptr, 1, static_cast<const TapsS*>(taps), stream);
}

int smos_bilinear_gather_backward(const float* grad_out, int64_t B, int64_t C, int64_t N, int64_t go_sb,
                                  int64_t go_sc, int64_t go_sn, const float* coord, int64_t co_sb,
                                  int64_t co_sn, int64_t co_sd, float scale_h, float scale_w, int32_t H,
                                  int32_t W, float* grad_grid, void* stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || N < 0) return SMOS_EINVAL;
  if (!grad_grid) return SMOS_EINVAL;
  if (N == 0)  // no point: the gradient is zero
    return static_cast<int>(smos_zero_async(grad_grid, static_cast<size_t>(B * C) * H * W * sizeof(float),
                                            smos_stream(stream)));
  if (!grad_out || !coord) return SMOS_EINVAL;
  if (N >= (int64_t(1) << 31) || B * C >= (int64_t(1) << 24)) return SMOS_EUNSUPPORTED;
  cudaStream_t st = smos_stream(stream);
  // global fp32 atomics (warp-aggregated over runs of equal pixels) into a gradient zero-filled here. Measured and
  // rejected: privatising row bands of the planes in shared memory ([pixel][32 channels] tiles, lane = channel) is
  // 1.1-1.8x SLOWER on LiDAR-shaped input — the hottest BEV row holds 7x the mean number of points and single cells
  // up to 1.5 k, so the few CTAs that own those tiles serialise on them.
  if (smos_zero_async(grad_grid, static_cast<size_t>(B * C) * H * W * sizeof(float), st) != cudaSuccess)
    return smos_launch_status();
  const int64_t total = B * C * N;
  const int fast_n = (go_sc == 1 && C > 1) ? 0 : 1;
  SMOS_LAUNCH((gather_backward_kernel), smos_ceil_div(total, 256), 256, 0, smos_stream(stream), 
      grad_out, static_cast<int32_t>(C), static_cast<int32_t>(N), total, go_sb, go_sc, go_sn, coord, co_sb, co_sn,
      co_sd, scale_h, scale_w, H, W, grad_grid, fast_n);
  return smos_launch_status();
}

}  // extern "C"
