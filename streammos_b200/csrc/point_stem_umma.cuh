// PointNet stem, layer 2 on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) — included by point_stem.cu.
//
// The scalar kernel in point_stem.cu runs the 64 -> 64 layer on the CUDA cores: 1.5 GFMA per scan, 46 us at the FP32
// peak, 91 us measured. Here the layer is a 128 x 64 x 64 GEMM per tile of 128 points on the tensor cores, in fp32
// accuracy through the 3xTF32 split (a = a_hi + a_lo with a_hi = tf32(a), a_lo = tf32(a - a_hi); a.b ~ a_lo.b_hi +
// a_hi.b_lo + a_hi.b_hi, fp32 accumulation in TMEM: relative error ~1e-6, inside the 1e-5 parity bar, no longer
// bit-identical to the scalar FMA order). A plain TF32 product (what torch/cuDNN use for these 1x1 convolutions) misses
// the bar by two orders of magnitude.
//
// Per CTA (one per SM, persistent over its tiles of 128 points; 512 threads = four per point, 16 channels each):
//   layer 1 (7 -> 64, CUDA cores, 448 FMA per point, four threads per point) -> BatchNorm + ReLU -> hi / lo split -> shared memory in the UMMA
//   canonical K-major layout without swizzle (core matrix = 8 rows x 16 bytes, 128 contiguous bytes; K-adjacent core
//   matrices 128 bytes apart = LBO, 8-row groups 2048 bytes apart = SBO);
//   fence.proxy.async + mbarrier arrive; the elected lane of a 17th warp issues 24 x tcgen05.mma.kind::tf32 (M = 128, N = 64, K = 8: eight K steps x
//   three split products) into a 64-column TMEM accumulator and commits them to an mbarrier;
//   while they run, all threads drain the PREVIOUS tile's accumulator (tcgen05.ld 32x32b.x16: lane = point, 16 channels),
//   apply BatchNorm + ReLU and store; operands and accumulators are double buffered.
// Measured on B200 (3 x 120 000 points): 62 us against 91 us for the CUDA-core kernel. Of those, 20 us are the 92 MB of
// output stores and 10 us the 448 FMA per point of layer 1; the tensor pipe is 11 % busy. What is left is the latency
// chain of a tile (operand stores -> proxy fence -> MMA -> TMEM load -> stores) with one CTA per SM: two CTAs per SM
// (single operand stage, 98 KB each) are the next step.
// W2 (hi and lo parts) sits in shared memory for the lifetime of the CTA in the same canonical layout (row = output
// channel, K-major: exactly its row-major (C2, C1) global layout).
#pragma once

namespace umma_stem {

constexpr int kPts = 128;
constexpr int kC = 64;
constexpr int kParts = 4;              // threads per point: each owns kC / kParts hidden / output channels
constexpr int kComputeThreads = kPts * kParts;  // 512: 16 warps per SM keep the CUDA-core half of the kernel busy
constexpr int kThreads = kComputeThreads + 32;  // + one warp whose elected lane issues the MMAs
constexpr int kCP = kC / kParts;       // 16 channels per thread
constexpr int kCinMax = 16;
constexpr uint32_t kLBO = 128;    // bytes between core matrices adjacent in K
constexpr uint32_t kSBO = 2048;   // bytes between 8-row groups: 16 K-chunks x 128 bytes
constexpr uint32_t kTmemCols = 128;  // two 64-column accumulators

struct Smem {
  float a_hi[2][kPts * kC];  // operand stages, canonical layout
  float a_lo[2][kPts * kC];
  float b_hi[kC * kC];
  float b_lo[kC * kC];
  alignas(16) float w1t[kCinMax][kC];  // W1 transposed: [input][hidden channel]
  float2 ab1[kC];
  alignas(16) float a2[kC];
  alignas(16) float b2[kC];
  float a0[kCinMax], b0[kCinMax];
  float4 rows[kComputeThreads / 32][32 * 4];  // per-warp staging of point-major output rows (64 bytes per row, swizzled)
  uint64_t full[2];  // operand stage written (one arrival per compute warp)
  uint64_t done[2];  // the stage's MMAs have completed (tcgen05.commit)
  uint32_t tmem_base;
};

// float index of element (row, k) of a K-major operand in the canonical no-swizzle layout
__device__ __forceinline__ int core_index(int row, int k) { return (row >> 3) * 512 + (k >> 2) * 32 + (row & 7) * 4 + (k & 3); }

using ::tf32_rna;  // cvt.rna.tf32.f32, defined in point_stem.cu

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO and SBO in 16-byte units,
// version 1 (sm_100), no swizzle
__device__ __forceinline__ uint64_t smem_desc(const void* p) {
  const uint32_t addr = smos_smem_u32(p);
  uint64_t d = static_cast<uint64_t>((addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(kLBO >> 4) << 16;
  d |= static_cast<uint64_t>(kSBO >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kC >> 3) << 17) | ((kPts >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smos_smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 consecutive accumulator columns of this thread's TMEM lane (= its point)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace umma_stem

template <int CIN, bool RAW>
__global__ void __launch_bounds__(umma_stem::kThreads, 1)
point_stem_umma_kernel(const float* __restrict__ x, int32_t Cin, int32_t N, int32_t B, int64_t x_sb, int64_t x_sc,
                       int64_t x_sn, const __grid_constant__ StemRaw raw, const float* __restrict__ a0,
                       const float* __restrict__ b0, const float* __restrict__ w1, const float* __restrict__ a1,
                       const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ a2,
                       const float* __restrict__ b2, float* __restrict__ y, int64_t y_sb, int64_t y_sc, int64_t y_sn) {
  using namespace umma_stem;
  SMOS_PDL_PROLOGUE();
  extern __shared__ __align__(128) unsigned char umma_raw[];
  Smem& S = *reinterpret_cast<Smem*>(umma_raw);
  constexpr int KI = CIN > 0 ? (CIN + 3) / 4 * 4 : kCinMax;
  const int tid = threadIdx.x, wid = tid >> 5;
  const int pt = tid & (kPts - 1), part = tid >> 7;  // my point of the tile, my quarter of the channels
  // ---- one-time setup: parameters, W2 split into its tf32 hi / lo parts, barriers, tensor memory ------------------------
  {
    // all loads of a thread first (8 in flight), then the splits and stores: one round trip instead of eight (the setup
    // was 5 % of the kernel's stall samples)
    constexpr int kW2PerThread = (kC * kC + kThreads - 1) / kThreads;
    float wv[kW2PerThread];
#pragma unroll
    for (int u = 0; u < kW2PerThread; ++u) wv[u] = __ldg(w2 + min(tid + u * kThreads, kC * kC - 1));
#pragma unroll
    for (int u = 0; u < kW2PerThread; ++u) {
      const int i = tid + u * kThreads;
      if (i < kC * kC) {
        const int n = i >> 6, k = i & 63;  // W2[n][k], row = output channel
        const float hi = tf32_rna(wv[u]);  // once per CTA: round to nearest here, the remainder is exact
        S.b_hi[core_index(n, k)] = hi;
        S.b_lo[core_index(n, k)] = wv[u] - hi;
      }
    }
  }
  for (int i = tid; i < kC * kCinMax; i += kThreads) {
    const int c = i / kCinMax, ci = i % kCinMax;
    S.w1t[ci][c] = ci < Cin ? __ldg(w1 + c * Cin + ci) : 0.f;
  }
  if (tid < kC) { S.ab1[tid] = make_float2(__ldg(a1 + tid), __ldg(b1 + tid)); S.a2[tid] = __ldg(a2 + tid); S.b2[tid] = __ldg(b2 + tid); }
  if (tid < Cin) { S.a0[tid] = a0 ? __ldg(a0 + tid) : 1.f; S.b0[tid] = b0 ? __ldg(b0 + tid) : 0.f; }
  if (tid == 0) {
    smos_mbar_init(&S.full[0], kComputeThreads / 32);  // one arrival per compute warp
    smos_mbar_init(&S.full[1], kComputeThreads / 32);
    smos_mbar_init(&S.done[0], 1);
    smos_mbar_init(&S.done[1], 1);
    smos_fence_mbar_init();
  }
  __syncwarp();
  if (wid == kComputeThreads / 32) {  // the MMA warp allocates (and later frees) the accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smos_smem_u32(&S.tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  smos_fence_proxy_async();  // W2 parts were written through the generic proxy; the tensor core reads through the async one
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  const int32_t tiles_per_b = (N + kPts - 1) / kPts;
  const int32_t ntiles = tiles_per_b * B;
  // tile -> (batch, first point) without an integer division per tile: the grid stride is split once
  const int32_t stride_b = static_cast<int32_t>(gridDim.x) / tiles_per_b;
  const int32_t stride_n = (static_cast<int32_t>(gridDim.x) - stride_b * tiles_per_b) * kPts;
  auto advance = [&](int32_t& b, int32_t& n0) {
    b += stride_b;
    n0 += stride_n;
    if (n0 >= tiles_per_b * kPts) { n0 -= tiles_per_b * kPts; ++b; }
  };
  // Inputs of my point of a tile, in two halves: fetch() only ISSUES the loads (one grid stride ahead, so that they are
  // in flight during the epilogue of the previous tile), expand() turns them into the layer's inputs at the top of
  // the tile's own iteration. (With the arithmetic inside fetch() the first multiply waited for the load right
  // there: 12 % of the kernel's stall samples.) Same arithmetic as point_stem_kernel.
  auto fetch = [&](int32_t b, int32_t n0, float (&xr)[KI], float4& rawq) {
    const int32_t n = min(n0 + pt, N - 1);
    if (RAW) {
      const float* p = raw.pts + (static_cast<int64_t>(b) * N + n) * raw.rs;
      if (raw.rs == 4 && (reinterpret_cast<uintptr_t>(raw.pts) & 15) == 0) {
        rawq = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        rawq = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
      }
      return;
    }
#pragma unroll
    for (int ci = 0; ci < KI; ++ci)
      xr[ci] = (CIN > 0 ? ci < CIN : ci < Cin) ? __ldg(x + b * x_sb + ci * x_sc + static_cast<int64_t>(n) * x_sn) : 0.f;
  };
  auto expand = [&](int32_t b, int32_t n0, float (&xr)[KI], const float4& rawq) {
    if (!RAW) return;
    const int32_t nn = n0 + pt;
    const int32_t n = min(nn, N - 1);
    const float px = __fmul_rn(rawq.x, raw.sx);
    const float py = __fmul_rn(rawq.y, raw.sy);
    const float pz = rawq.z, pw = rawq.w;
    const float qx = __fdiv_rn(__fsub_rn(px, raw.mx), raw.dx);
    const float qy = __fdiv_rn(__fsub_rn(py, raw.my), raw.dy);
    // (the third quotient only feeds the coordinate output: one of the four threads of a point computes it)
    const float qz = part == 0 ? __fdiv_rn(__fsub_rn(pz, raw.mz), raw.dz) : 0.f;
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz));
    xr[0] = px; xr[1] = py; xr[2] = pz; xr[3] = pw;
    xr[4] = __fadd_rn(__fsqrt_rn(d2), 1e-12f);
    xr[5] = __fsub_rn(qx, floorf(qx));
    xr[6] = __fsub_rn(qy, floorf(qy));
#pragma unroll
    for (int ci = 7; ci < KI; ++ci) xr[ci] = 0.f;
    if (nn < N && part == 0) {
      float* c = raw.coord + (static_cast<int64_t>(b) * N + n) * 3;
      c[0] = qx; c[1] = qy; c[2] = qz;
    }
  };
  // accumulator of tile `t` (stage s) -> BatchNorm + ReLU -> y
  auto drain = [&](int32_t b, int32_t n0, int s, uint32_t parity) {
    smos_mbar_wait_sleepy(&S.done[s], parity);  // the tile's 24 MMAs have completed
    fence_after_sync();
    const int32_t n = n0 + pt;
    // warp w may touch TMEM lanes 32 * (w % 4) ...: exactly the 32 points of its quarter of the tile
    const uint32_t taddr = tmem + (static_cast<uint32_t>((wid & 3) * 32) << 16) + static_cast<uint32_t>(s * kC + part * kCP);
    const bool rows_out = y_sc == 1 && (y_sn & 3) == 0 && (y_sb & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    float v[kCP];
    tmem_ld16(taddr, v);
#pragma unroll
    for (int j4 = 0; j4 < kCP / 4; ++j4) {
      const float4 al = *reinterpret_cast<const float4*>(&S.a2[part * kCP + 4 * j4]);
      const float4 be = *reinterpret_cast<const float4*>(&S.b2[part * kCP + 4 * j4]);
      v[4 * j4] = fmaxf(fmaf(v[4 * j4], al.x, be.x), 0.f);
      v[4 * j4 + 1] = fmaxf(fmaf(v[4 * j4 + 1], al.y, be.y), 0.f);
      v[4 * j4 + 2] = fmaxf(fmaf(v[4 * j4 + 2], al.z, be.z), 0.f);
      v[4 * j4 + 3] = fmaxf(fmaf(v[4 * j4 + 3], al.w, be.w), 0.f);
    }
    if (rows_out) {
      // point-major rows: my 16 channels are 64 contiguous bytes of my point's row. Written straight from the registers,
      // every store instruction would touch 32 rows with 16 bytes each (half sectors: measured 25 of the kernel's 66 us).
      // Through the warp's staging buffer four consecutive lanes write one 64-byte segment: whole sectors only.
      // Staging layout: row r = 64 contiguous bytes, its 16-byte chunk c at slot c ^ ((r >> 1) & 3). Writes (lane = row,
      // one chunk index per instruction) and reads (four lanes = the four chunks of a row, eight lanes = two adjacent
      // rows = 128 contiguous bytes) are both conflict free; the earlier pitch of 80 bytes cost the reads a second
      // wavefront (ncu: 2x the ideal on the four LDS.128 of the epilogue, 8 % of the kernel's shared-memory wavefronts).
      float4* stg = S.rows[wid];
      const int lane = tid & 31;
#pragma unroll
      for (int j = 0; j < kCP / 4; ++j)
        stg[lane * 4 + (j ^ ((lane >> 1) & 3))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
      const int32_t nw = n0 + (wid & 3) * 32;  // first point of my warp's quarter of the tile
      // lane l moves chunk l & 3 of rows (l >> 2) + 8 k: one base pointer per tile, a constant step per store, and the
      // staged chunks fetched together before the first store (the stores used to wait on their LDS one by one)
      const int r0 = lane >> 2, seg = lane & 3;
      float* yb = y + b * y_sb + static_cast<int64_t>(nw + r0) * y_sn + part * kCP + seg * 4;
      const int64_t step = 8 * y_sn;
      float4 o[kCP / 4];
#pragma unroll
      for (int k = 0; k < kCP / 4; ++k) {
        const int row = r0 + 8 * k;
        o[k] = stg[row * 4 + (seg ^ ((row >> 1) & 3))];
      }
      if (nw + 32 <= N) {
#pragma unroll
        for (int k = 0; k < kCP / 4; ++k) *reinterpret_cast<float4*>(yb + k * step) = o[k];
      } else {
#pragma unroll
        for (int k = 0; k < kCP / 4; ++k)
          if (nw + r0 + 8 * k < N) *reinterpret_cast<float4*>(yb + k * step) = o[k];
      }
      __syncwarp();
    }
    if (n < N) {
      float* dst = y + b * y_sb + static_cast<int64_t>(n) * y_sn + static_cast<int64_t>(part * kCP) * y_sc;
      if (!rows_out) {  // channel-major: a warp's 32 points are contiguous per channel
#pragma unroll
        for (int j = 0; j < kCP; ++j) dst[static_cast<int64_t>(j) * y_sc] = v[j];
      }
    }
    fence_before_sync();  // my tcgen05.ld of this stage are done before my next arrival lets an MMA overwrite it
  };

  if (wid == kComputeThreads / 32) {
    // ---- MMA issuer: one lane. Waits for an operand stage, issues its 24 MMAs, commits them to the stage's barrier ----
    if ((tid & 31) == 0) {
      const uint64_t dbh0 = smem_desc(S.b_hi), dbl0 = smem_desc(S.b_lo);
      int32_t it = 0;
      for (int32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it & 1;
        smos_mbar_wait_sleepy(&S.full[s], static_cast<uint32_t>((it >> 1) & 1));
        fence_after_sync();
        const uint32_t d = tmem + static_cast<uint32_t>(s * kC);
        const uint64_t dah0 = smem_desc(S.a_hi[s]), dal0 = smem_desc(S.a_lo[s]);
#pragma unroll
        for (int ks = 0; ks < kC / 8; ++ks) {  // K = 8 per instruction: two 16-byte K-chunks, 256 bytes (16 units) per step
          const uint64_t step = static_cast<uint64_t>(ks * 16);
          mma_tf32(d, dal0 + step, dbh0 + step, ks > 0 ? 1u : 0u);  // small terms first
          mma_tf32(d, dah0 + step, dbl0 + step, 1u);
          mma_tf32(d, dah0 + step, dbh0 + step, 1u);
        }
        mma_commit(&S.done[s]);  // arrives when all MMAs issued so far have completed
      }
    }
  } else {
    // ---- compute warps: layer 1 of tile i, then (while the tensor core works on it) the epilogue of tile i - 1 ------------
    float xr[KI];
    float4 rawq = make_float4(0.f, 0.f, 0.f, 0.f);
    int32_t t = blockIdx.x;
    int32_t b_cur = t / tiles_per_b, n_cur = (t - b_cur * tiles_per_b) * kPts;  // this tile
    int32_t b_nxt = b_cur, n_nxt = n_cur;                                        // the tile one grid stride ahead
    int32_t b_prev = 0, n_prev = 0;
    if (t < ntiles) fetch(b_cur, n_cur, xr, rawq);
    int32_t it = 0, t_prev = -1;
    for (; t < ntiles; t += gridDim.x, ++it) {
      const int s = it & 1;
      expand(b_cur, n_cur, xr, rawq);
      {
        float xin[KI];
#pragma unroll
        for (int ci = 0; ci < KI; ++ci) xin[ci] = (CIN > 0 ? ci < CIN : ci < Cin) ? fmaf(xr[ci], S.a0[ci], S.b0[ci]) : 0.f;
        float* ah = S.a_hi[s] + (pt >> 3) * 512 + (pt & 7) * 4;
        float* al = S.a_lo[s] + (pt >> 3) * 512 + (pt & 7) * 4;
        // 16 independent accumulators (my hidden channels), inputs outermost: every FMA of a step is independent of
        // the others, and one 128-bit broadcast load brings the weights of four channels
        float acc[kCP];
#pragma unroll
        for (int j = 0; j < kCP; ++j) acc[j] = 0.f;
#pragma unroll
        for (int ci = 0; ci < KI; ++ci) {
          if (CIN > 0 && ci >= CIN) continue;  // zero-padded inputs
#pragma unroll
          for (int g = 0; g < kCP / 4; ++g) {
            const float4 w = *reinterpret_cast<const float4*>(&S.w1t[ci][part * kCP + 4 * g]);
            acc[4 * g] = fmaf(w.x, xin[ci], acc[4 * g]);
            acc[4 * g + 1] = fmaf(w.y, xin[ci], acc[4 * g + 1]);
            acc[4 * g + 2] = fmaf(w.z, xin[ci], acc[4 * g + 2]);
            acc[4 * g + 3] = fmaf(w.w, xin[ci], acc[4 * g + 3]);
          }
        }
#pragma unroll
        for (int qq = 0; qq < kCP / 4; ++qq) {
          const int q = part * (kCP / 4) + qq;  // 16-byte K-chunk = four hidden channels
          float hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 ab = S.ab1[4 * q + j];
            const float h = fmaxf(fmaf(acc[4 * qq + j], ab.x, ab.y), 0.f);
            // hi = the tf32 the tensor core would read anyway (it ignores the low 13 mantissa bits), lo = the exact
            // remainder (< 2^-10 |h|, of which the tensor core again keeps the top 11 bits): |h - hi - tf32(lo)| < 2^-21 |h|.
            // cvt.rna.tf32 is an 8-instruction sequence on sm_100; this is one LOP3 and one FADD
            hi[j] = __uint_as_float(__float_as_uint(h) & 0xffffe000u);
            lo[j] = h - hi[j];
          }
          *reinterpret_cast<float4*>(ah + q * 32) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(al + q * 32) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      advance(b_nxt, n_nxt);
      if (t + gridDim.x < ntiles) fetch(b_nxt, n_nxt, xr, rawq);  // in flight during the epilogue below
      smos_fence_proxy_async();  // my operand stores -> visible to the tensor core (async proxy)
      fence_before_sync();       // and my earlier tcgen05.ld of this stage's accumulator are complete
      __syncwarp();
      if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smos_smem_u32(&S.full[s])) : "memory");
      if (t_prev >= 0) drain(b_prev, n_prev, s ^ 1, static_cast<uint32_t>(((it - 1) >> 1) & 1));
      t_prev = t;
      b_prev = b_cur; n_prev = n_cur;
      b_cur = b_nxt; n_cur = n_nxt;
    }
    if (t_prev >= 0) drain(b_prev, n_prev, (it - 1) & 1, static_cast<uint32_t>(((it - 1) >> 1) & 1));
  }
  __syncthreads();
  if (wid == kComputeThreads / 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}
