// PointNet stem for sm_100a: the two stacked point-wise layers in front of VoxelMaxPool #1.
//
// Replaces networks/backbone.py:199-250 PointNetStacker(7, 64, pre_bn=True, stack_num=2) in eval mode as
// models/StreamMOS.py:77,101 uses it: BatchNorm2d(Cin) -> Conv2d 1x1 (Cin -> C1, no bias) -> BatchNorm2d -> ReLU ->
// Conv2d 1x1 (C1 -> C2) -> BatchNorm2d -> ReLU, i.e. 7 torch kernels and six passes over (B', 64, N) tensors. Here
// one kernel reads the (B', Cin, N) input once and writes the (B', C2, N) output once (the channel-major tensor
// VoxelMaxPool #1 consumes); the hidden activations live in shared memory. Every eval-mode BatchNorm is the
// per-channel affine torch applies (y = x * alpha + beta, alpha = weight / sqrt(var + eps), beta = bias - mean *
// alpha), kept separate from the convolutions so the arithmetic follows the reference's sequence. fp32 FMA on the
// CUDA cores: TF32 tensor-core products would miss the 1e-5 parity bar.
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int kStemPts = 128;     // points per CTA
constexpr int kStemThreads = 128;  // 4 warps: one thread per point in layer 1, 16 channels x 4 points per thread in layer 2
constexpr int kStemC = 64;        // C1 == C2 == 64 (the only configuration StreamMOS builds)
constexpr int kStemCinMax = 16;
#ifndef SMOS_STEM_MIN_CTAS
#define SMOS_STEM_MIN_CTAS 3  // four CTAs per SM (128 registers) spill with the packed accumulators: 98.8 vs 90.6 us
#endif

struct StemSmem {
  float h[kStemC][kStemPts];      // hidden activations, [channel][point]
  float w2t[kStemC][kStemC];      // W2 transposed: [k][c]
  float w1[kStemC][kStemCinMax];
  float2 ab1[kStemC];             // (alpha, beta) of the hidden BatchNorm, one 64-bit broadcast load per channel
  float a2[kStemC], b2[kStemC];
  float a0[kStemCinMax], b0[kStemCinMax];
};

// Tensor-core variant of layer 2 (SMOS_STEM_TC=1): 3xTF32 split products on mma.sync.m16n8k8 (a = a_hi + a_lo with
// a_hi = tf32(a), a_lo = tf32(a - a_hi); a*b ~ a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulation): relative error
// ~1e-6, inside the 1e-5 parity bar but no longer bit-identical to the scalar FMA order. Row pitches = 8 mod 32 floats
// make every fragment load (8 consecutive elements x 4 rows per warp) bank-conflict free.
constexpr int kHPitchTC = kStemPts + 8;   // 136
constexpr int kWPitchTC = kStemC + 8;     // 72
struct StemSmemTC {
  float h[kStemC][kHPitchTC];      // hidden activations, [k][point]
  float w2hi[kStemC][kWPitchTC];   // tf32(W2) transposed: [k][c]
  float w2lo[kStemC][kWPitchTC];   // tf32(W2 - tf32(W2))
  float w1[kStemC][kStemCinMax];
  float2 ab1[kStemC];
  float a2[kStemC], b2[kStemC];
  float a0[kStemCinMax], b0[kStemCinMax];
};

__device__ __forceinline__ float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const float (&a)[4], float b0, float b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// RAW: the 7 input channels are built on the fly from raw points (x, y, z, intensity): the loader's Quantize +
// make_point_feat (see form_batch.cu, same float32 sequence, bit-exact), and the (B, N, 3) quantised coordinates leave
// as a side output — the (B, 7, N) tensor never exists.
struct StemRaw {
  const float* pts;  // (B*N, rs) raw points, or null
  int64_t rs;
  float sx, sy, mx, my, mz, dx, dy, dz;
  float* coord;      // (B, N, 3) out
};

template <int CIN, bool RAW, bool TC>
__global__ void __launch_bounds__(kStemThreads, SMOS_STEM_MIN_CTAS)
point_stem_kernel(const float* __restrict__ x, int32_t Cin, int32_t N, int32_t B, int64_t x_sb, int64_t x_sc, int64_t x_sn,
                  const __grid_constant__ StemRaw raw,
                  const float* __restrict__ a0, const float* __restrict__ b0, const float* __restrict__ w1,
                  const float* __restrict__ a1, const float* __restrict__ b1, const float* __restrict__ w2,
                  const float* __restrict__ a2, const float* __restrict__ b2, float* __restrict__ y, int64_t y_sb,
                  int64_t y_sc) {
  SMOS_PDL_PROLOGUE();
  extern __shared__ __align__(16) unsigned char stem_raw[];
  using Smem = typename std::conditional<TC, StemSmemTC, StemSmem>::type;
  Smem& S = *reinterpret_cast<Smem*>(stem_raw);
  constexpr int KI = CIN > 0 ? (CIN + 3) / 4 * 4 : kStemCinMax;  // inputs rounded up to whole float4 (zero weights)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // parameters -> shared memory, ONCE per CTA (the CTAs are persistent and walk the 128-point tiles). W2 is kept
  // transposed ([k][c]: the 16 output channels of a warp are contiguous per k); lanes run over c so the transposing
  // stores are conflict free (the 128-bit global reads are strided, 16 KB from L2 per CTA)
  for (int i = tid; i < kStemC * (kStemC / 4); i += kStemThreads) {
    const int c = i & (kStemC - 1), kq = i >> 6;
    const float4 v = __ldg(reinterpret_cast<const float4*>(w2 + c * kStemC) + kq);
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if constexpr (TC) {
        const float hi = tf32_rna(vv[q]);
        S.w2hi[4 * kq + q][c] = hi;
        S.w2lo[4 * kq + q][c] = tf32_rna(vv[q] - hi);
      } else {
        S.w2t[4 * kq + q][c] = vv[q];
      }
    }
  }
  for (int i = tid; i < kStemC * kStemCinMax; i += kStemThreads) {
    const int c = i / kStemCinMax, ci = i % kStemCinMax;
    S.w1[c][ci] = ci < Cin ? __ldg(w1 + c * Cin + ci) : 0.f;  // zero padded rows: whole 128-bit reads below
  }
  if (tid < kStemC) { S.ab1[tid] = make_float2(__ldg(a1 + tid), __ldg(b1 + tid)); S.a2[tid] = __ldg(a2 + tid); S.b2[tid] = __ldg(b2 + tid); }
  if (tid < Cin) { S.a0[tid] = a0 ? __ldg(a0 + tid) : 1.f; S.b0[tid] = b0 ? __ldg(b0 + tid) : 0.f; }
  __syncthreads();

  const int32_t tiles_per_b = (N + kStemPts - 1) / kStemPts;
  const int32_t ntiles = tiles_per_b * B;
  // raw inputs of my point in tile t (one point per thread), fetched one tile ahead of their use
  auto fetch = [&](int32_t t, float (&xr)[KI]) {
    const int32_t b = t / tiles_per_b;
    const int32_t nn = (t - b * tiles_per_b) * kStemPts + tid;
    const int32_t n = min(nn, N - 1);
    if (RAW) {
      const float* p = raw.pts + (static_cast<int64_t>(b) * N + n) * raw.rs;
      float px, py, pz, pw;
      if (raw.rs == 4 && (reinterpret_cast<uintptr_t>(raw.pts) & 15) == 0) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        px = q.x; py = q.y; pz = q.z; pw = q.w;
      } else {
        px = __ldg(p); py = __ldg(p + 1); pz = __ldg(p + 2); pw = __ldg(p + 3);
      }
      px = __fmul_rn(px, raw.sx);
      py = __fmul_rn(py, raw.sy);
      const float qx = __fdiv_rn(__fsub_rn(px, raw.mx), raw.dx);
      const float qy = __fdiv_rn(__fsub_rn(py, raw.my), raw.dy);
      const float qz = __fdiv_rn(__fsub_rn(pz, raw.mz), raw.dz);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz));
      xr[0] = px; xr[1] = py; xr[2] = pz; xr[3] = pw;
      xr[4] = __fadd_rn(__fsqrt_rn(d2), 1e-12f);
      xr[5] = __fsub_rn(qx, floorf(qx));
      xr[6] = __fsub_rn(qy, floorf(qy));
#pragma unroll
      for (int ci = 7; ci < KI; ++ci) xr[ci] = 0.f;
      if (nn < N) {
        float* c = raw.coord + (static_cast<int64_t>(b) * N + n) * 3;
        c[0] = qx; c[1] = qy; c[2] = qz;
      }
      return;
    }
#pragma unroll
    for (int ci = 0; ci < KI; ++ci)
      xr[ci] = ci < Cin ? __ldg(x + b * x_sb + ci * x_sc + static_cast<int64_t>(n) * x_sn) : 0.f;
  };
  float xr[KI];
  int32_t t = blockIdx.x;
  if (t < ntiles) fetch(t, xr);
  for (; t < ntiles; t += gridDim.x) {
    const int32_t b = t / tiles_per_b;
    const int32_t n0 = (t - b * tiles_per_b) * kStemPts;
    // layer 1: thread = point, all hidden channels (weights are warp-uniform broadcast loads)
    {
      float xin[KI];
#pragma unroll
      for (int ci = 0; ci < KI; ++ci) xin[ci] = ci < Cin ? fmaf(xr[ci], S.a0[ci], S.b0[ci]) : 0.f;
#pragma unroll 16  // 16 independent 8-FMA chains in flight: the layer is a dependent-FMA / LDS latency chain otherwise
      for (int c = 0; c < kStemC; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < KI / 4; ++q) {
          const float4 w = *reinterpret_cast<const float4*>(&S.w1[c][4 * q]);
          acc = fmaf(w.x, xin[4 * q], acc); acc = fmaf(w.y, xin[4 * q + 1], acc);
          acc = fmaf(w.z, xin[4 * q + 2], acc); acc = fmaf(w.w, xin[4 * q + 3], acc);
        }
        const float2 ab = S.ab1[c];
        S.h[c][tid] = fmaxf(fmaf(acc, ab.x, ab.y), 0.f);
      }
    }
    if (t + gridDim.x < ntiles) fetch(t + gridDim.x, xr);  // in flight during layer 2
    __syncthreads();
    if constexpr (TC) {
      // layer 2 on the tensor cores: warp = 32 points (2 m-tiles of 16) x all 64 output channels (8 n-tiles of 8),
      // K = 64 in 8 steps; A = hidden activations (row = point, col = k), B = W2^T (row = k, col = channel)
      const int g = lane >> 2, tq = lane & 3;
      const int p0 = wid * 32;
      float acc[2][8][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;
#pragma unroll 2
      for (int ks = 0; ks < 8; ++ks) {
        const int k0 = ks * 8;
        float ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int m = p0 + mt * 16 + g;
          const float a[4] = {S.h[k0 + tq][m], S.h[k0 + tq][m + 8], S.h[k0 + tq + 4][m], S.h[k0 + tq + 4][m + 8]};
#pragma unroll
          for (int r = 0; r < 4; ++r) { ahi[mt][r] = tf32_rna(a[r]); alo[mt][r] = tf32_rna(a[r] - ahi[mt][r]); }
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int n = nt * 8 + g;
          const float bh0 = S.w2hi[k0 + tq][n], bh1 = S.w2hi[k0 + tq + 4][n];
          const float bl0 = S.w2lo[k0 + tq][n], bl1 = S.w2lo[k0 + tq + 4][n];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma_tf32(acc[mt][nt], alo[mt], bh0, bh1);  // small terms first
            mma_tf32(acc[mt][nt], ahi[mt], bl0, bl1);
            mma_tf32(acc[mt][nt], ahi[mt], bh0, bh1);
          }
        }
      }
      // epilogue: c0:(point g, channel 2t) c1:(g, 2t+1) c2:(g+8, 2t) c3:(g+8, 2t+1) of every 16 x 8 tile
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * tq;
        const float al0 = S.a2[c], be0 = S.b2[c], al1 = S.a2[c + 1], be1 = S.b2[c + 1];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int32_t n = n0 + p0 + mt * 16 + g;
          float* d0 = y + b * y_sb + c * y_sc + n;
          float* d1 = d0 + y_sc;
          if (n < N) { d0[0] = fmaxf(fmaf(acc[mt][nt][0], al0, be0), 0.f); d1[0] = fmaxf(fmaf(acc[mt][nt][1], al1, be1), 0.f); }
          if (n + 8 < N) { d0[8] = fmaxf(fmaf(acc[mt][nt][2], al0, be0), 0.f); d1[8] = fmaxf(fmaf(acc[mt][nt][3], al1, be1), 0.f); }
        }
      }
    } else {
      // layer 2: warp = 16 output channels, lane = 4 consecutive points: 64 FMAs per 5 shared-memory loads (8
      // wavefronts) — with 8 channels per warp the kernel sat at 79 % of the shared-memory pipe and 52 % of the FMA pipe
      constexpr int CW = 16;
      const int c0 = wid * CW;
      // accumulators as channel PAIRS: fma.rn.f32x2 (sm_100 packed fp32 FMA, two round-to-nearest FMAs per instruction:
      // bit-identical to two scalar FMAs) halves the issue slots of the inner loop
      float2 acc2[CW / 2][4];
  #pragma unroll
      for (int j = 0; j < CW / 2; ++j)
  #pragma unroll
        for (int i = 0; i < 4; ++i) acc2[j][i] = make_float2(0.f, 0.f);
  #pragma unroll 4
      for (int k = 0; k < kStemC; ++k) {
        const float4 hv = *reinterpret_cast<const float4*>(&S.h[k][lane * 4]);
        const float2 hh[4] = {make_float2(hv.x, hv.x), make_float2(hv.y, hv.y), make_float2(hv.z, hv.z),
                              make_float2(hv.w, hv.w)};
  #pragma unroll
        for (int jq = 0; jq < CW / 4; ++jq) {
          const float4 wv = *reinterpret_cast<const float4*>(&S.w2t[k][c0 + 4 * jq]);
          const float2 wa = make_float2(wv.x, wv.y), wb = make_float2(wv.z, wv.w);
  #pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc2[2 * jq][i] = __ffma2_rn(wa, hh[i], acc2[2 * jq][i]);
            acc2[2 * jq + 1][i] = __ffma2_rn(wb, hh[i], acc2[2 * jq + 1][i]);
          }
        }
      }
      float acc[CW][4];
  #pragma unroll
      for (int j = 0; j < CW / 2; ++j)
  #pragma unroll
        for (int i = 0; i < 4; ++i) { acc[2 * j][i] = acc2[j][i].x; acc[2 * j + 1][i] = acc2[j][i].y; }
      const int32_t n = n0 + lane * 4;
      const bool vec = (n + 3 < N) && ((y_sc & 3) == 0) && ((y_sb & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  #pragma unroll
      for (int j = 0; j < CW; ++j) {
        const int c = c0 + j;
        const float al = S.a2[c], be = S.b2[c];
        float o[4];
  #pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = fmaxf(fmaf(acc[j][i], al, be), 0.f);
        float* dst = y + b * y_sb + c * y_sc + n;
        if (vec) {
          *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
  #pragma unroll
          for (int i = 0; i < 4; ++i)
            if (n + i < N) dst[i] = o[i];
        }
      }
    }
    __syncthreads();  // every warp is done reading h: the next tile's layer 1 may overwrite it
  }
}

#include "point_stem_umma.cuh"

}  // namespace

static int stem_launch(const float* x, int64_t B, int32_t Cin, int64_t N, int64_t x_sb, int64_t x_sc, int64_t x_sn,
                       const StemRaw& raw, const float* bn0_alpha, const float* bn0_beta, const float* w1,
                       const float* bn1_alpha, const float* bn1_beta, const float* w2, const float* bn2_alpha,
                       const float* bn2_beta, int32_t C1, int32_t C2, float* y, int64_t y_sb, int64_t y_sc, int64_t y_sn,
                       void* stream, int32_t max_ctas = 0) {
  if (B <= 0 || N < 0 || Cin <= 0 || max_ctas < 0) return SMOS_EINVAL;
  if (N == 0) return SMOS_OK;
  if ((!x && !raw.pts) || !w1 || !bn1_alpha || !bn1_beta || !w2 || !bn2_alpha || !bn2_beta || !y) return SMOS_EINVAL;
  if ((bn0_alpha == nullptr) != (bn0_beta == nullptr)) return SMOS_EINVAL;
  if (C1 != kStemC || C2 != kStemC || Cin > kStemCinMax || B > 65535 || N >= (int64_t(1) << 31)) return SMOS_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(w2) & 15) != 0) return SMOS_EINVAL;
  const bool is_raw = raw.pts != nullptr;
  if (is_raw && (Cin != 7 || raw.coord == nullptr || raw.rs < 4)) return SMOS_EINVAL;
  // SMOS_STEM_TC=1: layer 2 as 3xTF32 split products on the tensor cores (mma.sync), see StemSmemTC
  const bool tc = smos_env_int("SMOS_STEM_TC", 0) != 0;
  const int32_t Ni = static_cast<int32_t>(N), Bi = static_cast<int32_t>(B);
  cudaStream_t st = smos_stream(stream);
  const int64_t ntiles = static_cast<int64_t>(smos_ceil_div(N, kStemPts)) * B;
  // default: layer 2 on the 5th-generation tensor cores (tcgen05.mma, 3xTF32 split, accumulators in TMEM).
  // SMOS_STEM_UMMA=0 selects the CUDA-core kernel below (bit-identical to the oracle's FMA order; channel-major
  // output only), SMOS_STEM_TC=1 its mma.sync variant.
  if (smos_env_int("SMOS_STEM_UMMA", 1) != 0 && !tc) {
#define SMOS_STEM_UMMA_LAUNCH(CINV, RAWV)                                                                            \
  do {                                                                                                                \
    auto kern = point_stem_umma_kernel<CINV, RAWV>;                                                                   \
    static std::atomic<unsigned long long> opted{0};                                                                  \
    if (cudaError_t e = smos_smem_opt_in(kern, opted, static_cast<int>(sizeof(umma_stem::Smem) + 128)); e != cudaSuccess) \
      return static_cast<int>(e);                                                                                     \
    /* 165 KB of shared memory: one persistent CTA per SM. max_ctas (or SMOS_STEM_CTAS=n) leaves SMs to the kernels of  \
       the other scans in flight: the stem is latency bound (10 % of the DRAM rate) and blocks its SM for everything  \
       else, so a pipelined stream runs it on ~60 % of the SMs */                                        \
    int64_t want = SMOS_SM_COUNT;                                                                                     \
    const int cap = max_ctas > 0 ? max_ctas : smos_env_int("SMOS_STEM_CTAS", 0);                                      \
    if (cap > 0 && cap < want) want = cap;                                                                            \
    dim3 grid(static_cast<unsigned>(ntiles < want ? ntiles : want));                                                  \
    SMOS_LAUNCH((kern), grid, umma_stem::kThreads, sizeof(umma_stem::Smem) + 128, st, x, Cin, Ni, Bi, x_sb, x_sc, x_sn, \
                raw, bn0_alpha, bn0_beta, w1, bn1_alpha, bn1_beta, w2, bn2_alpha, bn2_beta, y, y_sb, y_sc, y_sn);     \
  } while (0)
    if (is_raw) SMOS_STEM_UMMA_LAUNCH(7, true);
    else if (Cin == 7) SMOS_STEM_UMMA_LAUNCH(7, false);
    else SMOS_STEM_UMMA_LAUNCH(0, false);
#undef SMOS_STEM_UMMA_LAUNCH
    return smos_launch_status();
  }
  if (y_sn != 1) return SMOS_EUNSUPPORTED;  // the CUDA-core kernels write channel-major rows
#define SMOS_STEM_LAUNCH(CINV, RAWV, TCV)                                                                             \
  do {                                                                                                                \
    auto kern = point_stem_kernel<CINV, RAWV, TCV>;                                                                   \
    const size_t smem = TCV ? sizeof(StemSmemTC) : sizeof(StemSmem);                                                   \
    static std::atomic<unsigned long long> opted{0};                                                                  \
    if (cudaError_t e = smos_smem_opt_in(kern, opted, static_cast<int>(smem)); e != cudaSuccess)                      \
      return static_cast<int>(e);                                                                                     \
    /* persistent CTAs, one wave, walking the 128-point tiles with a grid stride */                                   \
    int per_sm = 0;                                                                                                   \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStemThreads, smem) != cudaSuccess || per_sm < 1) \
      per_sm = 1;                                                                                                     \
    const int64_t want = static_cast<int64_t>(SMOS_SM_COUNT) * per_sm;                                                \
    dim3 grid(static_cast<unsigned>(ntiles < want ? ntiles : want));                                                  \
    SMOS_LAUNCH((kern), grid, kStemThreads, smem, st, x, Cin, Ni, Bi, x_sb, x_sc, x_sn, raw, bn0_alpha, bn0_beta, w1, bn1_alpha,   \
                                           bn1_beta, w2, bn2_alpha, bn2_beta, y, y_sb, y_sc);                          \
  } while (0)
  if (is_raw) { if (tc) SMOS_STEM_LAUNCH(7, true, true); else SMOS_STEM_LAUNCH(7, true, false); }
  else if (Cin == 7) { if (tc) SMOS_STEM_LAUNCH(7, false, true); else SMOS_STEM_LAUNCH(7, false, false); }
  else { if (tc) SMOS_STEM_LAUNCH(0, false, true); else SMOS_STEM_LAUNCH(0, false, false); }
#undef SMOS_STEM_LAUNCH
  return smos_launch_status();
}

extern "C" int smos_point_stem_forward(const float* x, int64_t B, int32_t Cin, int64_t N, int64_t x_sb, int64_t x_sc,
                                       int64_t x_sn, const float* bn0_alpha, const float* bn0_beta, const float* w1,
                                       const float* bn1_alpha, const float* bn1_beta, const float* w2,
                                       const float* bn2_alpha, const float* bn2_beta, int32_t C1, int32_t C2, float* y,
                                       int64_t y_sb, int64_t y_sc, int64_t y_sn, void* stream) {
  StemRaw raw = {};
  return stem_launch(x, B, Cin, N, x_sb, x_sc, x_sn, raw, bn0_alpha, bn0_beta, w1, bn1_alpha, bn1_beta, w2, bn2_alpha,
                     bn2_beta, C1, C2, y, y_sb, y_sc, y_sn, stream);
}

// smos_form_batch + smos_point_stem_forward in one kernel: raw points in, 64-channel features and the quantised
// coordinates out (same arithmetic as the two calls: bit-identical results).
extern "C" int smos_point_stem_forward_raw(const float* points, int64_t T, int64_t N, int64_t row_stride, float x_sign,
                                           float y_sign, float min_x, float min_y, float min_z, float dx, float dy,
                                           float dz, const float* bn0_alpha, const float* bn0_beta, const float* w1,
                                           const float* bn1_alpha, const float* bn1_beta, const float* w2,
                                           const float* bn2_alpha, const float* bn2_beta, int32_t C1, int32_t C2,
                                           float* pcds_coord, float* y, int64_t y_sb, int64_t y_sc, int64_t y_sn,
                                           void* stream) {
  return smos_point_stem_forward_raw_capped(points, T, N, row_stride, x_sign, y_sign, min_x, min_y, min_z, dx, dy, dz,
                                            bn0_alpha, bn0_beta, w1, bn1_alpha, bn1_beta, w2, bn2_alpha, bn2_beta, C1, C2,
                                            pcds_coord, y, y_sb, y_sc, y_sn, 0, stream);
}

// The same call with the number of persistent CTAs capped (a scheduling hint: results are identical).
extern "C" int smos_point_stem_forward_raw_capped(const float* points, int64_t T, int64_t N, int64_t row_stride,
                                                  float x_sign, float y_sign, float min_x, float min_y, float min_z,
                                                  float dx, float dy, float dz, const float* bn0_alpha,
                                                  const float* bn0_beta, const float* w1, const float* bn1_alpha,
                                                  const float* bn1_beta, const float* w2, const float* bn2_alpha,
                                                  const float* bn2_beta, int32_t C1, int32_t C2, float* pcds_coord,
                                                  float* y, int64_t y_sb, int64_t y_sc, int64_t y_sn, int32_t max_ctas,
                                                  void* stream) {
  if (!points || !pcds_coord) return SMOS_EINVAL;
  StemRaw raw = {points, row_stride, x_sign, y_sign, min_x, min_y, min_z, dx, dy, dz, pcds_coord};
  return stem_launch(nullptr, T, 7, N, 0, 0, 0, raw, bn0_alpha, bn0_beta, w1, bn1_alpha, bn1_beta, w2, bn2_alpha, bn2_beta,
                     C1, C2, y, y_sb, y_sc, y_sn, stream, max_ctas);
}
