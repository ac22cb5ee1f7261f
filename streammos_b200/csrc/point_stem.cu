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

// RAW: the 7 input channels are built on the fly from raw points (x, y, z, intensity): the loader's Quantize +
// make_point_feat (see form_batch.cu, same float32 sequence, bit-exact), and the (B, N, 3) quantised coordinates leave
// as a side output — the (B, 7, N) tensor never exists.
struct StemRaw {
  const float* pts;  // (B*N, rs) raw points, or null
  int64_t rs;
  float sx, sy, mx, my, mz, dx, dy, dz;
  float* coord;      // (B, N, 3) out
};

template <int CIN, bool RAW>
__global__ void __launch_bounds__(kStemThreads, SMOS_STEM_MIN_CTAS)
point_stem_kernel(const float* __restrict__ x, int32_t Cin, int32_t N, int32_t B, int64_t x_sb, int64_t x_sc, int64_t x_sn,
                  const __grid_constant__ StemRaw raw,
                  const float* __restrict__ a0, const float* __restrict__ b0, const float* __restrict__ w1,
                  const float* __restrict__ a1, const float* __restrict__ b1, const float* __restrict__ w2,
                  const float* __restrict__ a2, const float* __restrict__ b2, float* __restrict__ y, int64_t y_sb,
                  int64_t y_sc) {
  extern __shared__ __align__(16) unsigned char stem_raw[];
  StemSmem& S = *reinterpret_cast<StemSmem*>(stem_raw);
  constexpr int KI = CIN > 0 ? (CIN + 3) / 4 * 4 : kStemCinMax;  // inputs rounded up to whole float4 (zero weights)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // parameters -> shared memory, ONCE per CTA (the CTAs are persistent and walk the 128-point tiles). W2 is kept
  // transposed ([k][c]: the 16 output channels of a warp are contiguous per k); lanes run over c so the transposing
  // stores are conflict free (the 128-bit global reads are strided, 16 KB from L2 per CTA)
  for (int i = tid; i < kStemC * (kStemC / 4); i += kStemThreads) {
    const int c = i & (kStemC - 1), kq = i >> 6;
    const float4 v = __ldg(reinterpret_cast<const float4*>(w2 + c * kStemC) + kq);
    S.w2t[4 * kq][c] = v.x; S.w2t[4 * kq + 1][c] = v.y; S.w2t[4 * kq + 2][c] = v.z; S.w2t[4 * kq + 3][c] = v.w;
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
#pragma unroll 4
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
    __syncthreads();  // every warp is done reading h: the next tile's layer 1 may overwrite it
  }
}

}  // namespace

static int stem_launch(const float* x, int64_t B, int32_t Cin, int64_t N, int64_t x_sb, int64_t x_sc, int64_t x_sn,
                       const StemRaw& raw, const float* bn0_alpha, const float* bn0_beta, const float* w1,
                       const float* bn1_alpha, const float* bn1_beta, const float* w2, const float* bn2_alpha,
                       const float* bn2_beta, int32_t C1, int32_t C2, float* y, int64_t y_sb, int64_t y_sc, void* stream) {
  if (B <= 0 || N < 0 || Cin <= 0) return SMOS_EINVAL;
  if (N == 0) return SMOS_OK;
  if ((!x && !raw.pts) || !w1 || !bn1_alpha || !bn1_beta || !w2 || !bn2_alpha || !bn2_beta || !y) return SMOS_EINVAL;
  if ((bn0_alpha == nullptr) != (bn0_beta == nullptr)) return SMOS_EINVAL;
  if (C1 != kStemC || C2 != kStemC || Cin > kStemCinMax || B > 65535 || N >= (int64_t(1) << 31)) return SMOS_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(w2) & 15) != 0) return SMOS_EINVAL;
  const bool is_raw = raw.pts != nullptr;
  if (is_raw && (Cin != 7 || raw.coord == nullptr || raw.rs < 4)) return SMOS_EINVAL;
  static bool opt_in[64] = {};
  int device = 0;
  cudaGetDevice(&device);
  if (device >= 0 && device < 64 && !opt_in[device]) {
    cudaError_t e = cudaFuncSetAttribute(point_stem_kernel<7, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(StemSmem)));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(point_stem_kernel<7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(StemSmem)));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(point_stem_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(StemSmem)));
    if (e != cudaSuccess) return static_cast<int>(e);
    opt_in[device] = true;
  }
  // persistent CTAs, one wave (3 per SM: 53 KB of shared memory and ~140 registers each), walking the 128-point tiles
  const int64_t ntiles = static_cast<int64_t>(smos_ceil_div(N, kStemPts)) * B;
  int per_sm = 0;
  cudaError_t oe = is_raw ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, point_stem_kernel<7, true>, kStemThreads, sizeof(StemSmem))
                  : Cin == 7 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, point_stem_kernel<7, false>, kStemThreads, sizeof(StemSmem))
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, point_stem_kernel<0, false>, kStemThreads, sizeof(StemSmem));
  if (oe != cudaSuccess || per_sm < 1) per_sm = 1;
  const int64_t want = static_cast<int64_t>(SMOS_SM_COUNT) * per_sm;
  dim3 grid(static_cast<unsigned>(ntiles < want ? ntiles : want));
  const int32_t Ni = static_cast<int32_t>(N), Bi = static_cast<int32_t>(B);
  cudaStream_t st = smos_stream(stream);
  if (is_raw)
    point_stem_kernel<7, true><<<grid, kStemThreads, sizeof(StemSmem), st>>>(
        x, Cin, Ni, Bi, x_sb, x_sc, x_sn, raw, bn0_alpha, bn0_beta, w1, bn1_alpha, bn1_beta, w2, bn2_alpha, bn2_beta, y, y_sb, y_sc);
  else if (Cin == 7)  // the StreamMOS stem: x, y, z, intensity, dist, diff_x, diff_y
    point_stem_kernel<7, false><<<grid, kStemThreads, sizeof(StemSmem), st>>>(
        x, Cin, Ni, Bi, x_sb, x_sc, x_sn, raw, bn0_alpha, bn0_beta, w1, bn1_alpha, bn1_beta, w2, bn2_alpha, bn2_beta, y, y_sb, y_sc);
  else
    point_stem_kernel<0, false><<<grid, kStemThreads, sizeof(StemSmem), st>>>(
        x, Cin, Ni, Bi, x_sb, x_sc, x_sn, raw, bn0_alpha, bn0_beta, w1, bn1_alpha, bn1_beta, w2, bn2_alpha, bn2_beta, y, y_sb, y_sc);
  return smos_launch_status();
}

extern "C" int smos_point_stem_forward(const float* x, int64_t B, int32_t Cin, int64_t N, int64_t x_sb, int64_t x_sc,
                                       int64_t x_sn, const float* bn0_alpha, const float* bn0_beta, const float* w1,
                                       const float* bn1_alpha, const float* bn1_beta, const float* w2,
                                       const float* bn2_alpha, const float* bn2_beta, int32_t C1, int32_t C2, float* y,
                                       int64_t y_sb, int64_t y_sc, void* stream) {
  StemRaw raw = {};
  return stem_launch(x, B, Cin, N, x_sb, x_sc, x_sn, raw, bn0_alpha, bn0_beta, w1, bn1_alpha, bn1_beta, w2, bn2_alpha,
                     bn2_beta, C1, C2, y, y_sb, y_sc, stream);
}

// smos_form_batch + smos_point_stem_forward in one kernel: raw points in, 64-channel features and the quantised
// coordinates out (same arithmetic as the two calls: bit-identical results).
extern "C" int smos_point_stem_forward_raw(const float* points, int64_t T, int64_t N, int64_t row_stride, float x_sign,
                                           float y_sign, float min_x, float min_y, float min_z, float dx, float dy,
                                           float dz, const float* bn0_alpha, const float* bn0_beta, const float* w1,
                                           const float* bn1_alpha, const float* bn1_beta, const float* w2,
                                           const float* bn2_alpha, const float* bn2_beta, int32_t C1, int32_t C2,
                                           float* pcds_coord, float* y, int64_t y_sb, int64_t y_sc, void* stream) {
  if (!points || !pcds_coord) return SMOS_EINVAL;
  StemRaw raw = {points, row_stride, x_sign, y_sign, min_x, min_y, min_z, dx, dy, dz, pcds_coord};
  return stem_launch(nullptr, T, 7, N, 0, 0, 0, raw, bn0_alpha, bn0_beta, w1, bn1_alpha, bn1_beta, w2, bn2_alpha, bn2_beta,
                     C1, C2, y, y_sb, y_sc, stream);
}
