// Sampling state of BilinearSample (networks/backbone.py:458-475), shared by the gather kernels and by the pooling
// plan builder, which can emit it as a by-product (one record per slot of the plan's cell order).
#pragma once

#include "common.cuh"

struct Taps {
  int32_t x0, y0;        // north-west integer pixel
  float w_nw, w_ne, w_sw, w_se;
  bool in_nw, in_ne, in_sw, in_se;
};

// pix = ((g + 1) / 2) * (size - 1), g = 2 * c * s / (size - 1) - 1   (backbone.py:469-470 + ATen
// grid_sampler_unnormalize with align_corners=True). Every step individually rounded.
static __device__ __forceinline__ float replay_pixel(float c, float s, int32_t size) {
  const float sm1 = static_cast<float>(size - 1);
  float g = __fmul_rn(__fmul_rn(2.0f, c), s);
  g = __fdiv_rn(g, sm1);
  g = __fsub_rn(g, 1.0f);
  float p = __fadd_rn(g, 1.0f);
  p = __fmul_rn(p, 0.5f);
  return __fmul_rn(p, sm1);
}

static __device__ __forceinline__ Taps make_taps(float cy, float cx, float sh, float sw, int32_t H, int32_t W) {
  Taps t;
  const float ix = replay_pixel(cx, sw, W);
  const float iy = replay_pixel(cy, sh, H);
  const float fx = floorf(ix), fy = floorf(iy);
  // weights as in ATen grid_sampler_2d: nw = (x_se - x)(y_se - y) ...
  const float x1 = __fadd_rn(fx, 1.0f), y1 = __fadd_rn(fy, 1.0f);
  t.w_nw = __fmul_rn(__fsub_rn(x1, ix), __fsub_rn(y1, iy));
  t.w_ne = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(y1, iy));
  t.w_sw = __fmul_rn(__fsub_rn(x1, ix), __fsub_rn(iy, fy));
  t.w_se = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(iy, fy));
  // clamp before the int cast so far-away pads (and NaN) become plain out-of-bounds
  const float cxl = fminf(fmaxf(fx, -2.0f), static_cast<float>(W) + 1.0f);
  const float cyl = fminf(fmaxf(fy, -2.0f), static_cast<float>(H) + 1.0f);
  t.x0 = (fx == fx) ? static_cast<int32_t>(cxl) : -2;
  t.y0 = (fy == fy) ? static_cast<int32_t>(cyl) : -2;
  const bool xl = t.x0 >= 0 && t.x0 < W, xh = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  const bool yl = t.y0 >= 0 && t.y0 < H, yh = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  t.in_nw = xl && yl; t.in_ne = xh && yl; t.in_sw = xl && yh; t.in_se = xh && yh;
  return t;
}

// Per-point sampling state as the gather kernels consume it: 48 bytes, three 128-bit words.
struct __align__(16) TapsS {
  int32_t o_nw, o_ne, o_sw, o_se;  // clamped (always valid) pixel offsets  y * W + x
  float w_nw, w_ne, w_sw, w_se;    // bilinear weights (out-of-image taps are zeroed through in_mask)
  uint32_t in_mask;                // bit0 nw, bit1 ne, bit2 sw, bit3 se
  int32_t n, b;                    // the point this slot samples for (n < 0: no point)
  int32_t pad;
};

static __device__ __forceinline__ TapsS make_taps_record(float cy, float cx, float sh, float sw, int32_t H, int32_t W,
                                                  int32_t n, int32_t b) {
  const Taps t = make_taps(cy, cx, sh, sw, H, W);
  const int32_t xa = min(max(t.x0, 0), W - 1), xb = min(max(t.x0 + 1, 0), W - 1);
  const int32_t ya = min(max(t.y0, 0), H - 1), yb = min(max(t.y0 + 1, 0), H - 1);
  TapsS r;
  r.o_nw = ya * W + xa; r.o_ne = ya * W + xb; r.o_sw = yb * W + xa; r.o_se = yb * W + xb;
  r.w_nw = t.w_nw; r.w_ne = t.w_ne; r.w_sw = t.w_sw; r.w_se = t.w_se;
  r.in_mask = (t.in_nw ? 1u : 0u) | (t.in_ne ? 2u : 0u) | (t.in_sw ? 4u : 0u) | (t.in_se ? 8u : 0u);
  r.n = n;
  r.b = b;
  r.pad = 0;
  return r;
}
