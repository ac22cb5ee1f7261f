// Multi-scale deformable-attention sampling core for sm_100a (forward + backward).
//
// Replaces deformattn/src/cuda/ms_deform_im2col_cuda.cuh:237-299 (forward, one thread per
// output scalar) and :301-403 (backward for D=32: one 32-thread block per (b,q,m), serial
// shared-memory reduction by thread 0, scalar atomics). Here a group of D/4 lanes owns one
// (b, q, m): every tap is a 16-byte load per lane (a full 128-byte line per group at D=32),
// the per-sample gradients are reduced with warp shuffles, and sampling locations/weights are
// loaded once per group. Semantics are the reference's: pixel = loc * size - 0.5, a sample
// contributes iff -1 < pixel < size in both axes, per-tap zero padding.
#include "common.cuh"

namespace {

constexpr int kMsdaThreads = 256;

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

template <typename T>
struct Bilinear {
  int h_low, w_low;
  T lh, lw, hh, hw;
  bool in1, in2, in3, in4;  // (low,low) (low,high) (high,low) (high,high)
};

template <typename T>
__device__ __forceinline__ Bilinear<T> make_bilinear(T h, T w, int H, int W) {
  Bilinear<T> s;
  s.h_low = static_cast<int>(floor(h));
  s.w_low = static_cast<int>(floor(w));
  s.lh = h - s.h_low;
  s.lw = w - s.w_low;
  s.hh = 1 - s.lh;
  s.hw = 1 - s.lw;
  const bool hl = s.h_low >= 0, wl = s.w_low >= 0;
  const bool hh_ok = s.h_low + 1 <= H - 1, wh_ok = s.w_low + 1 <= W - 1;
  s.in1 = hl && wl; s.in2 = hl && wh_ok; s.in3 = hh_ok && wl; s.in4 = hh_ok && wh_ok;
  return s;
}

// ------------------------------- forward -----------------------------------------------
// VEC channels per lane (4 when D % 4 == 0, else 1); LPG lanes per (b,q,m) group (power of 2).
//
// The op is 5 MB and L2 resident: what it costs is its chain of dependent round trips. The first version walked the
// L*P samples one after the other (location -> four taps -> next location ...: 2 L P round trips, 6.5 us per layer at
// L*P = 4). Here the group first spreads its samples over its lanes — lane j fetches location and weight of sample
// j and works out its four clamped tap offsets and masked corner weights — and then every lane takes the samples
// back by shuffle, SB at a time, with all 4 SB tap loads of a batch in flight: two round trips per batch.
//
// FUSED: the kernel starts one step earlier, at what the attention module computes in front of the sampling core
// (deformattn/modules/ms_deform_attn.py:96-108): `loc` holds the raw sampling offsets and `attn` the attention
// logits; the softmax over the L*P logits of a (b, q, m) and sampling_locations = reference_points + offsets /
// (W_l, H_l) (or the 4-d reference box form) happen in registers, replacing five elementwise torch kernels and
// their (B, Lq, M, L, P, 2) intermediates.
template <typename T>
struct SampleState {
  int32_t o1, o2, o3, o4;  // clamped tap offsets (spatial positions from the start of `value`); o1 < 0: sample unused
  int32_t in;              // bit k: tap k lies inside the map (the others read a clamped pixel and are zeroed)
  T u1, u2, u3, u4, wgt;
};

template <typename T, int VEC, typename I, bool FUSED, int SB>
__global__ void __launch_bounds__(kMsdaThreads)
msda_forward_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const T* __restrict__ loc,
                    const T* __restrict__ attn, int32_t S, int32_t M, int32_t D, int32_t L,
                    int32_t Q, int32_t P, int64_t n_groups, int32_t lpg, T* __restrict__ out,
                    const T* __restrict__ ref, int32_t ref_dim) {
  SMOS_PDL_PROLOGUE();
  // I = int32_t when every tensor has < 2^31 elements (always true for StreamMOS): halves the integer
  // instruction count of the address arithmetic, which dominated this latency-bound kernel
  const I tid = static_cast<I>(blockIdx.x) * kMsdaThreads + threadIdx.x;
  const I grp_raw = tid / lpg;  // (b*Q + q)*M + m
  const int32_t gl = static_cast<int32_t>(tid - grp_raw * lpg);
  // groups are lane aligned and never straddle a warp; lanes of groups past the end stay in the shuffles
  const bool live = grp_raw < static_cast<I>(n_groups);
  const I grp = live ? grp_raw : 0;
  const int32_t m = static_cast<int32_t>(grp % M);
  const I bq = grp / M;
  const int32_t b = static_cast<int32_t>(bq / Q);
  const I row = static_cast<I>(M) * D;  // elements per spatial position
  const T* vb = value + static_cast<I>(b) * S * row + static_cast<I>(m) * D;
  const int32_t LP = L * P;
  const T* lp = loc + grp * (LP * 2);
  const T* ap = attn + grp * LP;
  const int lane = threadIdx.x & 31;
  const int gbase = lane - gl;  // first lane of my group
  T sm_max = 0, sm_inv = 1;
  if (FUSED) {  // softmax statistics of my (b, q, m): every lane of the group walks the (cached) logits
    sm_max = ap[0];
    for (int32_t i = 1; i < LP; ++i) sm_max = max(sm_max, ap[i]);
    T sum = 0;
    for (int32_t i = 0; i < LP; ++i) sum += exp(ap[i] - sm_max);
    sm_inv = T(1) / sum;
  }
  const int32_t dvec = D / VEC;
  for (int32_t dv0 = 0; dv0 < dvec; dv0 += lpg) {  // one pass unless D > 4 * 32
    const int32_t dv = dv0 + gl;
    const bool d_ok = dv < dvec;
    const int32_t d0 = (d_ok ? dv : 0) * VEC;
    T acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0;
    for (int32_t s0 = 0; s0 < LP; s0 += lpg) {
      // ---- lane gl prepares sample s0 + gl ----
      SampleState<T> st;
      st.o1 = st.o2 = st.o3 = st.o4 = -1;
      st.in = 0;
      st.u1 = st.u2 = st.u3 = st.u4 = st.wgt = 0;
      int32_t lvl_off = 0;
      const int32_t si = s0 + gl;
      if (si < LP) {
        const int32_t l = si / P;
        const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
        lvl_off = static_cast<int32_t>(lsi[l]);
        T loc_w = lp[si * 2], loc_h = lp[si * 2 + 1];
        T wgt = ap[si];
        if (FUSED) {
          wgt = exp(wgt - sm_max) * sm_inv;
          const T* rp = ref + (bq * L + l) * ref_dim;
          if (ref_dim == 2) {  // reference_points + offsets / (W_l, H_l)
            loc_w = rp[0] + loc_w / static_cast<T>(W);
            loc_h = rp[1] + loc_h / static_cast<T>(H);
          } else {             // reference boxes: xy + offsets / n_points * wh * 0.5
            loc_w = rp[0] + loc_w / static_cast<T>(P) * rp[2] * T(0.5);
            loc_h = rp[1] + loc_h / static_cast<T>(P) * rp[3] * T(0.5);
          }
        }
        const T h_im = loc_h * H - T(0.5);
        const T w_im = loc_w * W - T(0.5);
        if (h_im > -1 && w_im > -1 && h_im < H && w_im < W) {
          const Bilinear<T> s = make_bilinear<T>(h_im, w_im, H, W);
          const T w1 = s.hh * s.hw, w2 = s.hh * s.lw, w3 = s.lh * s.hw, w4 = s.lh * s.lw;
          // taps outside the map read a clamped (valid) pixel and get weight 0 via a select: unconditional
          // loads stay independent, predicated ones are serialised through one register by ptxas
          const int hl = max(s.h_low, 0), hh = min(s.h_low + 1, H - 1);
          const int wl = max(s.w_low, 0), wh = min(s.w_low + 1, W - 1);
          st.o1 = lvl_off + hl * W + wl; st.o2 = lvl_off + hl * W + wh;
          st.o3 = lvl_off + hh * W + wl; st.o4 = lvl_off + hh * W + wh;
          st.u1 = s.in1 ? w1 : T(0); st.u2 = s.in2 ? w2 : T(0); st.u3 = s.in3 ? w3 : T(0); st.u4 = s.in4 ? w4 : T(0);
          st.wgt = wgt;
          st.in = (s.in1 ? 1 : 0) | (s.in2 ? 2 : 0) | (s.in3 ? 4 : 0) | (s.in4 ? 8 : 0);
        }
      }
      // ---- every lane consumes the group's samples, SB at a time ----
      const int32_t ns = min(lpg, LP - s0);
      for (int32_t j0 = 0; j0 < ns; j0 += SB) {
        SampleState<T> q[SB];
#pragma unroll
        for (int u = 0; u < SB; ++u) {
          const int src = gbase + min(j0 + u, ns - 1);
          q[u].o1 = __shfl_sync(0xffffffffu, st.o1, src); q[u].o2 = __shfl_sync(0xffffffffu, st.o2, src);
          q[u].o3 = __shfl_sync(0xffffffffu, st.o3, src); q[u].o4 = __shfl_sync(0xffffffffu, st.o4, src);
          q[u].u1 = __shfl_sync(0xffffffffu, st.u1, src); q[u].u2 = __shfl_sync(0xffffffffu, st.u2, src);
          q[u].u3 = __shfl_sync(0xffffffffu, st.u3, src); q[u].u4 = __shfl_sync(0xffffffffu, st.u4, src);
          q[u].wgt = __shfl_sync(0xffffffffu, st.wgt, src);
          q[u].in = __shfl_sync(0xffffffffu, st.in, src);
          if (j0 + u >= ns) q[u].o1 = -1;  // padding of the last batch
        }
        if constexpr (VEC == 4) {
          using V = typename Vec4<T>::type;
          V t[SB][4];
#pragma unroll
          for (int u = 0; u < SB; ++u) {  // unused samples re-read position 0 of the map: valid, weight 0
            const bool use = q[u].o1 >= 0;
            t[u][0] = *reinterpret_cast<const V*>(vb + static_cast<I>(use ? q[u].o1 : 0) * row + d0);
            t[u][1] = *reinterpret_cast<const V*>(vb + static_cast<I>(use ? q[u].o2 : 0) * row + d0);
            t[u][2] = *reinterpret_cast<const V*>(vb + static_cast<I>(use ? q[u].o3 : 0) * row + d0);
            t[u][3] = *reinterpret_cast<const V*>(vb + static_cast<I>(use ? q[u].o4 : 0) * row + d0);
          }
#pragma unroll
          for (int u = 0; u < SB; ++u) {
            if (q[u].o1 < 0) continue;  // sample outside the map (the reference skips it, no 0 * inf)
            V z; z.x = z.y = z.z = z.w = 0;
            const V a1 = (q[u].in & 1) ? t[u][0] : z, a2 = (q[u].in & 2) ? t[u][1] : z;
            const V a3 = (q[u].in & 4) ? t[u][2] : z, a4 = (q[u].in & 8) ? t[u][3] : z;
            acc[0] += (q[u].u1 * a1.x + q[u].u2 * a2.x + q[u].u3 * a3.x + q[u].u4 * a4.x) * q[u].wgt;
            acc[1] += (q[u].u1 * a1.y + q[u].u2 * a2.y + q[u].u3 * a3.y + q[u].u4 * a4.y) * q[u].wgt;
            acc[2] += (q[u].u1 * a1.z + q[u].u2 * a2.z + q[u].u3 * a3.z + q[u].u4 * a4.z) * q[u].wgt;
            acc[3] += (q[u].u1 * a1.w + q[u].u2 * a2.w + q[u].u3 * a3.w + q[u].u4 * a4.w) * q[u].wgt;
          }
        } else {
          T t[SB][4];
#pragma unroll
          for (int u = 0; u < SB; ++u) {
            const bool use = q[u].o1 >= 0;
            t[u][0] = vb[static_cast<I>(use ? q[u].o1 : 0) * row + d0];
            t[u][1] = vb[static_cast<I>(use ? q[u].o2 : 0) * row + d0];
            t[u][2] = vb[static_cast<I>(use ? q[u].o3 : 0) * row + d0];
            t[u][3] = vb[static_cast<I>(use ? q[u].o4 : 0) * row + d0];
          }
#pragma unroll
          for (int u = 0; u < SB; ++u) {
            if (q[u].o1 < 0) continue;
            const T a1 = (q[u].in & 1) ? t[u][0] : T(0), a2 = (q[u].in & 2) ? t[u][1] : T(0);
            const T a3 = (q[u].in & 4) ? t[u][2] : T(0), a4 = (q[u].in & 8) ? t[u][3] : T(0);
            acc[0] += (q[u].u1 * a1 + q[u].u2 * a2 + q[u].u3 * a3 + q[u].u4 * a4) * q[u].wgt;
          }
        }
      }
    }
    if (live && d_ok) {
      T* o = out + grp * D + d0;
      if constexpr (VEC == 4) {
        using V = typename Vec4<T>::type;
        V r; r.x = acc[0]; r.y = acc[1]; r.z = acc[2]; r.w = acc[3];
        *reinterpret_cast<V*>(o) = r;
      } else {
        o[0] = acc[0];
      }
    }
  }
}

// ------------------------------- backward ----------------------------------------------
// One warp-aligned group of `lpg` lanes per (b,q,m); lanes stride over channels. For every
// sample the group reduces d/d(loc) and d/d(attn) with shuffles; lane 0 stores them (each
// (b,q,m,l,p) is owned by exactly one group, so plain stores). grad_value uses atomics.
template <typename T>
__global__ void __launch_bounds__(kMsdaThreads)
msda_backward_kernel(const T* __restrict__ gout, const T* __restrict__ value,
                     const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                     const T* __restrict__ loc, const T* __restrict__ attn, int32_t S, int32_t M,
                     int32_t D, int32_t L, int32_t Q, int32_t P, int64_t n_groups, int32_t lpg,
                     T* __restrict__ gvalue, T* __restrict__ gloc, T* __restrict__ gattn) {
  SMOS_PDL_PROLOGUE();
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kMsdaThreads + threadIdx.x;
  const int64_t grp = tid / lpg;
  const int32_t gl = static_cast<int32_t>(tid - grp * lpg);
  const bool active = grp < n_groups;
  const int64_t g = active ? grp : 0;
  const int32_t m = static_cast<int32_t>(g % M);
  const int32_t b = static_cast<int32_t>((g / M) / Q);
  const int64_t row = static_cast<int64_t>(M) * D;
  const int64_t voff = static_cast<int64_t>(b) * S * row + static_cast<int64_t>(m) * D;
  const T* lp = loc + g * L * P * 2;
  const T* ap = attn + g * L * P;
  const T* go = gout + g * D;
  for (int32_t l = 0; l < L; ++l) {
    const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
    const int64_t loff = voff + lsi[l] * row;
    for (int32_t p = 0; p < P; ++p) {
      const T loc_w = lp[(l * P + p) * 2], loc_h = lp[(l * P + p) * 2 + 1];
      const T wgt = ap[l * P + p];
      const T h_im = loc_h * H - T(0.5);
      const T w_im = loc_w * W - T(0.5);
      T g_w = 0, g_h = 0, g_a = 0;
      if (active && h_im > -1 && w_im > -1 && h_im < H && w_im < W) {
        const Bilinear<T> s = make_bilinear<T>(h_im, w_im, H, W);
        const T w1 = s.hh * s.hw, w2 = s.hh * s.lw, w3 = s.lh * s.hw, w4 = s.lh * s.lw;
        const int64_t o1 = loff + (static_cast<int64_t>(s.h_low) * W + s.w_low) * row;
        const int64_t o2 = o1 + row, o3 = o1 + static_cast<int64_t>(W) * row, o4 = o3 + row;
        for (int32_t d = gl; d < D; d += lpg) {
          const T top = go[d];
          const T tgv = top * wgt;
          T gh = 0, gw = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0;
          if (s.in1) { v1 = value[o1 + d]; gh -= s.hw * v1; gw -= s.hh * v1; atomicAdd(gvalue + o1 + d, w1 * tgv); }
          if (s.in2) { v2 = value[o2 + d]; gh -= s.lw * v2; gw += s.hh * v2; atomicAdd(gvalue + o2 + d, w2 * tgv); }
          if (s.in3) { v3 = value[o3 + d]; gh += s.hw * v3; gw -= s.lh * v3; atomicAdd(gvalue + o3 + d, w3 * tgv); }
          if (s.in4) { v4 = value[o4 + d]; gh += s.lw * v4; gw += s.lh * v4; atomicAdd(gvalue + o4 + d, w4 * tgv); }
          const T val = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
          g_a += top * val;
          g_w += W * gw * tgv;
          g_h += H * gh * tgv;
        }
      }
      // group reduction (lpg is a power of two <= 32 and groups are lane-aligned)
      for (int o = lpg >> 1; o > 0; o >>= 1) {
        g_w += __shfl_xor_sync(0xffffffffu, g_w, o);
        g_h += __shfl_xor_sync(0xffffffffu, g_h, o);
        g_a += __shfl_xor_sync(0xffffffffu, g_a, o);
      }
      if (active && gl == 0) {
        gloc[(g * L * P + l * P + p) * 2] = g_w;
        gloc[(g * L * P + l * P + p) * 2 + 1] = g_h;
        gattn[g * L * P + l * P + p] = g_a;
      }
    }
  }
}

// fp32, D % 4 == 0: the same backward with a lane owning FOUR channels. The reference (and the scalar kernel above)
// issue one 4-byte atomic per tap and channel — 8.4 M atomics per layer at the config shape; here a tap costs one
// 16-byte vector atomic per lane (atomicAdd on float4, sm_90+: RED.E.ADD.F32x4), a quarter of the atomic instructions
// and whole 16-byte pieces at L2, and the value taps are 16-byte loads as in the forward. Groups of D/4 lanes per
// (b, q, m); d/d(loc) and d/d(attn) reduced with shuffles as above.
__global__ void __launch_bounds__(kMsdaThreads)
msda_backward_vec4_kernel(const float* __restrict__ gout, const float* __restrict__ value,
                          const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                          const float* __restrict__ loc, const float* __restrict__ attn, int32_t S, int32_t M,
                          int32_t D, int32_t L, int32_t Q, int32_t P, int64_t n_groups, int32_t lpg,
                          float* __restrict__ gvalue, float* __restrict__ gloc, float* __restrict__ gattn) {
  SMOS_PDL_PROLOGUE();
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kMsdaThreads + threadIdx.x;
  const int64_t grp = tid / lpg;
  const int32_t gl = static_cast<int32_t>(tid - grp * lpg);
  const bool active = grp < n_groups;
  const int64_t g = active ? grp : 0;
  const int32_t m = static_cast<int32_t>(g % M);
  const int32_t b = static_cast<int32_t>((g / M) / Q);
  const int64_t row = static_cast<int64_t>(M) * D;
  const int64_t voff = static_cast<int64_t>(b) * S * row + static_cast<int64_t>(m) * D;
  const float* lp = loc + g * L * P * 2;
  const float* ap = attn + g * L * P;
  const float* go = gout + g * D;
  const int32_t dvec = D >> 2;
  for (int32_t l = 0; l < L; ++l) {
    const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
    const int64_t loff = voff + lsi[l] * row;
    for (int32_t p = 0; p < P; ++p) {
      const float loc_w = lp[(l * P + p) * 2], loc_h = lp[(l * P + p) * 2 + 1];
      const float wgt = ap[l * P + p];
      const float h_im = loc_h * H - 0.5f;
      const float w_im = loc_w * W - 0.5f;
      float g_w = 0, g_h = 0, g_a = 0;
      if (active && h_im > -1 && w_im > -1 && h_im < H && w_im < W) {
        const Bilinear<float> s = make_bilinear<float>(h_im, w_im, H, W);
        const float w1 = s.hh * s.hw, w2 = s.hh * s.lw, w3 = s.lh * s.hw, w4 = s.lh * s.lw;
        const int64_t o1 = loff + (static_cast<int64_t>(s.h_low) * W + s.w_low) * row;
        const int64_t o2 = o1 + row, o3 = o1 + static_cast<int64_t>(W) * row, o4 = o3 + row;
        for (int32_t dv = gl; dv < dvec; dv += lpg) {
          const int32_t d = dv << 2;
          const float4 top = *reinterpret_cast<const float4*>(go + d);
          const float t[4] = {top.x, top.y, top.z, top.w};
          float v[4][4];  // [tap][channel]
          const int64_t o[4] = {o1, o2, o3, o4};
          const bool in[4] = {s.in1, s.in2, s.in3, s.in4};
          const float w[4] = {w1, w2, w3, w4};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in[k]) {
              x = *reinterpret_cast<const float4*>(value + o[k] + d);
              atomicAdd(reinterpret_cast<float4*>(gvalue + o[k] + d),
                        make_float4(w[k] * (t[0] * wgt), w[k] * (t[1] * wgt), w[k] * (t[2] * wgt), w[k] * (t[3] * wgt)));
            }
            v[k][0] = x.x; v[k][1] = x.y; v[k][2] = x.z; v[k][3] = x.w;
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {  // same expressions as the scalar kernel, channel by channel
            const float tgv = t[c] * wgt;
            float gh = 0, gw = 0;
            if (s.in1) { gh -= s.hw * v[0][c]; gw -= s.hh * v[0][c]; }
            if (s.in2) { gh -= s.lw * v[1][c]; gw += s.hh * v[1][c]; }
            if (s.in3) { gh += s.hw * v[2][c]; gw -= s.lh * v[2][c]; }
            if (s.in4) { gh += s.lw * v[3][c]; gw += s.lh * v[3][c]; }
            const float val = w1 * v[0][c] + w2 * v[1][c] + w3 * v[2][c] + w4 * v[3][c];
            g_a += t[c] * val;
            g_w += W * gw * tgv;
            g_h += H * gh * tgv;
          }
        }
      }
      for (int o = lpg >> 1; o > 0; o >>= 1) {
        g_w += __shfl_xor_sync(0xffffffffu, g_w, o);
        g_h += __shfl_xor_sync(0xffffffffu, g_h, o);
        g_a += __shfl_xor_sync(0xffffffffu, g_a, o);
      }
      if (active && gl == 0) {
        gloc[(g * L * P + l * P + p) * 2] = g_w;
        gloc[(g * L * P + l * P + p) * 2 + 1] = g_h;
        gattn[g * L * P + l * P + p] = g_a;
      }
    }
  }
}

int32_t pick_lpg(int32_t dvec) {
  int32_t lpg = 1;
  while (lpg < dvec && lpg < 32) lpg <<= 1;
  return lpg;
}

template <typename T>
int forward_impl(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc, const void* attn,
                 int32_t B, int32_t S, int32_t M, int32_t D, int32_t L, int32_t Q, int32_t P, void* out,
                 cudaStream_t st, const void* ref = nullptr, int32_t ref_dim = 0) {
  const int64_t n_groups = static_cast<int64_t>(B) * Q * M;
  const size_t vbytes = 4 * sizeof(T);
  const bool vec = (D % 4 == 0) && (reinterpret_cast<uintptr_t>(value) % vbytes == 0) &&
                   (reinterpret_cast<uintptr_t>(out) % vbytes == 0);
  const int32_t lpg = pick_lpg(vec ? D / 4 : D);
  const int64_t threads = n_groups * lpg;
  const int grid = smos_ceil_div(threads, kMsdaThreads);
  const bool idx32 = static_cast<int64_t>(B) * S * M * D < (int64_t(1) << 31) &&
                     threads + kMsdaThreads < (int64_t(1) << 31) && n_groups * L * P * 2 < (int64_t(1) << 31);
  if (static_cast<int64_t>(S) >= (int64_t(1) << 31)) return SMOS_EUNSUPPORTED;  // tap offsets are 32-bit positions
  // samples per batch: 4 in fp32 (sixteen 16-byte loads per lane in flight), 2 in fp64 (registers); SMOS_MSDA_SB=2
  // selects the smaller batch for experiments
  const bool sb4 = sizeof(T) == 4 && smos_env_int("SMOS_MSDA_SB", 4) == 4;
#define SMOS_MSDA_FWD(V, I, F)                                                                                      \
  do {                                                                                                              \
    if (sb4)                                                                                                        \
      SMOS_LAUNCH((msda_forward_kernel<T, V, I, F, 4>), grid, kMsdaThreads, 0, st, static_cast<const T*>(value),    \
                  shapes, lsi, static_cast<const T*>(loc), static_cast<const T*>(attn), S, M, D, L, Q, P, n_groups, \
                  lpg, static_cast<T*>(out), static_cast<const T*>(ref), ref_dim);                                  \
    else                                                                                                            \
      SMOS_LAUNCH((msda_forward_kernel<T, V, I, F, 2>), grid, kMsdaThreads, 0, st, static_cast<const T*>(value),    \
                  shapes, lsi, static_cast<const T*>(loc), static_cast<const T*>(attn), S, M, D, L, Q, P, n_groups, \
                  lpg, static_cast<T*>(out), static_cast<const T*>(ref), ref_dim);                                  \
  } while (0)
  if (ref != nullptr) {
    if (vec && idx32) SMOS_MSDA_FWD(4, int32_t, true);
    else if (vec) SMOS_MSDA_FWD(4, int64_t, true);
    else if (idx32) SMOS_MSDA_FWD(1, int32_t, true);
    else SMOS_MSDA_FWD(1, int64_t, true);
  } else {
    if (vec && idx32) SMOS_MSDA_FWD(4, int32_t, false);
    else if (vec) SMOS_MSDA_FWD(4, int64_t, false);
    else if (idx32) SMOS_MSDA_FWD(1, int32_t, false);
    else SMOS_MSDA_FWD(1, int64_t, false);
  }
#undef SMOS_MSDA_FWD
  return smos_launch_status();
}

template <typename T>
int backward_impl(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc, const void* attn,
                  const void* gout, int32_t B, int32_t S, int32_t M, int32_t D, int32_t L, int32_t Q, int32_t P,
                  void* gvalue, void* gloc, void* gattn, cudaStream_t st) {
  const int64_t n_groups = static_cast<int64_t>(B) * Q * M;
  if constexpr (sizeof(T) == 4) {
    auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if ((D & 3) == 0 && a16(value) && a16(gvalue) && a16(gout) && smos_env_int("SMOS_MSDA_BWD_VEC", 1)) {
      const int32_t lpg4 = pick_lpg(D / 4);
      const int64_t threads4 = n_groups * lpg4;
      SMOS_LAUNCH((msda_backward_vec4_kernel), smos_ceil_div(threads4, kMsdaThreads), kMsdaThreads, 0, st,
                  static_cast<const float*>(gout), static_cast<const float*>(value), shapes, lsi,
                  static_cast<const float*>(loc), static_cast<const float*>(attn), S, M, D, L, Q, P, n_groups, lpg4,
                  static_cast<float*>(gvalue), static_cast<float*>(gloc), static_cast<float*>(gattn));
      return smos_launch_status();
    }
  }
  const int32_t lpg = pick_lpg(D);
  const int64_t threads = n_groups * lpg;
  SMOS_LAUNCH((msda_backward_kernel<T>), smos_ceil_div(threads, kMsdaThreads), kMsdaThreads, 0, st, 
      static_cast<const T*>(gout), static_cast<const T*>(value), shapes, lsi, static_cast<const T*>(loc),
      static_cast<const T*>(attn), S, M, D, L, Q, P, n_groups, lpg, static_cast<T*>(gvalue), static_cast<T*>(gloc),
      static_cast<T*>(gattn));
  return smos_launch_status();
}

bool bad_dims(int32_t B, int32_t S, int32_t M, int32_t D, int32_t L, int32_t Q, int32_t P) {
  return B <= 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || Q <= 0 || P <= 0;
}

}  // namespace

extern "C" {

int smos_ms_deform_attn_forward(int32_t dtype, const void* value, const int64_t* spatial_shapes,
                                const int64_t* level_start_index, const void* sampling_loc,
                                const void* attn_weight, int32_t B, int32_t S, int32_t M, int32_t D, int32_t L,
                                int32_t Q, int32_t P, void* output, void* stream) {
  if (bad_dims(B, S, M, D, L, Q, P)) return SMOS_EINVAL;
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output)
    return SMOS_EINVAL;
  if (static_cast<int64_t>(B) * S * M * D >= (int64_t(1) << 40)) return SMOS_EUNSUPPORTED;
  cudaStream_t st = smos_stream(stream);
  if (dtype == SMOS_F32)
    return forward_impl<float>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, B, S, M, D, L,
                               Q, P, output, st);
  if (dtype == SMOS_F64)
    return forward_impl<double>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, B, S, M, D, L,
                                Q, P, output, st);
  return SMOS_EUNSUPPORTED;
}

int smos_ms_deform_attn_fused_forward(int32_t dtype, const void* value, const int64_t* spatial_shapes,
                                      const int64_t* level_start_index, const void* sampling_offsets,
                                      const void* attn_logits, const void* reference_points, int32_t ref_dim,
                                      int32_t B, int32_t S, int32_t M, int32_t D, int32_t L, int32_t Q, int32_t P,
                                      void* output, void* stream) {
  if (bad_dims(B, S, M, D, L, Q, P) || (ref_dim != 2 && ref_dim != 4)) return SMOS_EINVAL;
  if (!value || !spatial_shapes || !level_start_index || !sampling_offsets || !attn_logits || !reference_points ||
      !output)
    return SMOS_EINVAL;
  if (static_cast<int64_t>(B) * S * M * D >= (int64_t(1) << 40)) return SMOS_EUNSUPPORTED;
  cudaStream_t st = smos_stream(stream);
  if (dtype == SMOS_F32)
    return forward_impl<float>(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, B, S, M, D, L,
                               Q, P, output, st, reference_points, ref_dim);
  if (dtype == SMOS_F64)
    return forward_impl<double>(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, B, S, M, D,
                                L, Q, P, output, st, reference_points, ref_dim);
  return SMOS_EUNSUPPORTED;
}

int smos_ms_deform_attn_backward(int32_t dtype, const void* value, const int64_t* spatial_shapes,
                                 const int64_t* level_start_index, const void* sampling_loc,
                                 const void* attn_weight, const void* grad_output, int32_t B, int32_t S,
                                 int32_t M, int32_t D, int32_t L, int32_t Q, int32_t P, void* grad_value,
                                 void* grad_sampling_loc, void* grad_attn_weight, void* stream) {
  if (bad_dims(B, S, M, D, L, Q, P)) return SMOS_EINVAL;
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_output ||
      !grad_value || !grad_sampling_loc || !grad_attn_weight)
    return SMOS_EINVAL;
  cudaStream_t st = smos_stream(stream);
  if (dtype == SMOS_F32)
    return backward_impl<float>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                B, S, M, D, L, Q, P, grad_value, grad_sampling_loc, grad_attn_weight, st);
  if (dtype == SMOS_F64)
    return backward_impl<double>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                 B, S, M, D, L, Q, P, grad_value, grad_sampling_loc, grad_attn_weight, st);
  return SMOS_EUNSUPPORTED;
}

}  // extern "C"
