// Device-side build of the model's point tensors from raw scans (the part that is exact).
//
// Replaces, for the val loader's form_batch (datasets/data_StreamMOS.py:471-493), the numpy code that turns the raw,
// range-filtered and padded points of T frames into the tensors the network consumes:
//   utils.Quantize (datasets/utils.py:151-169)            -> pcds_coord  (T, N, 3, 1)  (x_quan, y_quan, z_quan)
//   make_point_feat (data_StreamMOS.py:25-50)             -> pcds_xyzi   (T, 7, N, 1)  (x, y, z, intensity, dist,
//                                                                                        diff_x, diff_y)
// with the TTA flips of form_batch_tta (:495-513) as sign arguments. Every float32 operation of the reference is
// IEEE-exact and replayed one to one (subtract, divide, multiply, add, sqrt, floor), so the outputs are BIT-EXACT.
// utils.SphereQuantize (datasets/utils.py:172-192; arctan2 / arcsin in numpy float32) is a kernel of its own below
// (smos_sphere_quantize): no device libm matches numpy's float32 arctan2 / arcsin bit for bit (numpy's SVML loops are
// not correctly rounded and differ between hosts), so that one is a FLOATING-POINT result with a stated tolerance —
// the two angles are evaluated in float64 and rounded to float32 once (the correctly rounded float32 value in all but
// ~1e-9 of the cases), every other operation of the reference is replayed in float32. The bit-exact path keeps the
// range-view coordinates of the current frame (the only ones the model reads, models/StreamMOS.py:99) a loader output.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
form_batch_kernel(const float* __restrict__ pts, int64_t total, int64_t N, int64_t rs, float sx, float sy, float mx,
                  float my, float mz, float dx, float dy, float dz, float* __restrict__ feat, float* __restrict__ coord) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total) return;
  const float* p = pts + i * rs;
  float x, y, z, w;
  if (rs == 4 && (reinterpret_cast<uintptr_t>(pts) & 15) == 0) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    x = q.x; y = q.y; z = q.z; w = q.w;
  } else {
    x = p[0]; y = p[1]; z = p[2]; w = p[3];
  }
  x = __fmul_rn(x, sx);  // TTA flips (exact)
  y = __fmul_rn(y, sy);
  const float qx = __fdiv_rn(__fsub_rn(x, mx), dx);
  const float qy = __fdiv_rn(__fsub_rn(y, my), dy);
  const float qz = __fdiv_rn(__fsub_rn(z, mz), dz);
  // dist = sqrt(x**2 + y**2 + z**2) + 1e-12, all float32, numpy's left-to-right sum
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  const float dist = __fadd_rn(__fsqrt_rn(d2), 1e-12f);
  const int64_t t = i / N, n = i - t * N;
  float* f = feat + t * 7 * N + n;  // (T, 7, N): channel planes, points contiguous
  f[0] = x; f[N] = y; f[2 * N] = z; f[3 * N] = w; f[4 * N] = dist;
  f[5 * N] = __fsub_rn(qx, floorf(qx));
  f[6 * N] = __fsub_rn(qy, floorf(qy));
  float* c = coord + i * 3;
  c[0] = qx; c[1] = qy; c[2] = qz;
}

// utils.SphereQuantize: (theta_quan, phi_quan) per point. numpy evaluates everything in float32 (the Python-float
// constants are weak scalars): d = sqrt(x*x + y*y + z*z) + 1e-12; phi = phi_hi - arctan2(x, y); theta = theta_hi -
// arcsin(z / d); each divided by its float32 step.
__global__ void __launch_bounds__(256)
sphere_quantize_kernel(const float* __restrict__ pts, int64_t total, int64_t rs, float sx, float sy, float phi_hi,
                       float theta_hi, float dphi, float dtheta, float2* __restrict__ out) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total) return;
  const float* p = pts + i * rs;
  float x, y, z;
  if (rs == 4 && (reinterpret_cast<uintptr_t>(pts) & 15) == 0) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    x = q.x; y = q.y; z = q.z;
  } else {
    x = p[0]; y = p[1]; z = p[2];
  }
  x = __fmul_rn(x, sx);
  y = __fmul_rn(y, sy);
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  const float dist = __fadd_rn(__fsqrt_rn(d2), 1e-12f);
  const float s = __fdiv_rn(z, dist);
  const float a_phi = static_cast<float>(atan2(static_cast<double>(x), static_cast<double>(y)));
  const float a_theta = static_cast<float>(asin(static_cast<double>(s)));
  const float phi = __fsub_rn(phi_hi, a_phi);
  const float theta = __fsub_rn(theta_hi, a_theta);
  out[i] = make_float2(__fdiv_rn(theta, dtheta), __fdiv_rn(phi, dphi));
}

}  // namespace

extern "C" int smos_sphere_quantize(const float* points, int64_t total, int64_t row_stride, float x_sign, float y_sign,
                                    float phi_hi, float theta_hi, float dphi, float dtheta, float* sphere_coord,
                                    void* stream) {
  if (total < 0 || row_stride < 3 || !(dphi != 0.0f) || !(dtheta != 0.0f)) return SMOS_EINVAL;
  if (total == 0) return SMOS_OK;
  if (!points || !sphere_coord || (reinterpret_cast<uintptr_t>(sphere_coord) & 7) != 0) return SMOS_EINVAL;
  SMOS_LAUNCH((sphere_quantize_kernel), smos_ceil_div(total, 256), 256, 0, smos_stream(stream),
      points, total, row_stride, x_sign, y_sign, phi_hi, theta_hi, dphi, dtheta, reinterpret_cast<float2*>(sphere_coord));
  return smos_launch_status();
}

extern "C" int smos_form_batch(const float* points, int64_t T, int64_t N, int64_t row_stride, float x_sign, float y_sign,
                               float min_x, float min_y, float min_z, float dx, float dy, float dz, float* pcds_xyzi,
                               float* pcds_coord, void* stream) {
  if (T <= 0 || N < 0 || row_stride < 4) return SMOS_EINVAL;
  if (N == 0) return SMOS_OK;
  if (!points || !pcds_xyzi || !pcds_coord) return SMOS_EINVAL;
  const int64_t total = T * N;
  SMOS_LAUNCH((form_batch_kernel), smos_ceil_div(total, 256), 256, 0, smos_stream(stream), 
      points, total, N, row_stride, x_sign, y_sign, min_x, min_y, min_z, dx, dy, dz, pcds_xyzi, pcds_coord);
  return smos_launch_status();
}
