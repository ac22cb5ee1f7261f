// Device-side build of the model's point tensors from raw scans (the part that is exact).
//
// Replaces, for the val loader's form_batch (datasets/data_StreamMOS.py:471-493), the numpy code that turns the raw,
// range-filtered and padded points of T frames into the tensors the network consumes:
//   utils.Quantize (datasets/utils.py:151-169)            -> pcds_coord  (T, N, 3, 1)  (x_quan, y_quan, z_quan)
//   make_point_feat (data_StreamMOS.py:25-50)             -> pcds_xyzi   (T, 7, N, 1)  (x, y, z, intensity, dist,
//                                                                                        diff_x, diff_y)
// with the TTA flips of form_batch_tta (:495-513) as sign arguments. Every float32 operation of the reference is
// IEEE-exact and replayed one to one (subtract, divide, multiply, add, sqrt, floor), so the outputs are BIT-EXACT.
// utils.SphereQuantize (arctan2 / arcsin in numpy float32) is NOT rebuilt here: no device libm matches numpy's
// results bit for bit, and a 1-ulp difference can move a point across a range-view cell boundary; the range-view
// coordinates of the current frame (the only ones the model reads, models/StreamMOS.py:99) stay a loader output.
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

}  // namespace

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
