// Device-side scan ingestion (the loader steps in front of form_batch).
//
// Replaces, for every frame of the model's T-frame input window, the numpy code of the val loader
// (datasets/data_StreamMOS.py:515-574):
//   utils.Trans (datasets/utils.py:116-126)            pose alignment: float64 pose_diff . (x, y, z, 1), stored float32
//   utils.filter_pcds_mask (datasets/utils.py:107-113) range filter  lo <= p < hi  on the aligned float32 point
//   pc_list[ht][valid_mask]                            ORDER-PRESERVING compaction
//   np.pad(..., -1000) ; [:, 2] = -4000                padding to frame_point_num rows
// so that a stream keeps the RAW scans of its window resident in HBM: per scan only the new raw scan (and the pose
// chain, 96 bytes per frame) has to cross PCIe, instead of T re-aligned, filtered and padded frames (the two older
// frames change with every new pose, so the host re-uploaded them every scan).
// Exactness: Trans is an FMA chain in k order in float64 — what numpy's dgemm computes for a (4,4).(4,N) product —
// rounded to float32 once; comparisons and copies are exact. Bit-exact against the loader's own functions
// (tests/golden/ingest_a.npz).
// Everything the kernels need from a frame — point count and pose — is read from DEVICE memory, so the launch
// parameters never change and the two kernels sit in a CUDA graph.
#include "common.cuh"

namespace {

constexpr int kIngestThreads = 256;
constexpr int kIngestPts = 4;                                  // consecutive points per thread
constexpr int kIngestTile = kIngestThreads * kIngestPts;       // points per CTA
constexpr int kMaxFrames = 8;

struct IngestArgs {
  const float* pts[kMaxFrames];
  const int32_t* n_dev[kMaxFrames];
  const double* pose[kMaxFrames];  // 12 doubles (rows 0..2 of pose_diff) or null: no transform
  int64_t n_cap[kMaxFrames];
  int64_t rs;
  float lo[3], hi[3];
  int64_t n_out;
  float pad_xy, pad_z;
  int32_t* cta_count;  // [T][tiles]
  int32_t tiles;       // tiles per frame (covers max(n_cap, n_out))
  float* out;          // (T, n_out, 4)
  int32_t* src;        // (T, n_out) or null
  int32_t* count;      // (T)
};

// aligned point + range test of raw row i of frame t (i < n)
__device__ __forceinline__ bool ingest_point(const IngestArgs& A, int t, int64_t i, const double* m, float4* q) {
  const float* p = A.pts[t] + i * A.rs;
  float x = p[0], y = p[1], z = p[2];
  const float w = p[3];
  if (m != nullptr) {
    const double dx = x, dy = y, dz = z;
    x = static_cast<float>(fma(m[3], 1.0, fma(m[2], dz, fma(m[1], dy, m[0] * dx))));
    y = static_cast<float>(fma(m[7], 1.0, fma(m[6], dz, fma(m[5], dy, m[4] * dx))));
    z = static_cast<float>(fma(m[11], 1.0, fma(m[10], dz, fma(m[9], dy, m[8] * dx))));
  }
  *q = make_float4(x, y, z, w);
  return x >= A.lo[0] && x < A.hi[0] && y >= A.lo[1] && y < A.hi[1] && z >= A.lo[2] && z < A.hi[2];
}

// pass 1: number of points of every tile that survive the range filter
__global__ void __launch_bounds__(kIngestThreads)
ingest_count_kernel(const __grid_constant__ IngestArgs A) {
  SMOS_PDL_PROLOGUE();
  const int t = blockIdx.y;
  const int64_t n = min(static_cast<int64_t>(__ldg(A.n_dev[t])), A.n_cap[t]);
  __shared__ double s_m[12];
  if (A.pose[t] != nullptr && threadIdx.x < 12) s_m[threadIdx.x] = A.pose[t][threadIdx.x];
  __syncthreads();
  const double* m = A.pose[t] != nullptr ? s_m : nullptr;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * kIngestTile + threadIdx.x * kIngestPts;
  int mine = 0;
#pragma unroll
  for (int k = 0; k < kIngestPts; ++k) {
    float4 q;
    if (i0 + k < n && ingest_point(A, t, i0 + k, m, &q)) ++mine;
  }
  // block sum: warp shuffles + one shared word per warp
  __shared__ int s_w[kIngestThreads / 32];
  int v = mine;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < kIngestThreads / 32; ++w) s += s_w[w];
    A.cta_count[t * A.tiles + blockIdx.x] = s;
  }
}

// pass 2: positions from the tile counts (sum of the tiles in front + scan inside the tile), rows written in order,
// padding behind them
__global__ void __launch_bounds__(kIngestThreads)
ingest_write_kernel(const __grid_constant__ IngestArgs A) {
  SMOS_PDL_PROLOGUE();
  const int t = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t n = min(static_cast<int64_t>(__ldg(A.n_dev[t])), A.n_cap[t]);
  __shared__ double s_m[12];
  __shared__ int s_w[kIngestThreads / 32];
  __shared__ int s_base, s_total;
  if (A.pose[t] != nullptr && threadIdx.x < 12) s_m[threadIdx.x] = A.pose[t][threadIdx.x];
  // tiles in front of mine and all tiles of the frame: one warp sums the (few hundred) counts
  if (wid == 0) {
    int before = 0, all = 0;
    for (int c = lane; c < A.tiles; c += 32) {
      const int v = A.cta_count[t * A.tiles + c];
      all += v;
      if (c < static_cast<int>(blockIdx.x)) before += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      before += __shfl_xor_sync(0xffffffffu, before, o);
      all += __shfl_xor_sync(0xffffffffu, all, o);
    }
    if (lane == 0) { s_base = before; s_total = all; }
  }
  __syncthreads();
  const double* m = A.pose[t] != nullptr ? s_m : nullptr;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * kIngestTile + threadIdx.x * kIngestPts;
  float4 q[kIngestPts];
  bool ok[kIngestPts];
  int mine = 0;
#pragma unroll
  for (int k = 0; k < kIngestPts; ++k) {
    ok[k] = i0 + k < n && ingest_point(A, t, i0 + k, m, &q[k]);
    mine += ok[k] ? 1 : 0;
  }
  // exclusive scan of `mine` over the CTA
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s_w[wid] = incl;
  __syncthreads();
  int warp_before = 0;
#pragma unroll
  for (int w = 0; w < kIngestThreads / 32; ++w)
    if (w < wid) warp_before += s_w[w];
  int64_t pos = static_cast<int64_t>(s_base) + warp_before + (incl - mine);
  float4* out = reinterpret_cast<float4*>(A.out) + static_cast<int64_t>(t) * A.n_out;
  int32_t* src = A.src != nullptr ? A.src + static_cast<int64_t>(t) * A.n_out : nullptr;
#pragma unroll
  for (int k = 0; k < kIngestPts; ++k) {
    if (ok[k]) {
      if (pos < A.n_out) {  // (the loader asserts pad_length > 0: a frame that does not fit is truncated, `count` says so)
        out[pos] = q[k];
        if (src != nullptr) src[pos] = static_cast<int32_t>(i0 + k);
      }
      ++pos;
    }
  }
  // padding rows [total, n_out): row index = my raw index (tiles cover max(n_cap, n_out))
  const int64_t total = s_total;
#pragma unroll
  for (int k = 0; k < kIngestPts; ++k) {
    const int64_t r = i0 + k;
    if (r >= total && r < A.n_out) {
      out[r] = make_float4(A.pad_xy, A.pad_xy, A.pad_z, A.pad_xy);
      if (src != nullptr) src[r] = -1;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && A.count != nullptr) A.count[t] = s_total;
}

}  // namespace

extern "C" {

int64_t smos_ingest_workspace_bytes(int32_t T, int64_t n_cap_max, int64_t n_out) {
  if (T <= 0 || T > kMaxFrames || n_cap_max < 0 || n_out <= 0) return SMOS_EINVAL;
  const int64_t span = n_cap_max > n_out ? n_cap_max : n_out;
  return smos_align_up(static_cast<int64_t>(T) * smos_ceil_div(span, kIngestTile) * 4, 256);
}

int smos_ingest_frames(const smos_ingest_frame* frames_host, int32_t T, int64_t row_floats, float x_lo, float x_hi,
                       float y_lo, float y_hi, float z_lo, float z_hi, int64_t n_out, float pad_xy, float pad_z,
                       void* workspace, float* out_points, int32_t* out_src, int32_t* out_count, void* stream) {
  if (frames_host == nullptr || T <= 0 || T > kMaxFrames || row_floats < 4 || n_out <= 0) return SMOS_EINVAL;
  if (workspace == nullptr || out_points == nullptr || (reinterpret_cast<uintptr_t>(out_points) & 15) != 0) return SMOS_EINVAL;
  IngestArgs A;
  int64_t span = n_out;
  for (int32_t t = 0; t < T; ++t) {
    const smos_ingest_frame& f = frames_host[t];
    if (f.n_cap < 0 || f.n_dev == nullptr || (f.n_cap > 0 && f.points == nullptr)) return SMOS_EINVAL;
    if (f.n_cap >= (int64_t(1) << 31)) return SMOS_EUNSUPPORTED;
    A.pts[t] = f.points; A.n_dev[t] = f.n_dev; A.pose[t] = f.pose_dev; A.n_cap[t] = f.n_cap;
    if (f.n_cap > span) span = f.n_cap;
  }
  for (int32_t t = T; t < kMaxFrames; ++t) { A.pts[t] = nullptr; A.n_dev[t] = nullptr; A.pose[t] = nullptr; A.n_cap[t] = 0; }
  if (n_out >= (int64_t(1) << 31)) return SMOS_EUNSUPPORTED;
  A.rs = row_floats;
  A.lo[0] = x_lo; A.lo[1] = y_lo; A.lo[2] = z_lo;
  A.hi[0] = x_hi; A.hi[1] = y_hi; A.hi[2] = z_hi;
  A.n_out = n_out; A.pad_xy = pad_xy; A.pad_z = pad_z;
  A.cta_count = static_cast<int32_t*>(workspace);
  A.tiles = smos_ceil_div(span, kIngestTile);
  A.out = out_points; A.src = out_src; A.count = out_count;
  cudaStream_t st = smos_stream(stream);
  dim3 grid(A.tiles, T);
  SMOS_LAUNCH((ingest_count_kernel), grid, kIngestThreads, 0, st, A);
  SMOS_LAUNCH((ingest_write_kernel), grid, kIngestThreads, 0, st, A);
  return smos_launch_status();
}

}  // extern "C"
