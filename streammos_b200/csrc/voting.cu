// Long-term-memory voting for sm_100a.
//
// Replaces determine_voxel_labels / get_point_labels_from_voxel_labels / Quantize
// (voxel_voting.py:38-91) and the in-box vote count of cluster()
// (voxel_instance_voting.py:169-187). The reference materialises a dense
// (X*Y*Z, num_classes) int64 vote tensor (188 MB), a (P, num_classes) int64 one-hot and runs a
// generic scatter_add + argmax, with a device sync to find num_classes. Here the class counts
// of a voxel are bit-packed into one 64-bit word (21 bits per class, <= 3 classes) so a vote is
// a single 8-byte L2 atomic and the argmax reads 8 bytes per voxel; more classes fall back to
// one 32-bit counter per (voxel, class).
#include "common.cuh"

namespace {

constexpr int kVoteThreads = 256;
constexpr int kPackBits = 21;
constexpr unsigned long long kPackMask = (1ull << kPackBits) - 1ull;
constexpr int64_t kPackMaxPoints = (int64_t(1) << kPackBits) - 1;
constexpr int kMaxClasses = 64;

bool use_packed(int64_t P, int32_t num_classes) { return num_classes <= 3 && P <= kPackMaxPoints; }

// Packed counters of one voxel: three 21-bit fields at bits 1 / 22 / 43 holding the votes of class 2 / 1 / 0; bit 0 is
// never set. A non-empty counter word is therefore even and >= 2, and the only one that is <= 2 is "one vote for
// class 2", whose value (2) equals its own label: for the int64 label grid of the reference API (where the label slot
// itself is the counter) a word <= 2 is always a valid final label, and a word > 2 is a counter still to be converted
// — which makes the per-point conversion pass below idempotent and free of read/write hazards.
__device__ __forceinline__ int vote_shift(int lab) { return 1 + kPackBits * (2 - lab); }

__device__ __forceinline__ unsigned long long packed_argmax(unsigned long long w) {
  const unsigned c2 = static_cast<unsigned>((w >> 1) & kPackMask);
  const unsigned c1 = static_cast<unsigned>((w >> (1 + kPackBits)) & kPackMask);
  const unsigned c0 = static_cast<unsigned>((w >> (1 + 2 * kPackBits)) & kPackMask);
  unsigned long long best = 0ull;  // ties -> lowest class, empty -> 0 (torch.argmax)
  unsigned bv = c0;
  if (c1 > bv) { bv = c1; best = 1ull; }
  if (c2 > bv) { bv = c2; best = 2ull; }
  return best;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kVoteThreads)
vote_i64_kernel(const int64_t* __restrict__ coords, const int64_t* __restrict__ labels, int64_t P,
                int32_t X, int32_t Y, int32_t Z, int32_t C, int packed, void* __restrict__ ws,
                uint32_t* __restrict__ lin_out) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x;
  if (i >= P) return;
  const int64_t x = coords[i * 3], y = coords[i * 3 + 1], z = coords[i * 3 + 2];
  const int64_t lab = labels[i];
  const bool ok = !(x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z || lab < 0 || lab >= C);
  const int64_t lin = ok ? (x * Y + y) * Z + z : -1;
  if (lin_out != nullptr) lin_out[i] = static_cast<uint32_t>(lin);  // 0xffffffff: no vote
  if (!ok) return;
  if (packed)
    atomicAdd(static_cast<unsigned long long*>(ws) + lin, 1ull << vote_shift(static_cast<int>(lab)));
  else
    atomicAdd(static_cast<unsigned int*>(ws) + lin * C + lab, 1u);
}

// Conversion of the voted slots of the int64 label grid, one thread per POINT instead of a pass over the whole grid
// (63 MB for 512 x 512 x 30): a slot holding a counter word (> 2) becomes its label. Several points of one voxel may
// convert it concurrently — they all compute the same label from the same final counts (the vote kernel has
// completed) — and a point that finds a word <= 2 has nothing to do.
__global__ void __launch_bounds__(kVoteThreads)
vote_convert_points_kernel(const uint32_t* __restrict__ lin_in, int64_t P, unsigned long long* __restrict__ slots) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x;
  if (i >= P) return;
  const uint32_t lin = __ldg(lin_in + i);
  if (lin == 0xffffffffu) return;
  // adjacent points mostly share a voxel: one lane per run of equal voxels does the work
  const uint32_t prev = __shfl_up_sync(__activemask(), lin, 1);
  if ((threadIdx.x & 31) != 0 && prev == lin) return;
  const unsigned long long w = __ldcg(slots + lin);
  if (w > 2ull) slots[lin] = packed_argmax(w);
}

__device__ __forceinline__ float quant(float v, float lo, float d) { return __fdiv_rn(__fsub_rn(v, lo), d); }

// float xyz -> voxel index with the reference's arithmetic: fp32 (x - min) / d, then .to(int64)
// truncation (voxel_voting.py:86-88,240). Returns false if outside the grid.
__device__ __forceinline__ bool quant_voxel(const float* p, float mx, float my, float mz, float dx, float dy,
                                            float dz, int32_t X, int32_t Y, int32_t Z, int64_t* lin) {
  const long long x = static_cast<long long>(quant(p[0], mx, dx));
  const long long y = static_cast<long long>(quant(p[1], my, dy));
  const long long z = static_cast<long long>(quant(p[2], mz, dz));
  if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) return false;
  *lin = (static_cast<int64_t>(x) * Y + y) * Z + z;
  return true;
}

__global__ void __launch_bounds__(kVoteThreads)
vote_fused_kernel(const float* __restrict__ pts, int64_t P, int64_t rs, const uint8_t* __restrict__ labels,
                  float mx, float my, float mz, float dx, float dy, float dz, int32_t X, int32_t Y, int32_t Z,
                  int32_t C, int packed, void* __restrict__ ws) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x;
  if (i >= P) return;
  const int lab = labels[i];
  int64_t lin;
  if (lab >= C || !quant_voxel(pts + i * rs, mx, my, mz, dx, dy, dz, X, Y, Z, &lin)) return;
  if (packed)
    atomicAdd(static_cast<unsigned long long*>(ws) + lin, 1ull << vote_shift(lab));
  else
    atomicAdd(static_cast<unsigned int*>(ws) + lin * C + lab, 1u);
}

// argmax with ties -> lowest class, empty voxel -> 0 (torch.argmax on an all-zero row)
template <typename OutT>
__global__ void __launch_bounds__(kVoteThreads)
vote_argmax_kernel(const void* __restrict__ ws, int64_t V, int32_t C, int packed, OutT* __restrict__ out) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x;
  if (i >= V) return;
  int best = 0;
  if (packed) {
    const unsigned long long w = static_cast<const unsigned long long*>(ws)[i];
    if (w != 0ull) best = static_cast<int>(packed_argmax(w));
  } else {
    const unsigned int* r = static_cast<const unsigned int*>(ws) + i * C;
    unsigned bv = r[0];
    for (int c = 1; c < C; ++c) {
      const unsigned v = r[c];
      if (v > bv) { bv = v; best = c; }
    }
  }
  out[i] = static_cast<OutT>(best);
}

// Dense in-place variant (grids with more than 2^32 - 2 voxels, where the per-point pass has no 32-bit voxel index):
// one pass over the int64 label grid turns every non-empty counter word into its label.
__global__ void __launch_bounds__(kVoteThreads)
vote_argmax_inplace_kernel(unsigned long long* __restrict__ slots, int64_t V) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x) * 2;
  if (i >= V) return;
  if (i + 1 < V && (reinterpret_cast<uintptr_t>(slots) & 15) == 0) {
    ulonglong2 w = *reinterpret_cast<const ulonglong2*>(slots + i);
    if ((w.x | w.y) == 0ull) return;
    w.x = packed_argmax(w.x);
    w.y = packed_argmax(w.y);
    *reinterpret_cast<ulonglong2*>(slots + i) = w;  // (packed_argmax of 0 is 0)
  } else {
    for (int64_t k = i; k < V && k < i + 2; ++k) {
      const unsigned long long w = slots[k];
      if (w != 0ull) slots[k] = packed_argmax(w);
    }
  }
}

__global__ void __launch_bounds__(kVoteThreads)
point_labels_i64_kernel(const int64_t* __restrict__ coords, int64_t Pc, const int64_t* __restrict__ vlabels,
                        int32_t X, int32_t Y, int32_t Z, int64_t* __restrict__ out) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x;
  if (i >= Pc) return;
  const int64_t x = coords[i * 3], y = coords[i * 3 + 1], z = coords[i * 3 + 2];
  int64_t r = 0;
  if (x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z) r = __ldg(vlabels + (x * Y + y) * Z + z);
  out[i] = r;
}

__global__ void __launch_bounds__(kVoteThreads)
point_labels_fused_kernel(const float* __restrict__ pts, int64_t Pc, int64_t rs, float mx, float my, float mz,
                          float dx, float dy, float dz, int32_t X, int32_t Y, int32_t Z,
                          const uint8_t* __restrict__ vlabels, int64_t* __restrict__ out) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x;
  if (i >= Pc) return;
  int64_t lin;
  int64_t r = 0;
  if (quant_voxel(pts + i * rs, mx, my, mz, dx, dy, dz, X, Y, Z, &lin)) r = vlabels[lin];
  out[i] = r;
}

__global__ void __launch_bounds__(kVoteThreads)
quantize_kernel(const float* __restrict__ pcds, int64_t P, int64_t rs, float mx, float my, float mz, float dx,
                float dy, float dz, float* __restrict__ out) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x;
  if (i >= P) return;
  const float* p = pcds + i * rs;
  out[i * 3] = quant(p[0], mx, dx);
  out[i * 3 + 1] = quant(p[1], my, dy);
  out[i * 3 + 2] = quant(p[2], mz, dz);
}

// The same expression as torch evaluates it ON A CUDA DEVICE, which is where the reference's scripts run it
// (voxel_voting.py:218-240): a float tensor divided by a Python scalar becomes a multiplication by the float32
// reciprocal of the scalar (ATen BinaryDivTrueKernel.cu, the is_cpu_scalar branch: inv_b = 1.0f / float(b), a * inv_b)
// — one bit away from the IEEE quotient for some inputs, enough to move a point that sits within an ulp of a voxel
// boundary into the neighbouring voxel. rx / ry / rz are those reciprocals.
__global__ void __launch_bounds__(kVoteThreads)
quantize_rcp_kernel(const float* __restrict__ pcds, int64_t P, int64_t rs, float mx, float my, float mz, float rx,
                    float ry, float rz, float* __restrict__ out) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x;
  if (i >= P) return;
  const float* p = pcds + i * rs;
  out[i * 3] = __fmul_rn(__fsub_rn(p[0], mx), rx);
  out[i * 3 + 1] = __fmul_rn(__fsub_rn(p[1], my), ry);
  out[i * 3 + 2] = __fmul_rn(__fsub_rn(p[2], mz), rz);
}

// ---- staging for the reference's int64 voting API (voxel_voting.py:234-241) ----------------------------------
// The script quantises the local map (Quantize -> float32), casts the result and the predictions to int64
// (`.to(torch.int64)`, truncation) and hands both to determine_voxel_labels. Here ONE kernel produces all three
// tensors — q (float32), coords (int64) and labels (int64) — from the float points and uint8 predictions resident in
// the long-term memory ring, and (optionally) performs the ring insert of the new scan on the way: thread i owns
// point i of EVERY slot, so it can read the old "current" point before overwriting it (smos_memory_push fused in).
// A warp parks its 32 x (3 int64 + 3 float) results in shared memory and writes them as whole 16-byte pieces of
// contiguous 768- and 384-byte blocks.
constexpr int kStageThreads = 128;

struct StageArgs {
  float* pts;            // (S, N, rs) ring, slot major
  uint8_t* pred;         // (S, N)
  const float* in_pts;   // (N, rs) new scan or null
  const uint8_t* in_pred;
  int64_t N, rs;
  int32_t S, cur, hist;  // hist < 0: no history slot to fill
  float mx, my, mz, dx, dy, dz;
  float clo[3], chi[3];  // open crop box (transforms.Crop); used when `crop` is set
  int32_t crop;
  float* q;              // (S*N, 3) or null
  int64_t* coords;       // (S*N, 3)
  int64_t* labels;       // (S*N)
  int32_t vec_io;        // rs == 4 and 16-byte aligned point rows
  int32_t vec_out;       // N % 4 == 0 and 16-byte aligned outputs: warp-staged 128-bit stores
};

__global__ void __launch_bounds__(kStageThreads)
vote_stage_kernel(const __grid_constant__ StageArgs A) {
  SMOS_PDL_PROLOGUE();
  __shared__ __align__(16) long long s_c[kStageThreads / 32][32 * 3];
  __shared__ __align__(16) float s_q[kStageThreads / 32][32 * 3];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kStageThreads + threadIdx.x;
  const int64_t i0 = i - lane;  // first point of my warp
  if (i0 >= A.N) return;
  const bool live = i < A.N;
  const bool full = i0 + 32 <= A.N;  // whole warp in range
  const bool push = A.in_pts != nullptr;
  auto load_pt = [&](const float* base, float& x, float& y, float& z, float4& raw) {
    if (A.vec_io) {
      raw = *reinterpret_cast<const float4*>(base + i * 4);
      x = raw.x; y = raw.y; z = raw.z;
    } else {
      const float* p = base + i * A.rs;
      x = p[0]; y = p[1]; z = p[2];
    }
  };
  auto copy_row = [&](float* dst, const float* src, const float4& raw) {
    if (A.vec_io) {
      *reinterpret_cast<float4*>(dst + i * 4) = raw;
    } else {
      for (int64_t k = 0; k < A.rs; ++k) dst[i * A.rs + k] = src[i * A.rs + k];
    }
  };
  float ox = 0.f, oy = 0.f, oz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  uint8_t op = 0, np = 0;
  if (push && live) {  // ring insert: old current -> history slot, new scan -> current slot
    float4 oraw = make_float4(0.f, 0.f, 0.f, 0.f), nraw = oraw;
    float* cur_pts = A.pts + static_cast<int64_t>(A.cur) * A.N * A.rs;
    uint8_t* cur_pred = A.pred + static_cast<int64_t>(A.cur) * A.N;
    load_pt(cur_pts, ox, oy, oz, oraw);
    op = cur_pred[i];
    load_pt(A.in_pts, nx, ny, nz, nraw);
    np = A.in_pred[i];
    if (A.hist >= 0) {
      copy_row(A.pts + static_cast<int64_t>(A.hist) * A.N * A.rs, cur_pts, oraw);
      A.pred[static_cast<int64_t>(A.hist) * A.N + i] = op;
    }
    copy_row(cur_pts, A.in_pts, nraw);
    cur_pred[i] = np;
  }
  // The slots are independent, but the loop body (shared-memory staging, warp syncs) keeps the compiler from hoisting
  // the next slot's loads: each slot's point and label are fetched one iteration AHEAD, so their latency hides behind
  // the previous slot's quantisation and stores (ncu: 37 % of the kernel's stall samples sat on the first FADD
  // behind the point load).
  auto from_ring = [&](int32_t s) { return live && s < A.S && !(push && (s == A.cur || s == A.hist)); };
  float px = 0.f, py = 0.f, pz = 0.f;
  uint8_t plab = 0;
  if (from_ring(0)) {
    float4 raw;
    load_pt(A.pts, px, py, pz, raw);
    plab = A.pred[i];
  }
  for (int32_t s = 0; s < A.S; ++s) {
    float x = px, y = py, z = pz;
    uint8_t lab = plab;
    if (from_ring(s + 1)) {  // prefetch of the next slot
      float4 raw;
      load_pt(A.pts + static_cast<int64_t>(s + 1) * A.N * A.rs, px, py, pz, raw);
      plab = A.pred[static_cast<int64_t>(s + 1) * A.N + i];
    }
    if (live) {
      if (push && s == A.cur) { x = nx; y = ny; z = nz; lab = np; }
      else if (push && s == A.hist) { x = ox; y = oy; z = oz; lab = op; }
    } else {
      x = y = z = 0.f;
      lab = 0;
    }
    const float qx = quant(x, A.mx, A.dx), qy = quant(y, A.my, A.dy), qz = quant(z, A.mz, A.dz);
    // .to(torch.int64): truncation toward zero (voxel_voting.py:240)
    long long cx = static_cast<long long>(qx), cy = static_cast<long long>(qy), cz = static_cast<long long>(qz);
    // the script crops before it quantises: a point outside the open box takes no part in the vote
    if (A.crop && !(x > A.clo[0] && x < A.chi[0] && y > A.clo[1] && y < A.chi[1] && z > A.clo[2] && z < A.chi[2]))
      cx = cy = cz = -1;
    const int64_t g = static_cast<int64_t>(s) * A.N + i;
    if (live) A.labels[g] = lab;
    if (A.vec_out && full) {
      __syncwarp();
      s_c[wid][lane * 3] = cx; s_c[wid][lane * 3 + 1] = cy; s_c[wid][lane * 3 + 2] = cz;
      if (A.q) { s_q[wid][lane * 3] = qx; s_q[wid][lane * 3 + 1] = qy; s_q[wid][lane * 3 + 2] = qz; }
      __syncwarp();
      const int64_t g0 = static_cast<int64_t>(s) * A.N + i0;
      uint4* dc = reinterpret_cast<uint4*>(A.coords + g0 * 3);          // 768 contiguous bytes = 48 pieces
      const uint4* sc = reinterpret_cast<const uint4*>(&s_c[wid][0]);
      dc[lane] = sc[lane];
      if (lane < 16) dc[32 + lane] = sc[32 + lane];
      if (A.q && lane < 24)                                                // 384 contiguous bytes = 24 pieces
        reinterpret_cast<uint4*>(A.q + g0 * 3)[lane] = reinterpret_cast<const uint4*>(&s_q[wid][0])[lane];
    } else if (live) {
      A.coords[g * 3] = cx; A.coords[g * 3 + 1] = cy; A.coords[g * 3 + 2] = cz;
      if (A.q) { A.q[g * 3] = qx; A.q[g * 3 + 1] = qy; A.q[g * 3 + 2] = qz; }
    }
  }
}

// Instance vote. Testing every point against every box is O(P*K) compares (P ~ 1.08 M); almost all
// points are near no box. Each CTA therefore first bins the boxes of its chunk into a coarse 16 x 16
// grid over their common xy extent (a bit mask of boxes per coarse cell, in shared memory); a point
// looks up its coarse cell — the mapping is monotonic and clamped, so a point inside a box always
// lands in a cell that lists the box — and runs the exact inclusive test only for the listed boxes.
constexpr int kBoxChunk = 256;
constexpr int kIvGrid = 16;
constexpr int kIvWords = kBoxChunk / 32;

__device__ __forceinline__ int iv_cell(float v, float lo, float inv) {
  const float f = (v - lo) * inv;
  return f > 0.f ? (f < static_cast<float>(kIvGrid - 1) ? static_cast<int>(f) : kIvGrid - 1) : 0;
}

__global__ void __launch_bounds__(kVoteThreads)
instance_vote_kernel(const float* __restrict__ pts, int64_t P, int64_t rs, const int64_t* __restrict__ pred,
                     const float* __restrict__ lo, const float* __restrict__ hi, int32_t K,
                     const int32_t* __restrict__ k_dev, unsigned long long* __restrict__ sums,
                     unsigned long long* __restrict__ acc_ws, unsigned long long* __restrict__ out_ws) {
  SMOS_PDL_PROLOGUE();
  // acc_ws == null: the CTAs add straight into `sums` (zero-filled by the caller).
  // acc_ws != null: they add into the persistent accumulator acc_ws[0 .. 2K) (zero on entry); the last CTA to finish
  // (ticket in acc_ws[2K]) moves the totals to out_ws and leaves accumulator and ticket zero for the next call — no
  // fill kernel in front of the vote.
  if (acc_ws != nullptr) sums = acc_ws;
  __shared__ float4 s_box[kBoxChunk * 2];  // (lo.x, lo.y, lo.z, hi.x) (hi.y, hi.z, -, -)
  __shared__ unsigned int s_cnt[kBoxChunk * 2];
  __shared__ unsigned int s_mask[kIvGrid * kIvGrid][kIvWords];
  __shared__ int s_ext[4];  // ordered-int keys of min x, min y, max x, max y
  // the number of boxes may live on the device (boxes produced by smos_cluster_boxes): no host read in between
  const int32_t Kt = k_dev ? min(__ldg(k_dev), K) : K;
  for (int32_t k0 = blockIdx.y * kBoxChunk; k0 < Kt; k0 += gridDim.y * kBoxChunk) {
    const int32_t kn = min(kBoxChunk, Kt - k0);
    if (threadIdx.x == 0) { s_ext[0] = s_ext[1] = 0x7fffffff; s_ext[2] = s_ext[3] = static_cast<int>(0x80000000u); }
    for (int i = threadIdx.x; i < kIvGrid * kIvGrid * kIvWords; i += kVoteThreads) (&s_mask[0][0])[i] = 0u;
    for (int i = threadIdx.x; i < kn * 2; i += kVoteThreads) s_cnt[i] = 0u;
    __syncthreads();
    auto key = [](float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; };  // monotonic
    auto unkey = [](int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); };
    for (int i = threadIdx.x; i < kn; i += kVoteThreads) {
      const float* l = lo + static_cast<int64_t>(k0 + i) * 3;
      const float* h = hi + static_cast<int64_t>(k0 + i) * 3;
      s_box[2 * i] = make_float4(l[0], l[1], l[2], h[0]);
      s_box[2 * i + 1] = make_float4(h[1], h[2], 0.f, 0.f);
      atomicMin(&s_ext[0], key(l[0])); atomicMin(&s_ext[1], key(l[1]));
      atomicMax(&s_ext[2], key(h[0])); atomicMax(&s_ext[3], key(h[1]));
    }
    __syncthreads();
    const float gx0 = unkey(s_ext[0]), gy0 = unkey(s_ext[1]);
    const float ex = unkey(s_ext[2]) - gx0, ey = unkey(s_ext[3]) - gy0;
    const float invx = ex > 0.f ? static_cast<float>(kIvGrid) / ex : 0.f;
    const float invy = ey > 0.f ? static_cast<float>(kIvGrid) / ey : 0.f;
    for (int i = threadIdx.x; i < kn; i += kVoteThreads) {
      const float4 a = s_box[2 * i], bb = s_box[2 * i + 1];
      const int x0 = iv_cell(a.x, gx0, invx), x1 = iv_cell(a.w, gx0, invx);
      const int y0 = iv_cell(a.y, gy0, invy), y1 = iv_cell(bb.x, gy0, invy);
      for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) atomicOr(&s_mask[y * kIvGrid + x][i >> 5], 1u << (i & 31));
    }
    __syncthreads();
    const int nwords = (kn + 31) >> 5;
    const bool vec = (rs == 4) && ((reinterpret_cast<uintptr_t>(pts) & 15) == 0);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x; i < P;
         i += static_cast<int64_t>(gridDim.x) * kVoteThreads) {
      // prediction and point are fetched together (the point of a background point is wasted bandwidth, but a
      // dependent second round trip per point was 18 % of the kernel's stall samples)
      const int64_t pr = __ldg(pred + i);
      float x, y, z;
      if (vec) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(pts) + i);
        x = q.x; y = q.y; z = q.z;
      } else {
        const float* q = pts + i * rs;
        x = __ldg(q); y = __ldg(q + 1); z = __ldg(q + 2);
      }
      if (pr != 1 && pr != 2) continue;
      const unsigned int* m = s_mask[iv_cell(y, gy0, invy) * kIvGrid + iv_cell(x, gx0, invx)];
      for (int w = 0; w < nwords; ++w) {
        unsigned int bits = m[w];
        while (bits) {
          const int k = (w << 5) + __ffs(bits) - 1;
          bits &= bits - 1;
          const float4 a = s_box[2 * k], bb = s_box[2 * k + 1];
          // inclusive AABB test == in_hull of the 8 corners (voxel_instance_voting.py:62-76,177)
          if (x >= a.x && x <= a.w && y >= a.y && y <= bb.x && z >= a.z && z <= bb.y)
            atomicAdd(&s_cnt[k * 2 + static_cast<int>(pr) - 1], 1u);
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kn * 2; i += kVoteThreads) {
      const unsigned int c = s_cnt[i];
      // sum(pred[pred==2]) counts 2 per dynamic point (voxel_instance_voting.py:182-184)
      if (c) atomicAdd(&sums[k0 * 2 + i], static_cast<unsigned long long>(c) * ((i & 1) ? 2ull : 1ull));
    }
    __syncthreads();  // the shared tables are rebuilt for the next chunk of boxes
  }
  if (acc_ws != nullptr) {
    __shared__ bool s_last;
    __threadfence();  // my atomics are performed before my ticket
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int* ticket = reinterpret_cast<unsigned int*>(acc_ws + 2 * static_cast<int64_t>(K));
      s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1u;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      for (int i = threadIdx.x; i < 2 * K; i += kVoteThreads) {
        out_ws[i] = __ldcg(acc_ws + i);
        acc_ws[i] = 0ull;
      }
      if (threadIdx.x == 0) *reinterpret_cast<unsigned int*>(acc_ws + 2 * static_cast<int64_t>(K)) = 0u;
    }
  }
}

// ---- streaming long-term memory -------------------------------------------------
// One call per scan over a history that stays resident in HBM: pose alignment of the history scans
// (datasets/utils.py:116-126 Trans: float64 4x4 x (x,y,z,1), stored as float32 — an FMA chain in k order
// reproduces numpy's dgemm bit for bit), crop to the open box (utils/transforms.py:151-161), quantise
// (voxel_voting.py:77-91), vote. Replaces the per-frame reload + numpy transform + crop + .cuda() round trip
// of voxel_voting.py:176-230.
constexpr int kMaxStreamScans = 16;

struct StreamScans {
  const float* pts[kMaxStreamScans];
  const uint8_t* lab[kMaxStreamScans];
  int64_t begin[kMaxStreamScans + 1];  // prefix of point counts
  double m[kMaxStreamScans][12];       // rows 0..2 of pose_diff = inv(pose_cur) . pose_scan
  int32_t transform[kMaxStreamScans];
  int32_t n;
};

struct CropBox { float lo[3], hi[3]; };  // thresholds already include eps: keep iff lo < p < hi

__device__ __forceinline__ bool stream_point(const StreamScans& S, int j, int64_t i, int64_t rs, const CropBox& box,
                                             float* q) {
  const float* p = S.pts[j] + i * rs;
  float x = p[0], y = p[1], z = p[2];
  if (S.transform[j]) {
    const double dx = x, dy = y, dz = z;
    const double* m = S.m[j];
    x = static_cast<float>(fma(m[3], 1.0, fma(m[2], dz, fma(m[1], dy, m[0] * dx))));
    y = static_cast<float>(fma(m[7], 1.0, fma(m[6], dz, fma(m[5], dy, m[4] * dx))));
    z = static_cast<float>(fma(m[11], 1.0, fma(m[10], dz, fma(m[9], dy, m[8] * dx))));
  }
  q[0] = x; q[1] = y; q[2] = z;
  return x > box.lo[0] && x < box.hi[0] && y > box.lo[1] && y < box.hi[1] && z > box.lo[2] && z < box.hi[2];
}

__global__ void __launch_bounds__(kVoteThreads)
vote_stream_kernel(const __grid_constant__ StreamScans S, int64_t rs, const __grid_constant__ CropBox box, float mx,
                   float my, float mz, float dx, float dy, float dz, int32_t X, int32_t Y, int32_t Z, int32_t C,
                   int packed, void* __restrict__ ws) {
  SMOS_PDL_PROLOGUE();
  const int64_t g = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x;
  if (g >= S.begin[S.n]) return;
  int j = 0;
  while (j + 1 < S.n && g >= S.begin[j + 1]) ++j;
  const int64_t i = g - S.begin[j];
  float q[3];
  if (!stream_point(S, j, i, rs, box, q)) return;
  const int lab = S.lab[j][i];
  int64_t lin;
  if (lab >= C || !quant_voxel(q, mx, my, mz, dx, dy, dz, X, Y, Z, &lin)) return;
  if (packed)
    atomicAdd(static_cast<unsigned long long*>(ws) + lin, 1ull << vote_shift(lab));
  else
    atomicAdd(static_cast<unsigned int*>(ws) + lin * C + lab, 1u);
}

// current scan: points inside the crop take their voxel's label, the others keep their own prediction
// (voxel_voting.py:243-244: current_pred_result_orin[mask] = pred_result_new)
__global__ void __launch_bounds__(kVoteThreads)
stream_point_labels_kernel(const __grid_constant__ StreamScans S, int cur, int64_t rs, const __grid_constant__ CropBox box,
                           float mx, float my, float mz, float dx, float dy, float dz, int32_t X, int32_t Y, int32_t Z,
                           const uint8_t* __restrict__ vlabels, int64_t* __restrict__ out) {
  SMOS_PDL_PROLOGUE();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x;
  if (i >= S.begin[cur + 1] - S.begin[cur]) return;
  float q[3];
  int64_t r = S.lab[cur][i];
  if (stream_point(S, cur, i, rs, box, q)) {
    int64_t lin;
    r = quant_voxel(q, mx, my, mz, dx, dy, dz, X, Y, Z, &lin) ? vlabels[lin] : 0;
  }
  out[i] = r;
}

// ---- long-term memory ring: insert the new scan -----------------------------------
// One kernel per scan instead of four memcpy nodes: the scan that was current until now moves from the "current"
// slot into its history slot and the new scan (points + predicted labels) takes the current slot.
__global__ void __launch_bounds__(kVoteThreads)
memory_push_kernel(const float* __restrict__ pts_in, const uint8_t* __restrict__ pred_in, int64_t nfloat, int64_t n,
                   float* __restrict__ cur_pts, uint8_t* __restrict__ cur_pred, float* __restrict__ hist_pts,
                   uint8_t* __restrict__ hist_pred, int vec) {
  SMOS_PDL_PROLOGUE();
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kVoteThreads + threadIdx.x;
  const int64_t nthreads = static_cast<int64_t>(gridDim.x) * kVoteThreads;
  if (vec) {  // every pointer 16-byte aligned, nfloat % 4 == 0, n % 16 == 0
    const int64_t n4 = nfloat >> 2, n16 = n >> 4;
    for (int64_t i = tid; i < n4; i += nthreads) {
      if (hist_pts) reinterpret_cast<float4*>(hist_pts)[i] = reinterpret_cast<const float4*>(cur_pts)[i];
      reinterpret_cast<float4*>(cur_pts)[i] = __ldg(reinterpret_cast<const float4*>(pts_in) + i);
    }
    for (int64_t i = tid; i < n16; i += nthreads) {
      if (hist_pred) reinterpret_cast<uint4*>(hist_pred)[i] = reinterpret_cast<const uint4*>(cur_pred)[i];
      reinterpret_cast<uint4*>(cur_pred)[i] = __ldg(reinterpret_cast<const uint4*>(pred_in) + i);
    }
  } else {
    for (int64_t i = tid; i < nfloat; i += nthreads) {
      if (hist_pts) hist_pts[i] = cur_pts[i];
      cur_pts[i] = pts_in[i];
    }
    for (int64_t i = tid; i < n; i += nthreads) {
      if (hist_pred) hist_pred[i] = cur_pred[i];
      cur_pred[i] = pred_in[i];
    }
  }
}

int64_t ws_bytes(int64_t P, int64_t V, int32_t C) {
  // packed: V counter words (the fused / streaming calls) or P 32-bit voxel indices (the int64 API, whose label grid
  // holds the counters itself), whichever is larger
  if (use_packed(P, C)) return V * 8 > P * 4 ? V * 8 : P * 4;
  return V * static_cast<int64_t>(C) * 4;
}

}  // namespace

extern "C" {

int smos_quantize(const float* pcds, int64_t P, int64_t row_stride, float min_x, float min_y, float min_z,
                  float dx, float dy, float dz, float* out, void* stream) {
  if (P < 0 || row_stride < 3) return SMOS_EINVAL;
  if (P == 0) return SMOS_OK;
  if (!pcds || !out) return SMOS_EINVAL;
  SMOS_LAUNCH((quantize_kernel), smos_ceil_div(P, kVoteThreads), kVoteThreads, 0, smos_stream(stream), 
      pcds, P, row_stride, min_x, min_y, min_z, dx, dy, dz, out);
  return smos_launch_status();
}

int smos_quantize_rcp(const float* pcds, int64_t P, int64_t row_stride, float min_x, float min_y, float min_z, float dx,
                      float dy, float dz, float* out, void* stream) {
  if (P < 0 || row_stride < 3 || dx == 0.f || dy == 0.f || dz == 0.f) return SMOS_EINVAL;
  if (P == 0) return SMOS_OK;
  if (!pcds || !out) return SMOS_EINVAL;
  // float32 reciprocals, IEEE division on the host: what `opmath_t(1.0) / b` gives in ATen
  const volatile float rx = 1.0f / dx, ry = 1.0f / dy, rz = 1.0f / dz;
  SMOS_LAUNCH((quantize_rcp_kernel), smos_ceil_div(P, kVoteThreads), kVoteThreads, 0, smos_stream(stream),
      pcds, P, row_stride, min_x, min_y, min_z, static_cast<float>(rx), static_cast<float>(ry), static_cast<float>(rz), out);
  return smos_launch_status();
}

int64_t smos_vote_workspace_bytes(int64_t P, int32_t X, int32_t Y, int32_t Z, int32_t num_classes) {
  if (P < 0 || X <= 0 || Y <= 0 || Z <= 0 || num_classes <= 0 || num_classes > kMaxClasses) return SMOS_EINVAL;
  return ws_bytes(P, static_cast<int64_t>(X) * Y * Z, num_classes);
}

int smos_vote_voxel_labels(const int64_t* voxel_coords, const int64_t* semantic_labels, int64_t P, int32_t X,
                           int32_t Y, int32_t Z, int32_t num_classes, void* workspace, int64_t* voxel_labels,
                           void* stream) {
  if (P < 0 || X <= 0 || Y <= 0 || Z <= 0 || num_classes <= 0 || num_classes > kMaxClasses) return SMOS_EINVAL;
  if (!workspace || !voxel_labels || (P > 0 && (!voxel_coords || !semantic_labels))) return SMOS_EINVAL;
  const int64_t V = static_cast<int64_t>(X) * Y * Z;
  const int packed = use_packed(P, num_classes) ? 1 : 0;
  cudaStream_t st = smos_stream(stream);
  if (packed) {  // the label slots double as the packed counters (workspace untouched)
    cudaError_t e = smos_zero_async(voxel_labels, static_cast<size_t>(V) * 8, st);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (P > 0 && V < 0xffffffffll) {
      uint32_t* lin = static_cast<uint32_t*>(workspace);
      SMOS_LAUNCH((vote_i64_kernel), smos_ceil_div(P, kVoteThreads), kVoteThreads, 0, st, voxel_coords, semantic_labels, P, X, Y,
                                                                               Z, num_classes, 1, voxel_labels, lin);
      SMOS_LAUNCH((vote_convert_points_kernel), smos_ceil_div(P, kVoteThreads), kVoteThreads, 0, st, 
          lin, P, reinterpret_cast<unsigned long long*>(voxel_labels));
    } else if (P > 0) {
      SMOS_LAUNCH((vote_i64_kernel), smos_ceil_div(P, kVoteThreads), kVoteThreads, 0, st, voxel_coords, semantic_labels, P, X, Y,
                                                                               Z, num_classes, 1, voxel_labels, nullptr);
      SMOS_LAUNCH((vote_argmax_inplace_kernel), smos_ceil_div((V + 1) / 2, kVoteThreads), kVoteThreads, 0, st, 
          reinterpret_cast<unsigned long long*>(voxel_labels), V);
    }
    return smos_launch_status();
  }
  cudaError_t e = smos_zero_async(workspace, static_cast<size_t>(ws_bytes(P, V, num_classes)), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (P > 0)
    SMOS_LAUNCH((vote_i64_kernel), smos_ceil_div(P, kVoteThreads), kVoteThreads, 0, st, voxel_coords, semantic_labels, P, X, Y,
                                                                             Z, num_classes, 0, workspace, nullptr);
  SMOS_LAUNCH((vote_argmax_kernel<int64_t>), smos_ceil_div(V, kVoteThreads), kVoteThreads, 0, st, workspace, V, num_classes,
                                                                                       0, voxel_labels);
  return smos_launch_status();
}

int smos_vote_point_labels(const int64_t* new_voxel_coords, int64_t Pc, const int64_t* voxel_labels, int32_t X,
                           int32_t Y, int32_t Z, int64_t* point_labels, void* stream) {
  if (Pc < 0 || X <= 0 || Y <= 0 || Z <= 0) return SMOS_EINVAL;
  if (Pc == 0) return SMOS_OK;
  if (!new_voxel_coords || !voxel_labels || !point_labels) return SMOS_EINVAL;
  SMOS_LAUNCH((point_labels_i64_kernel), smos_ceil_div(Pc, kVoteThreads), kVoteThreads, 0, smos_stream(stream), 
      new_voxel_coords, Pc, voxel_labels, X, Y, Z, point_labels);
  return smos_launch_status();
}

int smos_vote_fused(const float* points, int64_t P, int64_t row_stride, const uint8_t* labels, int64_t Pc,
                    float min_x, float min_y, float min_z, float dx, float dy, float dz, int32_t X, int32_t Y,
                    int32_t Z, int32_t num_classes, void* workspace, uint8_t* voxel_labels_u8,
                    int64_t* point_labels, void* stream) {
  if (P < 0 || Pc < 0 || Pc > P || row_stride < 3 || X <= 0 || Y <= 0 || Z <= 0 || num_classes <= 0 ||
      num_classes > kMaxClasses)
    return SMOS_EINVAL;
  if (!workspace || !voxel_labels_u8 || (P > 0 && (!points || !labels)) || (Pc > 0 && !point_labels))
    return SMOS_EINVAL;
  const int64_t V = static_cast<int64_t>(X) * Y * Z;
  const int packed = use_packed(P, num_classes) ? 1 : 0;
  cudaStream_t st = smos_stream(stream);
  cudaError_t e = smos_zero_async(workspace, static_cast<size_t>(ws_bytes(P, V, num_classes)), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (P > 0)
    SMOS_LAUNCH((vote_fused_kernel), smos_ceil_div(P, kVoteThreads), kVoteThreads, 0, st, 
        points, P, row_stride, labels, min_x, min_y, min_z, dx, dy, dz, X, Y, Z, num_classes, packed, workspace);
  SMOS_LAUNCH((vote_argmax_kernel<uint8_t>), smos_ceil_div(V, kVoteThreads), kVoteThreads, 0, st, workspace, V, num_classes,
                                                                                       packed, voxel_labels_u8);
  if (Pc > 0)
    SMOS_LAUNCH((point_labels_fused_kernel), smos_ceil_div(Pc, kVoteThreads), kVoteThreads, 0, st, 
        points + (P - Pc) * row_stride, Pc, row_stride, min_x, min_y, min_z, dx, dy, dz, X, Y, Z, voxel_labels_u8,
        point_labels);
  return smos_launch_status();
}

int smos_vote_stream(const smos_vote_stream_scan* scans_host, int32_t n_scans, int32_t current, int64_t row_stride,
                     const float* crop_lo_host, const float* crop_hi_host, float min_x, float min_y, float min_z,
                     float dx, float dy, float dz, int32_t X, int32_t Y, int32_t Z, int32_t num_classes,
                     void* workspace, uint8_t* voxel_labels_u8, int64_t* point_labels, void* stream) {
  if (!scans_host || n_scans <= 0 || n_scans > kMaxStreamScans || current < 0 || current >= n_scans || row_stride < 3 ||
      X <= 0 || Y <= 0 || Z <= 0 || num_classes <= 0 || num_classes > kMaxClasses || !crop_lo_host || !crop_hi_host)
    return SMOS_EINVAL;
  if (!workspace || !voxel_labels_u8 || !point_labels) return SMOS_EINVAL;
  StreamScans S;
  S.n = n_scans;
  int64_t total = 0;
  for (int32_t j = 0; j < n_scans; ++j) {
    const smos_vote_stream_scan& d = scans_host[j];
    if (d.n < 0 || (d.n > 0 && (!d.points || !d.labels))) return SMOS_EINVAL;
    S.pts[j] = d.points; S.lab[j] = d.labels; S.begin[j] = total; S.transform[j] = d.transform;
    for (int k = 0; k < 12; ++k) S.m[j][k] = d.pose_diff[k];
    total += d.n;
  }
  S.begin[n_scans] = total;
  CropBox box;
  for (int k = 0; k < 3; ++k) { box.lo[k] = crop_lo_host[k]; box.hi[k] = crop_hi_host[k]; }
  const int64_t V = static_cast<int64_t>(X) * Y * Z;
  const int packed = use_packed(total, num_classes) ? 1 : 0;
  cudaStream_t st = smos_stream(stream);
  cudaError_t e = smos_zero_async(workspace, static_cast<size_t>(ws_bytes(total, V, num_classes)), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (total > 0)
    SMOS_LAUNCH((vote_stream_kernel), smos_ceil_div(total, kVoteThreads), kVoteThreads, 0, st, 
        S, row_stride, box, min_x, min_y, min_z, dx, dy, dz, X, Y, Z, num_classes, packed, workspace);
  SMOS_LAUNCH((vote_argmax_kernel<uint8_t>), smos_ceil_div(V, kVoteThreads), kVoteThreads, 0, st, workspace, V, num_classes,
                                                                                       packed, voxel_labels_u8);
  const int64_t nc = scans_host[current].n;
  if (nc > 0)
    SMOS_LAUNCH((stream_point_labels_kernel), smos_ceil_div(nc, kVoteThreads), kVoteThreads, 0, st, 
        S, current, row_stride, box, min_x, min_y, min_z, dx, dy, dz, X, Y, Z, voxel_labels_u8, point_labels);
  return smos_launch_status();
}

int smos_vote_stage(float* ring_points, uint8_t* ring_pred, int32_t n_slots, int64_t n, int64_t row_floats,
                    const float* new_points, const uint8_t* new_pred, int32_t cur_slot, int32_t hist_slot,
                    float min_x, float min_y, float min_z, float dx, float dy, float dz,
                    const float* crop_lo_host, const float* crop_hi_host,
                    float* q_out, int64_t* coords_out, int64_t* labels_out, void* stream) {
  if (n_slots <= 0 || n < 0 || row_floats < 3) return SMOS_EINVAL;
  if (n == 0) return SMOS_OK;
  if (!ring_points || !ring_pred || !coords_out || !labels_out) return SMOS_EINVAL;
  if ((new_points == nullptr) != (new_pred == nullptr)) return SMOS_EINVAL;
  if (new_points && (cur_slot < 0 || cur_slot >= n_slots || hist_slot >= n_slots || hist_slot == cur_slot)) return SMOS_EINVAL;
  auto a16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  StageArgs A;
  A.pts = ring_points; A.pred = ring_pred; A.in_pts = new_points; A.in_pred = new_pred;
  A.N = n; A.rs = row_floats; A.S = n_slots; A.cur = cur_slot; A.hist = new_points ? hist_slot : -1;
  A.mx = min_x; A.my = min_y; A.mz = min_z; A.dx = dx; A.dy = dy; A.dz = dz;
  A.q = q_out; A.coords = coords_out; A.labels = labels_out;
  if ((crop_lo_host == nullptr) != (crop_hi_host == nullptr)) return SMOS_EINVAL;
  A.crop = crop_lo_host != nullptr ? 1 : 0;
  for (int k = 0; k < 3; ++k) { A.clo[k] = A.crop ? crop_lo_host[k] : 0.f; A.chi[k] = A.crop ? crop_hi_host[k] : 0.f; }
  A.vec_io = (row_floats == 4 && a16(ring_points) && a16(new_points)) ? 1 : 0;
  A.vec_out = ((n & 3) == 0 && a16(q_out) && a16(coords_out)) ? 1 : 0;
  SMOS_LAUNCH((vote_stage_kernel), smos_ceil_div(n, kStageThreads), kStageThreads, 0, smos_stream(stream), A);
  return smos_launch_status();
}

int smos_memory_push(const float* points, const uint8_t* pred, int64_t n, int64_t row_floats, float* cur_points,
                     uint8_t* cur_pred, float* hist_points, uint8_t* hist_pred, void* stream) {
  if (n < 0 || row_floats < 3) return SMOS_EINVAL;
  if (n == 0) return SMOS_OK;
  if (!points || !pred || !cur_points || !cur_pred || ((hist_points == nullptr) != (hist_pred == nullptr))) return SMOS_EINVAL;
  const int64_t nfloat = n * row_floats;
  auto a16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int vec = (a16(points) && a16(pred) && a16(cur_points) && a16(cur_pred) && a16(hist_points) && a16(hist_pred) &&
                   (nfloat & 3) == 0 && (n & 15) == 0) ? 1 : 0;
  int64_t blocks = smos_ceil_div(vec ? (nfloat >> 2) : nfloat, kVoteThreads);
  if (blocks > SMOS_SM_COUNT * 8) blocks = SMOS_SM_COUNT * 8;
  SMOS_LAUNCH((memory_push_kernel), static_cast<unsigned>(blocks), kVoteThreads, 0, smos_stream(stream), 
      points, pred, nfloat, n, cur_points, cur_pred, hist_points, hist_pred, vec);
  return smos_launch_status();
}

int smos_instance_vote(const float* points, int64_t P, int64_t row_stride, const int64_t* pred,
                       const float* box_lo, const float* box_hi, int32_t K, int64_t* sums, void* stream) {
  if (P < 0 || K < 0 || row_stride < 3) return SMOS_EINVAL;
  if (K == 0 || P == 0) return SMOS_OK;
  if (!points || !pred || !box_lo || !box_hi || !sums) return SMOS_EINVAL;
  int gx = smos_ceil_div(P, kVoteThreads * 4);
  if (gx > 8 * SMOS_SM_COUNT) gx = 8 * SMOS_SM_COUNT;  // persistent-style grid-stride loop
  dim3 grid(gx, smos_ceil_div(K, kBoxChunk));
  SMOS_LAUNCH((instance_vote_kernel), grid, kVoteThreads, 0, smos_stream(stream), 
      points, P, row_stride, pred, box_lo, box_hi, K, nullptr, reinterpret_cast<unsigned long long*>(sums), nullptr,
      nullptr);
  return smos_launch_status();
}

int64_t smos_instance_vote_workspace_bytes(int32_t K) {
  if (K < 0) return SMOS_EINVAL;
  return (2 * static_cast<int64_t>(K) + 2) * 8;
}

int smos_instance_vote_ws(const float* points, int64_t P, int64_t row_stride, const int64_t* pred,
                          const float* box_lo, const float* box_hi, int32_t K_cap, const int32_t* K_dev,
                          void* workspace, int64_t* sums, void* stream) {
  if (P < 0 || K_cap < 0 || row_stride < 3) return SMOS_EINVAL;
  if (K_cap == 0) return SMOS_OK;
  if (!sums || !workspace || (reinterpret_cast<uintptr_t>(workspace) & 7) != 0) return SMOS_EINVAL;
  if (P > 0 && (!points || !pred || !box_lo || !box_hi)) return SMOS_EINVAL;
  int gx = smos_ceil_div(P > 0 ? P : 1, kVoteThreads * 4);
  if (gx > 8 * SMOS_SM_COUNT) gx = 8 * SMOS_SM_COUNT;
  // one grid row when the box count lives on the device (every CTA walks the chunks that exist)
  dim3 grid(gx, K_dev ? 1 : smos_ceil_div(K_cap, kBoxChunk));
  SMOS_LAUNCH((instance_vote_kernel), grid, kVoteThreads, 0, smos_stream(stream), 
      points, P, row_stride, pred, box_lo, box_hi, K_cap, K_dev, nullptr,
      static_cast<unsigned long long*>(workspace), reinterpret_cast<unsigned long long*>(sums));
  return smos_launch_status();
}

int smos_instance_vote_counted(const float* points, int64_t P, int64_t row_stride, const int64_t* pred,
                               const float* box_lo, const float* box_hi, int32_t K_cap, const int32_t* K_dev,
                               int64_t* sums, void* stream) {
  if (P < 0 || K_cap < 0 || row_stride < 3 || !K_dev) return SMOS_EINVAL;
  if (K_cap == 0 || P == 0) return SMOS_OK;
  if (!points || !pred || !box_lo || !box_hi || !sums) return SMOS_EINVAL;
  int gx = smos_ceil_div(P, kVoteThreads * 4);
  if (gx > 8 * SMOS_SM_COUNT) gx = 8 * SMOS_SM_COUNT;
  // one grid row: every CTA walks the chunks of boxes that exist (usually one), none when *K_dev == 0
  SMOS_LAUNCH((instance_vote_kernel), dim3(gx, 1), kVoteThreads, 0, smos_stream(stream), 
      points, P, row_stride, pred, box_lo, box_hi, K_cap, K_dev, reinterpret_cast<unsigned long long*>(sums), nullptr,
      nullptr);
  return smos_launch_status();
}

}  // extern "C"
