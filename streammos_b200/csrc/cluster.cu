// Instance clustering of the long-term voting stage: the part of cluster()
// (voxel_instance_voting.py:144-175) that precedes the vote block — foreground selection, DBSCAN(eps 0.3,
// min_samples 5), the >30-point cut, and one axis-aligned box per cluster with its floor lifted by 0.2 — and the
// write-back that follows it (:189-191).
//
// The reference runs scikit-learn's DBSCAN on the host: a KD-tree radius query per point and a sequential
// depth-first expansion whose border-point labels depend on the visiting order. Here the same labels come from an
// order-free formulation that maps to the GPU:
//   * neighbour test: ((dx*dx) + dy*dy) + dz*dz <= eps*eps in float64 without contraction (what the KD-tree's
//     reduced distance evaluates), brute force over M^2 pairs in shared-memory tiles with a float32 reject first
//     (M = moving points of one scan, a few thousand);
//   * core points: >= min_samples neighbours, the point itself included;
//   * clusters: connected components of core points under the neighbour relation — lock-free union-find whose
//     root is the smallest index of the component; sklearn numbers clusters in the order of their first core
//     point, i.e. by that root, so cluster id = rank of the root among roots;
//   * border points: sklearn expands cluster 0 completely, then cluster 1, ...; a border point therefore takes the
//     smallest cluster id among its core neighbours; points without a core neighbour are noise (-1).
// Ordered compactions (foreground indices ascending, root ranks, kept clusters in id order) use per-CTA counts
// and a sum over the preceding CTAs, so every result is deterministic.
//
// Nothing here synchronises: M and the cluster counts stay on the device; every kernel is launched for the upper
// bound n and reads the actual count.
#include "common.cuh"

namespace {

constexpr int kClThreads = 256;
constexpr int kClWarps = kClThreads / 32;

struct ClusterWs {
  float4* xyz;        // [n] foreground points (x, y, z, -)
  int32_t* nnb;       // [n] neighbour count (the point itself included); core = nnb >= min_samples
  int32_t* minlab;    // [n] smallest cluster id among the core neighbours of a non-core point
  int32_t* parent;    // [n] union-find over core points (only the union sweep writes it)
  int32_t* rootof;    // [n] root of a core point, written once the unions are complete (never touched by cl_find)
  int32_t* rootlab;   // [n] cluster id of a root, -1 elsewhere
  int32_t* cl_cnt;    // [n] points per cluster
  int32_t* cl_min;    // [3n] ordered-int keys of the per-cluster minima
  int32_t* cl_max;    // [3n]
  int32_t* slot;      // [n] cluster id -> index among the kept clusters, or -1
  int32_t* blk;       // [2 * ceil(n / 256)] per-CTA counts of the two ordered compactions
};

__host__ __device__ inline int64_t cl_align(int64_t x) { return (x + 255) / 256 * 256; }

inline int64_t cluster_ws_layout(int64_t n, char* base, ClusterWs* w) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) { char* p = base ? base + off : nullptr; off += cl_align(bytes); return p; };
  const int64_t nb = (n + kClThreads - 1) / kClThreads;
  ClusterWs t;
  t.xyz = reinterpret_cast<float4*>(take(n * 16));
  t.nnb = reinterpret_cast<int32_t*>(take(n * 4));
  t.minlab = reinterpret_cast<int32_t*>(take(n * 4));
  t.parent = reinterpret_cast<int32_t*>(take(n * 4));
  t.rootof = reinterpret_cast<int32_t*>(take(n * 4));
  t.rootlab = reinterpret_cast<int32_t*>(take(n * 4));
  t.cl_cnt = reinterpret_cast<int32_t*>(take(n * 4));
  t.cl_min = reinterpret_cast<int32_t*>(take(n * 12));
  t.cl_max = reinterpret_cast<int32_t*>(take(n * 12));
  t.slot = reinterpret_cast<int32_t*>(take(n * 4));
  t.blk = reinterpret_cast<int32_t*>(take(nb * 8));
  if (w) *w = t;
  return off;
}

__device__ __forceinline__ int cl_key(float f) {  // monotonic float -> int
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float cl_unkey(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// exclusive position of this thread's flag inside the CTA and the CTA total; s_warp: kClWarps ints
__device__ __forceinline__ int cl_block_scan(bool flag, int* s_warp, int* total) {
  const unsigned ballot = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(ballot);
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < kClWarps; ++k) {
    const int c = s_warp[k];
    base += k < warp ? c : 0;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return base + __popc(ballot & smos_lanemask_lt());
}

// sum of the counts of the CTAs before this one (blk has gridDim.x entries)
__device__ __forceinline__ int cl_blocks_before(const int32_t* __restrict__ blk, int* s_red) {
  int part = 0;
  for (int c = threadIdx.x; c < static_cast<int>(blockIdx.x); c += kClThreads) part += blk[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
  __syncthreads();
  int base = 0;
#pragma unroll
  for (int k = 0; k < kClWarps; ++k) base += s_red[k];
  __syncthreads();
  return base;
}

// ---- foreground selection (voxel_instance_voting.py:145-148): indices where pred_bf == 2, ascending ----------
__global__ void __launch_bounds__(kClThreads)
cl_fg_count_kernel(const int32_t* __restrict__ bf, int64_t n, ClusterWs w) {
  SMOS_PDL_PROLOGUE();
  __shared__ int s_warp[kClWarps];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kClThreads + threadIdx.x;
  if (i < n) {  // per-cluster accumulators, reset for this call
    w.cl_cnt[i] = 0;
    w.cl_min[3 * i] = w.cl_min[3 * i + 1] = w.cl_min[3 * i + 2] = 0x7fffffff;
    w.cl_max[3 * i] = w.cl_max[3 * i + 1] = w.cl_max[3 * i + 2] = static_cast<int>(0x80000000u);
    w.slot[i] = -1;
    w.nnb[i] = 0;
    w.minlab[i] = 0x7fffffff;
    w.parent[i] = static_cast<int32_t>(i);
    w.rootof[i] = static_cast<int32_t>(i);
    w.rootlab[i] = -1;
  }
  int total;
  cl_block_scan(i < n && bf[i] == 2, s_warp, &total);
  if (threadIdx.x == 0) w.blk[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kClThreads)
cl_fg_write_kernel(const int32_t* __restrict__ bf, const float* __restrict__ pts, int64_t n, int64_t rs,
                   ClusterWs w, int32_t* __restrict__ fg_index, int32_t* __restrict__ fg_label,
                   int32_t* __restrict__ counts) {
  SMOS_PDL_PROLOGUE();
  __shared__ int s_warp[kClWarps];
  const int base = cl_blocks_before(w.blk, s_warp);
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kClThreads + threadIdx.x;
  const bool fg = i < n && bf[i] == 2;
  int total;
  const int pos = base + cl_block_scan(fg, s_warp, &total);
  if (fg) {
    const float* p = pts + i * rs;
    fg_index[pos] = static_cast<int32_t>(i);
    w.xyz[pos] = make_float4(p[0], p[1], p[2], 0.f);
  }
  if (i < n) fg_label[i] = -1;
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    counts[0] = base + total;  // M
    counts[1] = 0;
    counts[2] = 0;
  }
}

// ---- the pair sweep --------------------------------------------------------------------------------------------
// sklearn's KD-tree compares the reduced distance in float64: d = 0; d += dx*dx; d += dy*dy; d += dz*dz; d <= r*r
__device__ __forceinline__ bool cl_near(float xi, float yi, float zi, float4 q, float reject, double r2) {
  // float32 reject: |xi - xj| rounded is within 2^-24 relative of the exact difference; reject carries a 1e-3 margin
  if (fabsf(xi - q.x) > reject || fabsf(yi - q.y) > reject || fabsf(zi - q.z) > reject) return false;
  const double dx = static_cast<double>(xi) - static_cast<double>(q.x);
  const double dy = static_cast<double>(yi) - static_cast<double>(q.y);
  const double dz = static_cast<double>(zi) - static_cast<double>(q.z);
  double d = __dmul_rn(dx, dx);
  d = __dadd_rn(d, __dmul_rn(dy, dy));
  d = __dadd_rn(d, __dmul_rn(dz, dz));
  return d <= r2;
}

__device__ __forceinline__ int cl_find(int32_t* parent, int x) {
  volatile int32_t* p = parent;
  for (;;) {
    const int a = p[x];
    if (a == x) return x;
    const int g = p[a];
    if (g != a) p[x] = g;  // path halving; only ever lowers a non-root's pointer towards its root
    x = a;
  }
}

// Read-only walk to the root: for kernels that run after the union sweep, when parent[] no longer changes. No
// path-halving stores, so concurrent readers of parent[] in the same launch always see a consistent forest.
__device__ __forceinline__ int cl_find_ro(const int32_t* __restrict__ parent, int x) {
  for (;;) {
    const int a = parent[x];
    if (a == x) return x;
    x = a;
  }
}

// Joins the sets of i and j. `ri` is the caller's running guess of i's root (an ancestor of i, kept in a register
// across the neighbours of i): most neighbours of a point already hang under it, and one load of parent[j] settles
// those without walking any chain. Returns the new guess.
__device__ __forceinline__ int cl_unite(int32_t* parent, int ri, int j) {
  if (static_cast<volatile int32_t*>(parent)[j] == ri) return ri;
  int a = cl_find(parent, ri), b = cl_find(parent, j);
  while (a != b) {
    if (a < b) { const int t = a; a = b; b = t; }
    if (atomicCAS(&parent[a], a, b) == a) return b;  // the larger root hangs under the smaller one
    a = cl_find(parent, a);
    b = cl_find(parent, b);
  }
  return a;
}

enum { CL_COUNT = 0, CL_UNION = 1, CL_LABEL = 2 };
constexpr int kClColSplit = 16;  // grid.y: column tiles are dealt round-robin to this many CTAs per row tile

// One CTA = one tile of 256 rows against the column tiles y, y + 16, ... (the launch is sized for the upper bound
// n; CTAs beyond the M x M problem exit at once). Partial results meet in global atomics.
//   CL_COUNT: nnb[i] += neighbours in these columns          (nnb zeroed by cl_fg_count_kernel)
//   CL_UNION: core i, core j < i, near -> unite(i, j)        (column tiles up to the row tile only)
//   CL_LABEL: non-core i: minlab[i] = min(cluster id of near core j)
template <int MODE>
__global__ void __launch_bounds__(kClThreads)
cl_pairs_kernel(ClusterWs w, const int32_t* __restrict__ counts, float reject, double r2, int32_t min_samples) {
  SMOS_PDL_PROLOGUE();
  __shared__ float4 s_q[kClThreads];
  __shared__ int32_t s_aux[kClThreads];
  const int M = counts[0];
  const int row0 = blockIdx.x * kClThreads;
  if (row0 >= M) return;
  const int tile_end = MODE == CL_UNION ? static_cast<int>(blockIdx.x) + 1 : (M + kClThreads - 1) / kClThreads;
  if (static_cast<int>(blockIdx.y) >= tile_end) return;
  const int i = row0 + threadIdx.x;
  const bool live = i < M;
  float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
  bool core = false;
  if (live) {
    me = w.xyz[i];
    if (MODE != CL_COUNT) core = w.nnb[i] >= min_samples;
  }
  const bool work = MODE == CL_COUNT ? live : MODE == CL_UNION ? (live && core) : (live && !core);
  int acc = MODE == CL_LABEL ? 0x7fffffff : 0;
  int root = i;  // CL_UNION: running guess of i's root
  // the action for one candidate column t of the tile at j0
  auto visit = [&](int j0, int t, float4 q) {
    if (MODE == CL_UNION && (j0 + t >= i || !s_aux[t])) return;
    if (MODE == CL_LABEL && s_aux[t] < 0) return;
    if (!cl_near(me.x, me.y, me.z, q, reject, r2)) return;
    if (MODE == CL_COUNT) acc += 1;
    if (MODE == CL_UNION) root = cl_unite(w.parent, root, j0 + t);
    if (MODE == CL_LABEL) acc = min(acc, s_aux[t]);
  };
  auto far = [&](float4 q) {  // Chebyshev distance in float32: above `reject` means certainly not a neighbour
    return fmaxf(fmaxf(fabsf(me.x - q.x), fabsf(me.y - q.y)), fabsf(me.z - q.z));
  };
  for (int tj = blockIdx.y; tj < tile_end; tj += gridDim.y) {
    const int j0 = tj * kClThreads;
    const int j = j0 + threadIdx.x;
    __syncthreads();
    if (j < M) {
      s_q[threadIdx.x] = w.xyz[j];
      if (MODE == CL_UNION) s_aux[threadIdx.x] = w.nnb[j] >= min_samples;
      if (MODE == CL_LABEL) s_aux[threadIdx.x] = w.nnb[j] >= min_samples ? w.rootlab[w.rootof[j]] : -1;
    }
    __syncthreads();
    if (!work) continue;
    const int jn = min(kClThreads, M - j0);
    int t = 0;
    for (; t + 4 <= jn; t += 4) {  // four columns per step: nearly all of them fail the float32 test together
      const float4 q0 = s_q[t], q1 = s_q[t + 1], q2 = s_q[t + 2], q3 = s_q[t + 3];
      const float m0 = far(q0), m1 = far(q1), m2 = far(q2), m3 = far(q3);
      if (fminf(fminf(m0, m1), fminf(m2, m3)) > reject) continue;
      if (m0 <= reject) visit(j0, t, q0);
      if (m1 <= reject) visit(j0, t + 1, q1);
      if (m2 <= reject) visit(j0, t + 2, q2);
      if (m3 <= reject) visit(j0, t + 3, q3);
    }
    for (; t < jn; ++t) {
      const float4 q = s_q[t];
      if (far(q) <= reject) visit(j0, t, q);
    }
  }
  if (MODE == CL_COUNT && live && acc) atomicAdd(&w.nnb[i], acc);
  if (MODE == CL_LABEL && work && acc != 0x7fffffff) atomicMin(&w.minlab[i], acc);
}

// final label of every foreground point and the per-cluster size / extent
__global__ void __launch_bounds__(kClThreads)
cl_stats_kernel(ClusterWs w, const int32_t* __restrict__ counts, int32_t min_samples,
                int32_t* __restrict__ fg_label) {
  SMOS_PDL_PROLOGUE();
  const int M = counts[0];
  const int i = blockIdx.x * kClThreads + threadIdx.x;
  if (i >= M) return;
  // core points: the id of their component; border points: the smallest id among core neighbours; else noise
  const int ml = w.minlab[i];
  int lab = w.nnb[i] >= min_samples ? w.rootlab[w.rootof[i]] : (ml == 0x7fffffff ? -1 : ml);
  if (lab >= M) lab = -1;  // cannot happen (ids are ranks of roots); keeps the atomics below in bounds regardless
  fg_label[i] = lab;
  if (lab >= 0) {
    const float4 me = w.xyz[i];
    atomicAdd(&w.cl_cnt[lab], 1);
    atomicMin(&w.cl_min[3 * lab], cl_key(me.x)); atomicMax(&w.cl_max[3 * lab], cl_key(me.x));
    atomicMin(&w.cl_min[3 * lab + 1], cl_key(me.y)); atomicMax(&w.cl_max[3 * lab + 1], cl_key(me.y));
    atomicMin(&w.cl_min[3 * lab + 2], cl_key(me.z)); atomicMax(&w.cl_max[3 * lab + 2], cl_key(me.z));
  }
}

// ---- cluster ids: rank of each root among the roots (sklearn numbers clusters by their first core point) ------
__global__ void __launch_bounds__(kClThreads)
cl_root_count_kernel(ClusterWs w, const int32_t* __restrict__ counts, int32_t min_samples,
                     int32_t* __restrict__ blk) {
  SMOS_PDL_PROLOGUE();
  __shared__ int s_warp[kClWarps];
  const int M = counts[0];
  const int i = blockIdx.x * kClThreads + threadIdx.x;
  bool root = false;
  if (i < M && w.nnb[i] >= min_samples) {
    // the forest is final here: a read-only walk, and the flattened root goes to its own array — storing it into
    // parent[] would race with the walks of other threads of this launch
    const int r = cl_find_ro(w.parent, i);
    root = r == i;
    w.rootof[i] = r;
  }
  int total;
  cl_block_scan(root, s_warp, &total);
  if (threadIdx.x == 0) blk[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kClThreads)
cl_root_label_kernel(ClusterWs w, int32_t* __restrict__ counts, int32_t min_samples,
                     const int32_t* __restrict__ blk) {
  SMOS_PDL_PROLOGUE();
  __shared__ int s_warp[kClWarps];
  const int M = counts[0];
  const int base = cl_blocks_before(blk, s_warp);
  const int i = blockIdx.x * kClThreads + threadIdx.x;
  const bool root = i < M && w.nnb[i] >= min_samples && w.rootof[i] == i;
  int total;
  const int pos = base + cl_block_scan(root, s_warp, &total);
  if (root) w.rootlab[i] = pos;
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) counts[1] = base + total;  // clusters
}

// ---- boxes of the clusters with more than min_points points, in id order (:157-175) ----------------------------
__global__ void __launch_bounds__(kClThreads)
cl_boxes_kernel(ClusterWs w, int32_t* __restrict__ counts, int32_t min_points, float z_lift,
                float* __restrict__ box_lo, float* __restrict__ box_hi, int32_t* __restrict__ kept_label) {
  SMOS_PDL_PROLOGUE();
  __shared__ int s_warp[kClWarps];
  const int C = counts[1];
  int base = 0;
  for (int c0 = 0; c0 < C; c0 += kClThreads) {
    const int c = c0 + threadIdx.x;
    const bool keep = c < C && w.cl_cnt[c] > min_points;
    int total;
    const int k = base + cl_block_scan(keep, s_warp, &total);
    if (keep) {
      w.slot[c] = k;
      kept_label[k] = c;
      const float zmin = cl_unkey(w.cl_min[3 * c + 2]), zmax = cl_unkey(w.cl_max[3 * c + 2]);
      // corners whose z equals the minimum get += 0.2 in float32 (:172-174); the box is then spanned by the
      // lifted floor and the ceiling, whichever is lower (a cluster thinner than the lift flips them)
      const float lift = __fadd_rn(zmin, z_lift);
      const float zl = zmin == zmax ? lift : fminf(lift, zmax);
      const float zh = zmin == zmax ? lift : fmaxf(lift, zmax);
      box_lo[3 * k] = cl_unkey(w.cl_min[3 * c]);
      box_lo[3 * k + 1] = cl_unkey(w.cl_min[3 * c + 1]);
      box_lo[3 * k + 2] = zl;
      box_hi[3 * k] = cl_unkey(w.cl_max[3 * c]);
      box_hi[3 * k + 1] = cl_unkey(w.cl_max[3 * c + 1]);
      box_hi[3 * k + 2] = zh;
    }
    base += total;
  }
  if (threadIdx.x == 0) counts[2] = base;
}

// ---- write-back (:184-191): every point of a kept cluster takes the label its box voted -------------------------
__global__ void __launch_bounds__(kClThreads)
cl_apply_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ fg_index,
                const int32_t* __restrict__ fg_label, const int32_t* __restrict__ slot,
                const int64_t* __restrict__ sums, int64_t* __restrict__ pred) {
  SMOS_PDL_PROLOGUE();
  const int M = counts[0];
  const int i = blockIdx.x * kClThreads + threadIdx.x;
  if (i >= M) return;
  const int lab = fg_label[i];
  if (lab < 0) return;
  const int k = slot[lab];
  if (k < 0) return;
  // static_points_num = sum(pred[pred==1]), dynamic_points_num = sum(pred[pred==2]); 2 if dynamic > static else 1
  pred[fg_index[i]] = sums[2 * k + 1] > sums[2 * k] ? 2 : 1;
}

}  // namespace

extern "C" {

int64_t smos_cluster_workspace_bytes(int64_t n) {
  if (n < 0) return -1;
  return cluster_ws_layout(n > 0 ? n : 1, nullptr, nullptr);
}

int smos_cluster_boxes(const float* points, int64_t n, int64_t row_stride, const int32_t* pred_bf, double eps,
                       int32_t min_samples, int32_t min_cluster_points, float z_lift, void* workspace,
                       int32_t* fg_index, int32_t* fg_label, int32_t* counts, float* box_lo, float* box_hi,
                       int32_t* kept_label, void* stream) {
  if (n < 0 || n > 0x7fffffff || row_stride < 3 || !(eps > 0.0) || min_samples < 1 ||
      min_cluster_points < 0)
    return SMOS_EINVAL;
  if (!counts) return SMOS_EINVAL;
  cudaStream_t st = smos_stream(stream);
  if (n == 0) return static_cast<int>(cudaMemsetAsync(counts, 0, 3 * sizeof(int32_t), st));
  if (!points || !pred_bf || !workspace || !fg_index || !fg_label || !box_lo || !box_hi || !kept_label)
    return SMOS_EINVAL;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return SMOS_EINVAL;
  ClusterWs w;
  cluster_ws_layout(n, static_cast<char*>(workspace), &w);
  const int grid = smos_ceil_div(n, kClThreads);
  const float reject = static_cast<float>(eps * 1.001);
  const double r2 = eps * eps;
  int32_t* blk_roots = w.blk + grid;
  SMOS_LAUNCH((cl_fg_count_kernel), grid, kClThreads, 0, st, pred_bf, n, w);
  SMOS_LAUNCH((cl_fg_write_kernel), grid, kClThreads, 0, st, pred_bf, points, n, row_stride, w, fg_index, fg_label, counts);
  const dim3 pgrid(grid, kClColSplit);
  SMOS_LAUNCH((cl_pairs_kernel<CL_COUNT>), pgrid, kClThreads, 0, st, w, counts, reject, r2, min_samples);
  SMOS_LAUNCH((cl_pairs_kernel<CL_UNION>), pgrid, kClThreads, 0, st, w, counts, reject, r2, min_samples);
  SMOS_LAUNCH((cl_root_count_kernel), grid, kClThreads, 0, st, w, counts, min_samples, blk_roots);
  SMOS_LAUNCH((cl_root_label_kernel), grid, kClThreads, 0, st, w, counts, min_samples, blk_roots);
  SMOS_LAUNCH((cl_pairs_kernel<CL_LABEL>), pgrid, kClThreads, 0, st, w, counts, reject, r2, min_samples);
  SMOS_LAUNCH((cl_stats_kernel), grid, kClThreads, 0, st, w, counts, min_samples, fg_label);
  SMOS_LAUNCH((cl_boxes_kernel), 1, kClThreads, 0, st, w, counts, min_cluster_points, z_lift, box_lo, box_hi, kept_label);
  return smos_launch_status();
}

int smos_cluster_apply(int64_t n, void* workspace, const int32_t* fg_index, const int32_t* fg_label,
                       const int32_t* counts, const int64_t* sums, int64_t* pred, void* stream) {
  if (n < 0 || n > 0x7fffffff) return SMOS_EINVAL;
  if (n == 0) return SMOS_OK;
  if (!workspace || !fg_index || !fg_label || !counts || !sums || !pred) return SMOS_EINVAL;
  ClusterWs w;
  cluster_ws_layout(n, static_cast<char*>(workspace), &w);
  SMOS_LAUNCH((cl_apply_kernel), smos_ceil_div(n, kClThreads), kClThreads, 0, smos_stream(stream), counts, fg_index, fg_label,
                                                                                       w.slot, sums, pred);
  return smos_launch_status();
}

}  // extern "C"
