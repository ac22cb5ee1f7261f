/*
 * smos_oracle.c — CPU restatement of the StreamMOS hot path. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library. The product (streammos_b200/) never does.
 *
 * Every function restates the reference algorithm in plain scalar C and cites the reference
 * file:line it follows (paths relative to the StreamMOS tree). Pinning: the reference ships no
 * golden vectors for these ops (its only test, deformattn/test.py, compares the CUDA op with
 * ms_deform_attn_core_pytorch at run time), so this oracle is pinned against OUTPUTS OF THE
 * REFERENCE ITSELF, generated in the build container by tools/make_golden.py (the reference's
 * point_deep.cpp compiled unmodified, its ms_deform_attn_core_pytorch, its BilinearSample module,
 * its voting functions and one whole frame of the voxel_voting.py loop body with its Trans and Crop,
 * its PointNetStacker module, its utils.Quantize / SphereQuantize / make_point_feat) and committed
 * under tests/golden/. Status: pinned (tests/test_oracle_golden.py).
 *
 * Third-party arithmetic restated here: torch's grid_sampler_2d (bilinear, zeros padding) —
 * the reference pins torch 1.11.0 (README.md:65,79); the algorithm restated is ATen's
 * aten/src/ATen/native/GridSampler.h (unnormalize + bilinear weights + bounds mask), whose
 * results in the image's torch 2.11 the golden vectors record.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------------- */
/* VoxelMaxPool — deep_point/src/point_deep.cpp:19-88 (Init loop then Max loop),            */
/* voxel_max_idx per deep_point/src/point_deep_cuda_kernel.cu:24-53.                         */
/* feat (B,C,N) ind (B,N,2) out (B,C,H,W) idx (B,N) or NULL                                   */
/* ---------------------------------------------------------------------------------------- */
static int cell_of(const float* ind2, float sh, float sw, int64_t H, int64_t W, int64_t* cell) {
  /* int64_t(static_cast<float>(ind) * scale): fp32 product, truncation (point_deep.cpp:38) */
  volatile float fh = ind2[0] * sh;
  volatile float fw = ind2[1] * sw;
  if (!(fh > -9.0e18f && fh < 9.0e18f && fw > -9.0e18f && fw < 9.0e18f)) return 0; /* NaN/overflow */
  int64_t ih = (int64_t)fh, iw = (int64_t)fw;
  if (ih < 0 || ih >= H || iw < 0 || iw >= W) return 0;
  *cell = ih * W + iw;
  return 1;
}

API void oracle_voxel_maxpool_forward(const float* feat, const float* ind, int64_t B, int64_t C, int64_t N,
                                      int64_t H, int64_t W, float sh, float sw, float* out, int64_t* idx) {
  const int64_t hw = H * W;
  memset(out, 0, sizeof(float) * (size_t)(B * C * hw)); /* torch.zeros, deep_point/__init__.py:26 */
  if (idx)
    for (int64_t i = 0; i < B * N; ++i) idx[i] = -1; /* torch.full(-1), :27 */
  /* VoxelMaxPoolUpdateOutputInit, point_deep.cpp:19-52 */
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c)
      for (int64_t n = 0; n < N; ++n) {
        int64_t cell;
        if (cell_of(ind + (b * N + n) * 2, sh, sw, H, W, &cell)) {
          out[(b * C + c) * hw + cell] = feat[(b * C + c) * N + n];
          if (idx && c == 0) idx[b * N + n] = b * C * hw + cell;
        }
      }
  /* VoxelMaxPoolUpdateOutputKernel, point_deep.cpp:54-88 */
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c)
      for (int64_t n = 0; n < N; ++n) {
        int64_t cell;
        if (cell_of(ind + (b * N + n) * 2, sh, sw, H, W, &cell)) {
          float* o = out + (b * C + c) * hw + cell;
          const float f = feat[(b * C + c) * N + n];
          if (*o < f) *o = f;
        }
      }
}

/* VoxelMaxPoolUpdateBackwardKernel, point_deep.cpp:97-132; grad_feat zero-filled (:50 of __init__.py) */
API void oracle_voxel_maxpool_backward(const float* feat, const float* ind, const float* out, const float* gout,
                                       int64_t B, int64_t C, int64_t N, int64_t H, int64_t W, float sh, float sw,
                                       float* gfeat) {
  const int64_t hw = H * W;
  memset(gfeat, 0, sizeof(float) * (size_t)(B * C * N));
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c)
      for (int64_t n = 0; n < N; ++n) {
        int64_t cell;
        if (cell_of(ind + (b * N + n) * 2, sh, sw, H, W, &cell)) {
          const int64_t o = (b * C + c) * hw + cell;
          if (out[o] == feat[(b * C + c) * N + n]) gfeat[(b * C + c) * N + n] = gout[o];
        }
      }
}

/* ---------------------------------------------------------------------------------------- */
/* BilinearSample — networks/backbone.py:458-475:                                            */
/*   gx = 2*coord[:,:,1]*scale[1]/(W-1) - 1 ; gy likewise with H ; F.grid_sample(bilinear,   */
/*   zeros, align_corners=True). grid_sampler: pix = ((g+1)/2)*(size-1); nw = floor;          */
/*   weights (x_se-x)(y_se-y) ... ; taps outside the image contribute 0.                      */
/* grid (B,C,H,W) coord (B,N,2) out (B,C,N); all fp32 arithmetic, no contraction.            */
/* ---------------------------------------------------------------------------------------- */
static float replay_pixel(float c, float s, int64_t size) {
  volatile float sm1 = (float)(size - 1);
  volatile float g = 2.0f * c;
  g = g * s;
  g = g / sm1;
  g = g - 1.0f;
  volatile float p = g + 1.0f;
  p = p / 2.0f;
  p = p * sm1;
  return p;
}

API void oracle_bilinear_sample(const float* grid, int64_t B, int64_t C, int64_t H, int64_t W, const float* coord,
                                int64_t N, float sh, float sw, float* out) {
  for (int64_t b = 0; b < B; ++b)
    for (int64_t n = 0; n < N; ++n) {
      const float ix = replay_pixel(coord[(b * N + n) * 2 + 1], sw, W);
      const float iy = replay_pixel(coord[(b * N + n) * 2 + 0], sh, H);
      const float fx = floorf(ix), fy = floorf(iy);
      volatile float dx1 = (fx + 1.0f) - ix, dy1 = (fy + 1.0f) - iy, dx0 = ix - fx, dy0 = iy - fy;
      volatile float w_nw = dx1 * dy1, w_ne = dx0 * dy1, w_sw = dx1 * dy0, w_se = dx0 * dy0;
      int64_t x0 = -2, y0 = -2;
      if (fx == fx && fx >= -2.0f && fx <= (float)W + 1.0f) x0 = (int64_t)fx;
      else if (fx > (float)W + 1.0f) x0 = W + 1;
      if (fy == fy && fy >= -2.0f && fy <= (float)H + 1.0f) y0 = (int64_t)fy;
      else if (fy > (float)H + 1.0f) y0 = H + 1;
      const int xl = x0 >= 0 && x0 < W, xh = x0 + 1 >= 0 && x0 + 1 < W;
      const int yl = y0 >= 0 && y0 < H, yh = y0 + 1 >= 0 && y0 + 1 < H;
      for (int64_t c = 0; c < C; ++c) {
        const float* g = grid + (b * C + c) * H * W;
        float acc = 0.0f;
        if (xl && yl) acc = fmaf(g[y0 * W + x0], w_nw, acc);
        if (xh && yl) acc = fmaf(g[y0 * W + x0 + 1], w_ne, acc);
        if (xl && yh) acc = fmaf(g[(y0 + 1) * W + x0], w_sw, acc);
        if (xh && yh) acc = fmaf(g[(y0 + 1) * W + x0 + 1], w_se, acc);
        out[(b * C + c) * N + n] = acc;
      }
    }
}

/* grid_sampler_2d backward wrt input only (coordinates carry no gradient in the model). */
API void oracle_bilinear_sample_backward(const float* gout, int64_t B, int64_t C, int64_t N, const float* coord,
                                         float sh, float sw, int64_t H, int64_t W, double* ggrid) {
  memset(ggrid, 0, sizeof(double) * (size_t)(B * C * H * W));
  for (int64_t b = 0; b < B; ++b)
    for (int64_t n = 0; n < N; ++n) {
      const float ix = replay_pixel(coord[(b * N + n) * 2 + 1], sw, W);
      const float iy = replay_pixel(coord[(b * N + n) * 2 + 0], sh, H);
      const float fx = floorf(ix), fy = floorf(iy);
      if (!(fx == fx) || !(fy == fy) || fx < -2.0f || fy < -2.0f || fx > (float)W + 1.0f || fy > (float)H + 1.0f)
        continue;
      const int64_t x0 = (int64_t)fx, y0 = (int64_t)fy;
      const double dx0 = (double)ix - fx, dy0 = (double)iy - fy, dx1 = 1.0 - dx0, dy1 = 1.0 - dy0;
      const int xl = x0 >= 0 && x0 < W, xh = x0 + 1 >= 0 && x0 + 1 < W;
      const int yl = y0 >= 0 && y0 < H, yh = y0 + 1 >= 0 && y0 + 1 < H;
      for (int64_t c = 0; c < C; ++c) {
        double* g = ggrid + (b * C + c) * H * W;
        const double go = gout[(b * C + c) * N + n];
        if (xl && yl) g[y0 * W + x0] += go * dx1 * dy1;
        if (xh && yl) g[y0 * W + x0 + 1] += go * dx0 * dy1;
        if (xl && yh) g[(y0 + 1) * W + x0] += go * dx1 * dy0;
        if (xh && yh) g[(y0 + 1) * W + x0 + 1] += go * dx0 * dy0;
      }
    }
}

/* ---------------------------------------------------------------------------------------- */
/* MSDeformAttn — ms_deform_attn_core_pytorch, deformattn/functions/ms_deform_attn_func.py:41-61: */
/*   grids = 2*loc - 1 ; grid_sample(value_l, bilinear, zeros, align_corners=False)             */
/*   (pix = ((g+1)*size - 1)/2) ; out = sum_{l,p} sampled * attn.  In double precision.         */
/* Backward follows the analytic gradients of the same expression (the reference's CUDA         */
/* col2im, deformattn/src/cuda/ms_deform_im2col_cuda.cuh:87-160, computes the same quantities). */
/* value (B,S,M,D) shapes (L,2) lsi (L) loc (B,Q,M,L,P,2) attn (B,Q,M,L,P) out (B,Q,M*D)         */
/* ---------------------------------------------------------------------------------------- */
API void oracle_ms_deform_attn_forward(const double* value, const int64_t* shapes, const int64_t* lsi,
                                       const double* loc, const double* attn, int64_t B, int64_t S, int64_t M,
                                       int64_t D, int64_t L, int64_t Q, int64_t P, double* out) {
  for (int64_t b = 0; b < B; ++b)
    for (int64_t q = 0; q < Q; ++q)
      for (int64_t m = 0; m < M; ++m) {
        const int64_t g = (b * Q + q) * M + m;
        double* o = out + g * D;
        for (int64_t d = 0; d < D; ++d) o[d] = 0.0;
        for (int64_t l = 0; l < L; ++l) {
          const int64_t H = shapes[2 * l], W = shapes[2 * l + 1];
          for (int64_t p = 0; p < P; ++p) {
            const double lx = loc[((g * L + l) * P + p) * 2], ly = loc[((g * L + l) * P + p) * 2 + 1];
            const double a = attn[(g * L + l) * P + p];
            const double gx = 2.0 * lx - 1.0, gy = 2.0 * ly - 1.0;
            const double ix = ((gx + 1.0) * (double)W - 1.0) / 2.0, iy = ((gy + 1.0) * (double)H - 1.0) / 2.0;
            const double fx = floor(ix), fy = floor(iy);
            if (!(fx > -3.0 && fx < (double)W + 2.0 && fy > -3.0 && fy < (double)H + 2.0)) continue;
            const int64_t x0 = (int64_t)fx, y0 = (int64_t)fy;
            const double dx0 = ix - fx, dy0 = iy - fy, dx1 = 1.0 - dx0, dy1 = 1.0 - dy0;
            for (int64_t t = 0; t < 4; ++t) {
              const int64_t x = x0 + (t & 1), y = y0 + (t >> 1);
              if (x < 0 || x >= W || y < 0 || y >= H) continue;
              const double w = ((t & 1) ? dx0 : dx1) * ((t >> 1) ? dy0 : dy1);
              const double* v = value + ((b * S + lsi[l] + y * W + x) * M + m) * D;
              for (int64_t d = 0; d < D; ++d) o[d] += a * w * v[d];
            }
          }
        }
      }
}

API void oracle_ms_deform_attn_backward(const double* value, const int64_t* shapes, const int64_t* lsi,
                                        const double* loc, const double* attn, const double* gout, int64_t B,
                                        int64_t S, int64_t M, int64_t D, int64_t L, int64_t Q, int64_t P,
                                        double* gvalue, double* gloc, double* gattn) {
  memset(gvalue, 0, sizeof(double) * (size_t)(B * S * M * D));
  for (int64_t b = 0; b < B; ++b)
    for (int64_t q = 0; q < Q; ++q)
      for (int64_t m = 0; m < M; ++m) {
        const int64_t g = (b * Q + q) * M + m;
        const double* go = gout + g * D;
        for (int64_t l = 0; l < L; ++l) {
          const int64_t H = shapes[2 * l], W = shapes[2 * l + 1];
          for (int64_t p = 0; p < P; ++p) {
            const int64_t sidx = (g * L + l) * P + p;
            const double a = attn[sidx];
            const double ix = loc[sidx * 2] * (double)W - 0.5, iy = loc[sidx * 2 + 1] * (double)H - 0.5;
            double g_x = 0.0, g_y = 0.0, g_a = 0.0;
            const double fx = floor(ix), fy = floor(iy);
            if (fx > -3.0 && fx < (double)W + 2.0 && fy > -3.0 && fy < (double)H + 2.0) {
              const int64_t x0 = (int64_t)fx, y0 = (int64_t)fy;
              const double dx0 = ix - fx, dy0 = iy - fy, dx1 = 1.0 - dx0, dy1 = 1.0 - dy0;
              for (int64_t t = 0; t < 4; ++t) {
                const int64_t x = x0 + (t & 1), y = y0 + (t >> 1);
                if (x < 0 || x >= W || y < 0 || y >= H) continue;
                const double wx = (t & 1) ? dx0 : dx1, wy = (t >> 1) ? dy0 : dy1;
                const double sx = (t & 1) ? 1.0 : -1.0, sy = (t >> 1) ? 1.0 : -1.0;
                const int64_t off = ((b * S + lsi[l] + y * W + x) * M + m) * D;
                for (int64_t d = 0; d < D; ++d) {
                  const double v = value[off + d];
                  gvalue[off + d] += go[d] * a * wx * wy;
                  g_a += go[d] * wx * wy * v;
                  g_x += go[d] * a * sx * wy * v * (double)W;
                  g_y += go[d] * a * wx * sy * v * (double)H;
                }
              }
            }
            gloc[sidx * 2] = g_x;
            gloc[sidx * 2 + 1] = g_y;
            gattn[sidx] = g_a;
          }
        }
      }
}

/* ---------------------------------------------------------------------------------------- */
/* Voting — voxel_voting.py:38-91 (== voxel_instance_voting.py:78-135)                       */
/* ---------------------------------------------------------------------------------------- */

/* Quantize, voxel_voting.py:77-91: fp32 (x - min) / d */
API void oracle_quantize(const float* pcds, int64_t P, int64_t row_stride, float mx, float my, float mz, float dx,
                         float dy, float dz, float* out) {
  for (int64_t i = 0; i < P; ++i) {
    volatile float x = pcds[i * row_stride] - mx, y = pcds[i * row_stride + 1] - my,
                   z = pcds[i * row_stride + 2] - mz;
    out[i * 3] = x / dx;
    out[i * 3 + 1] = y / dy;
    out[i * 3 + 2] = z / dz;
  }
}

/* determine_voxel_labels, voxel_voting.py:55-75: votes[lin, label] += 1 ; argmax(-1)
 * (first maximum = lowest class on ties; all-zero row -> 0). Returns -1 on allocation failure. */
API int oracle_determine_voxel_labels(const int64_t* coords, const int64_t* labels, int64_t P, int64_t X,
                                      int64_t Y, int64_t Z, int64_t num_classes, int64_t* voxel_labels) {
  const int64_t V = X * Y * Z;
  int32_t* votes = (int32_t*)calloc((size_t)(V * num_classes), sizeof(int32_t));
  if (!votes) return -1;
  for (int64_t i = 0; i < P; ++i) {
    const int64_t x = coords[i * 3], y = coords[i * 3 + 1], z = coords[i * 3 + 2];
    if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) continue; /* reference input is pre-cropped */
    if (labels[i] < 0 || labels[i] >= num_classes) continue;
    votes[(x * Y * Z + y * Z + z) * num_classes + labels[i]] += 1; /* :67,:71 */
  }
  for (int64_t v = 0; v < V; ++v) {
    int64_t best = 0;
    for (int64_t c = 1; c < num_classes; ++c)
      if (votes[v * num_classes + c] > votes[v * num_classes + best]) best = c;
    voxel_labels[v] = best; /* :73 */
  }
  free(votes);
  return 0;
}

/* get_point_labels_from_voxel_labels, voxel_voting.py:38-53 */
API void oracle_point_labels(const int64_t* coords, int64_t Pc, const int64_t* voxel_labels, int64_t X, int64_t Y,
                             int64_t Z, int64_t* out) {
  for (int64_t i = 0; i < Pc; ++i) {
    const int64_t x = coords[i * 3], y = coords[i * 3 + 1], z = coords[i * 3 + 2];
    out[i] = 0;
    if (x >= 0 && y >= 0 && z >= 0 && x < X && y < Y && z < Z) out[i] = voxel_labels[x * Y * Z + y * Z + z];
  }
}

/* Vote block of cluster(), voxel_instance_voting.py:177-187, with in_hull(:62-76) of the 8 AABB
 * corners restated as the inclusive box test it is equivalent to. sums (K,2). */
API void oracle_instance_vote(const float* pts, int64_t P, int64_t row_stride, const int64_t* pred,
                              const float* lo, const float* hi, int64_t K, int64_t* sums) {
  for (int64_t k = 0; k < K; ++k) {
    int64_t st = 0, dy = 0;
    for (int64_t i = 0; i < P; ++i) {
      const float* p = pts + i * row_stride;
      if (p[0] >= lo[k * 3] && p[0] <= hi[k * 3] && p[1] >= lo[k * 3 + 1] && p[1] <= hi[k * 3 + 1] &&
          p[2] >= lo[k * 3 + 2] && p[2] <= hi[k * 3 + 2]) {
        if (pred[i] == 1) st += 1; /* sum(pred[pred == 1]) */
        if (pred[i] == 2) dy += 2; /* sum(pred[pred == 2]) : 2 per point */
      }
    }
    sums[k * 2] = st;
    sums[k * 2 + 1] = dy;
  }
}

/* ---------------------------------------------------------------------------------------- */
/* DBSCAN as cluster() runs it (voxel_instance_voting.py:150-153:                              */
/*   DBSCAN(eps=0.3, min_samples=5).fit_predict(foreground_points), float32 xyz).               */
/* The algorithm lives in a third-party dependency that is not vendored in the reference:      */
/* scikit-learn (requirements.txt names it without a version; this image has 1.9.0). Restated   */
/* from its published sources:                                                                  */
/*   sklearn/cluster/_dbscan.py  fit(): neighborhoods = radius_neighbors(X, eps) (a point is    */
/*     its own neighbour), core = n_neighbors >= min_samples;                                   */
/*   sklearn/neighbors (KDTree, float64): a point is a neighbour when the reduced distance      */
/*     ((0 + dx*dx) + dy*dy) + dz*dz, accumulated in float64 in that order, is <= eps*eps;      */
/*   sklearn/cluster/_dbscan_inner.pyx  dbscan_inner(): depth-first expansion with a stack,     */
/*     seeds visited in index order, label_num incremented after each expansion.                */
/* Pinned by tests/golden/cluster_a.npz (labels produced by sklearn through the reference's      */
/* own cluster()) and, where sklearn is importable, against fit_predict directly.               */
/* ---------------------------------------------------------------------------------------- */
static int dbscan_near(const float* a, const float* b, double r2) {
  const double dx = (double)a[0] - (double)b[0], dy = (double)a[1] - (double)b[1], dz = (double)a[2] - (double)b[2];
  double d = 0.0;
  d += dx * dx;
  d += dy * dy;
  d += dz * dz;
  return d <= r2;
}

API int oracle_dbscan(const float* pts, int64_t M, int64_t row_stride, double eps, int64_t min_samples,
                      int32_t* labels) {
  const double r2 = eps * eps;
  int64_t* begin = (int64_t*)calloc((size_t)M + 1, sizeof(int64_t));
  if (!begin) return 1;
  for (int64_t i = 0; i < M; ++i) {
    int64_t c = 0;
    for (int64_t j = 0; j < M; ++j) c += dbscan_near(pts + i * row_stride, pts + j * row_stride, r2);
    begin[i + 1] = begin[i] + c;
  }
  const int64_t nnz = begin[M];
  int32_t* nbr = (int32_t*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t));
  int32_t* stack = (int32_t*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t));
  if (!nbr || !stack) { free(nbr); free(stack); free(begin); return 1; }
  for (int64_t i = 0; i < M; ++i) {
    int64_t w = begin[i];
    for (int64_t j = 0; j < M; ++j)
      if (dbscan_near(pts + i * row_stride, pts + j * row_stride, r2)) nbr[w++] = (int32_t)j;
  }
  for (int64_t i = 0; i < M; ++i) labels[i] = -1;
  int32_t label_num = 0;
  for (int64_t seed = 0; seed < M; ++seed) {
    if (labels[seed] != -1 || begin[seed + 1] - begin[seed] < min_samples) continue;
    int64_t top = 0, i = seed;
    for (;;) { /* dbscan_inner: label, push unlabelled neighbours of core points, pop */
      if (labels[i] == -1) {
        labels[i] = label_num;
        if (begin[i + 1] - begin[i] >= min_samples)
          for (int64_t k = begin[i]; k < begin[i + 1]; ++k)
            if (labels[nbr[k]] == -1) stack[top++] = nbr[k];
      }
      if (top == 0) break;
      i = stack[--top];
    }
    label_num += 1;
  }
  free(stack);
  free(nbr);
  free(begin);
  return 0;
}

/* ---------------------------------------------------------------------------------------- */
/* Streaming long-term voting, one frame — the loop body of voxel_voting.py:176-244:          */
/*   Trans (datasets/utils.py:116-126): float64 pose_diff . (x,y,z,1) -> float32 (numpy's dgemm */
/*   accumulates the four products in k order with FMAs; restated as an fma chain);            */
/*   Crop (utils/transforms.py:151-161): keep lo+eps < p < hi-eps, compared in float32;         */
/*   Quantize + .to(int64) + determine_voxel_labels + get_point_labels_from_voxel_labels;       */
/*   current_pred_result_orin[mask] = pred_result_new (voxel_voting.py:243-244).                */
/* scans are concatenated: pts (sum n, rs), lab (sum n), begin[n_scans+1]; m[n_scans][12].     */
/* ---------------------------------------------------------------------------------------- */
static int stream_point(const float* p, const double* m, int transform, const float* lo, const float* hi,
                        float* q) {
  float x = p[0], y = p[1], z = p[2];
  if (transform) {
    const double dx = x, dy = y, dz = z;
    x = (float)fma(m[3], 1.0, fma(m[2], dz, fma(m[1], dy, m[0] * dx)));
    y = (float)fma(m[7], 1.0, fma(m[6], dz, fma(m[5], dy, m[4] * dx)));
    z = (float)fma(m[11], 1.0, fma(m[10], dz, fma(m[9], dy, m[8] * dx)));
  }
  q[0] = x; q[1] = y; q[2] = z;
  return x > lo[0] && x < hi[0] && y > lo[1] && y < hi[1] && z > lo[2] && z < hi[2];
}

static int quant_lin(const float* q, const float* mn, const float* d, int64_t X, int64_t Y, int64_t Z, int64_t* lin) {
  volatile float fx = q[0] - mn[0], fy = q[1] - mn[1], fz = q[2] - mn[2];
  volatile float gx = fx / d[0], gy = fy / d[1], gz = fz / d[2];
  const int64_t x = (int64_t)gx, y = (int64_t)gy, z = (int64_t)gz;
  if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) return 0;
  *lin = x * Y * Z + y * Z + z;
  return 1;
}

API int oracle_vote_stream(const float* pts, const uint8_t* lab, const int64_t* begin, const double* m,
                           const int32_t* transform, int64_t n_scans, int64_t current, int64_t rs, const float* lo,
                           const float* hi, const float* mn, const float* d, int64_t X, int64_t Y, int64_t Z,
                           int64_t num_classes, uint8_t* voxel_labels, int64_t* point_labels, float* trans_out) {
  const int64_t V = X * Y * Z;
  int32_t* votes = (int32_t*)calloc((size_t)(V * num_classes), sizeof(int32_t));
  if (!votes) return -1;
  for (int64_t j = 0; j < n_scans; ++j)
    for (int64_t i = begin[j]; i < begin[j + 1]; ++i) {
      float q[3];
      const int in = stream_point(pts + i * rs, m + j * 12, transform[j], lo, hi, q);
      if (trans_out) { trans_out[i * 3] = q[0]; trans_out[i * 3 + 1] = q[1]; trans_out[i * 3 + 2] = q[2]; }
      int64_t lin;
      if (in && lab[i] < num_classes && quant_lin(q, mn, d, X, Y, Z, &lin)) votes[lin * num_classes + lab[i]] += 1;
    }
  for (int64_t v = 0; v < V; ++v) {
    int64_t best = 0;
    for (int64_t c = 1; c < num_classes; ++c)
      if (votes[v * num_classes + c] > votes[v * num_classes + best]) best = c;
    voxel_labels[v] = (uint8_t)best;
  }
  free(votes);
  for (int64_t i = begin[current]; i < begin[current + 1]; ++i) {
    float q[3];
    int64_t r = lab[i];
    if (stream_point(pts + i * rs, m + current * 12, transform[current], lo, hi, q)) {
      int64_t lin;
      r = quant_lin(q, mn, d, X, Y, Z, &lin) ? voxel_labels[lin] : 0;
    }
    point_labels[i - begin[current]] = r;
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------- */
/* PointNet stem — networks/backbone.py:199-250 PointNetStacker(cin, 64,   */
/* pre_bn=True, stack_num=2) in eval mode as models/StreamMOS.py:77,101 runs it:              */
/*   BatchNorm2d -> Conv2d 1x1 (no bias) -> BatchNorm2d -> ReLU -> Conv2d 1x1 -> BatchNorm2d   */
/*   -> ReLU. Eval BatchNorm = per-channel affine y = x*alpha + beta (alpha = weight /         */
/*   sqrt(running_var + eps), beta = bias - running_mean*alpha, as ATen's CPU kernel folds it).*/
/* x (B, Cin, N), w1 (C1, Cin), w2 (C2, C1), y (B, C2, N); a0/b0 may be NULL.                  */
/* ---------------------------------------------------------------------------------------- */
API void oracle_point_stem(const float* x, int64_t B, int64_t Cin, int64_t N, const float* a0, const float* b0,
                           const float* w1, const float* a1, const float* b1, int64_t C1, const float* w2,
                           const float* a2, const float* b2, int64_t C2, float* y) {
  float* xin = (float*)malloc(sizeof(float) * (size_t)Cin);
  float* h = (float*)malloc(sizeof(float) * (size_t)C1);
  for (int64_t b = 0; b < B; ++b)
    for (int64_t n = 0; n < N; ++n) {
      for (int64_t ci = 0; ci < Cin; ++ci) {
        const float v = x[(b * Cin + ci) * N + n];
        xin[ci] = a0 ? fmaf(v, a0[ci], b0[ci]) : v;
      }
      for (int64_t c = 0; c < C1; ++c) {
        float acc = 0.f;
        for (int64_t ci = 0; ci < Cin; ++ci) acc = fmaf(w1[c * Cin + ci], xin[ci], acc);
        const float t = fmaf(acc, a1[c], b1[c]);
        h[c] = t > 0.f ? t : 0.f;
      }
      for (int64_t c = 0; c < C2; ++c) {
        float acc = 0.f;
        for (int64_t k = 0; k < C1; ++k) acc = fmaf(w2[c * C1 + k], h[k], acc);
        const float t = fmaf(acc, a2[c], b2[c]);
        y[(b * C2 + c) * N + n] = t > 0.f ? t : 0.f;
      }
    }
  free(xin);
  free(h);
}

/* ---------------------------------------------------------------------------------------- */
/* Loader form_batch without SphereQuantize (exact part):                   */
/*   utils.Quantize (datasets/utils.py:151-169): (v - min) / d in float32;                    */
/*   make_point_feat (datasets/data_StreamMOS.py:25-50): x, y, z, intensity,                   */
/*   dist = sqrt(x**2 + y**2 + z**2) + 1e-12, diff = coord - floor(coord), all float32;        */
/*   TTA flips of form_batch_tta (:495-513) as sign factors.                                   */
/* pts (T*N, rs), feat (T, 7, N), coord (T, N, 3).                                             */
/* ---------------------------------------------------------------------------------------- */
API void oracle_form_batch(const float* pts, int64_t T, int64_t N, int64_t rs, float sx, float sy, const float* mn,
                           const float* d, float* feat, float* coord) {
  for (int64_t t = 0; t < T; ++t)
    for (int64_t n = 0; n < N; ++n) {
      const float* p = pts + (t * N + n) * rs;
      volatile float x = p[0] * sx, y = p[1] * sy, z = p[2];
      volatile float ax = x - mn[0], ay = y - mn[1], az = z - mn[2];
      volatile float qx = ax / d[0], qy = ay / d[1], qz = az / d[2];
      volatile float xx = x * x, yy = y * y, zz = z * z;
      volatile float s1 = xx + yy;
      volatile float s2 = s1 + zz;
      volatile float r = sqrtf(s2);
      volatile float dist = r + 1e-12f;
      float* f = feat + t * 7 * N + n;
      f[0] = x; f[N] = y; f[2 * N] = z; f[3 * N] = p[3]; f[4 * N] = dist;
      f[5 * N] = qx - floorf(qx);
      f[6 * N] = qy - floorf(qy);
      float* c = coord + (t * N + n) * 3;
      c[0] = qx; c[1] = qy; c[2] = qz;
    }
}

/* ---------------------------------------------------------------------------------------- */
/* utils.SphereQuantize (datasets/utils.py:172-192), the range-view coordinates of the         */
/* loader's form_batch (datasets/data_StreamMOS.py:481-484), numpy float32 throughout:         */
/*   d = sqrt(x**2 + y**2 + z**2) + 1e-12; phi = phi_hi - arctan2(x, y);                        */
/*   theta = theta_hi - arcsin(z / d); out = (theta / dtheta, phi / dphi).                      */
/* FLOATING POINT: numpy's float32 arctan2 / arcsin (SVML or libm, by host) are not correctly  */
/* rounded, so this restatement takes the float64 functions rounded to float32 once — the      */
/* correctly rounded angle in all but ~1e-9 of the cases — and tests/test_oracle_golden.py      */
/* bounds its distance to the reference's own output (1 ulp of the angle).                     */
/* pts (total, rs) -> out (total, 2); c = {phi_hi, theta_hi, dphi, dtheta} as float32.         */
/* ---------------------------------------------------------------------------------------- */
API void oracle_sphere_quantize(const float* pts, int64_t total, int64_t rs, float sx, float sy, const float* c,
                                float* out) {
  for (int64_t i = 0; i < total; ++i) {
    const float* p = pts + i * rs;
    volatile float x = p[0] * sx, y = p[1] * sy, z = p[2];
    volatile float xx = x * x, yy = y * y, zz = z * z;
    volatile float s1 = xx + yy;
    volatile float s2 = s1 + zz;
    volatile float r = sqrtf(s2);
    volatile float dist = r + 1e-12f;
    volatile float q = z / dist;
    volatile float a_phi = (float)atan2((double)x, (double)y);
    volatile float a_theta = (float)asin((double)q);
    volatile float phi = c[0] - a_phi;
    volatile float theta = c[1] - a_theta;
    out[2 * i] = theta / c[3];
    out[2 * i + 1] = phi / c[2];
  }
}

/* ---------------------------------------------------------------------------------------- */
/* Scan ingestion: one frame of the val loader's window,                   */
/* datasets/data_StreamMOS.py:515-574:                                                         */
/*   utils.Trans (datasets/utils.py:116-126): pcds_tmp = mat.dot([x, y, z, 1]) in float64      */
/*   (numpy's dgemm: the four products accumulated in k order with FMAs), xyz stored back as  */
/*   float32, intensity untouched;                                                             */
/*   utils.filter_pcds_mask (:107-113): lo <= p < hi on the aligned float32 point;             */
/*   pc_list[ht][valid_mask]: order-preserving compaction;                                     */
/*   np.pad(..., constant_values=-1000) and [-pad_length:, 2] = -4000 (:566-571).              */
/* pts (n, rs) -> out (n_out, 4), src (n_out) raw row of every output row (-1 = padding).     */
/* Returns the number of points that passed the filter (the loader asserts it is < n_out).    */
/* ---------------------------------------------------------------------------------------- */
API int64_t oracle_ingest_frame(const float* pts, int64_t n, int64_t rs, const double* m, int transform,
                                const float* lo, const float* hi, int64_t n_out, float pad_xy, float pad_z,
                                float* out, int32_t* src) {
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float* p = pts + i * rs;
    float x = p[0], y = p[1], z = p[2];
    if (transform) {
      const double dx = x, dy = y, dz = z;
      x = (float)fma(m[3], 1.0, fma(m[2], dz, fma(m[1], dy, m[0] * dx)));
      y = (float)fma(m[7], 1.0, fma(m[6], dz, fma(m[5], dy, m[4] * dx)));
      z = (float)fma(m[11], 1.0, fma(m[10], dz, fma(m[9], dy, m[8] * dx)));
    }
    if (x >= lo[0] && x < hi[0] && y >= lo[1] && y < hi[1] && z >= lo[2] && z < hi[2]) {
      if (k < n_out) {
        out[k * 4] = x; out[k * 4 + 1] = y; out[k * 4 + 2] = z; out[k * 4 + 3] = p[3];
        if (src) src[k] = (int32_t)i;
      }
      ++k;
    }
  }
  for (int64_t r = k; r < n_out; ++r) {
    out[r * 4] = pad_xy; out[r * 4 + 1] = pad_xy; out[r * 4 + 2] = pad_z; out[r * 4 + 3] = pad_xy;
    if (src) src[r] = -1;
  }
  return k;
}
