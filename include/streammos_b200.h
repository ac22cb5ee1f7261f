/*
 * streammos_b200.h — C-ABI of the B200 (sm_100a) StreamMOS hot-path library.
 *
 * Drop-in boundary for the per-scan data-parallel hot path of StreamMOS:
 *   (A) point -> BEV / range-view scatter-max pooling and bilinear gather-back,
 *   (B) multi-scale deformable-attention sampling (forward / backward),
 *   (C) long-term-memory voxel voting and per-instance vote counting.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - strides are in ELEMENTS, not bytes;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - no allocation, no synchronisation, no host<->device copy happens inside
 *     the library: the caller owns all buffers including scratch (`plan`,
 *     `workspace`), whose sizes come from the *_bytes() functions;
 *   - return value: 0 on success, a negative SMOS_E* code for bad arguments,
 *     or a positive cudaError_t if a launch failed. smos_error_string() maps
 *     either to text. The Python shims raise RuntimeError on non-zero.
 *
 * Each entry point cites the reference interface (file:line under the
 * StreamMOS tree) it replaces.
 */
#ifndef STREAMMOS_B200_H_
#define STREAMMOS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMOS_OK 0
#define SMOS_EINVAL (-1)       /* bad shape / null pointer / misaligned */
#define SMOS_EUNSUPPORTED (-2) /* valid request the library does not implement */

#define SMOS_ABI_VERSION 2

int smos_abi_version(void);
const char* smos_error_string(int code);

/* Identity of the CUDA-graph capture `stream` is currently part of (cudaStreamGetCaptureInfo): *id_host = 0 when
 * the stream is not capturing, else the capture sequence's unique id. Host-side query, no synchronisation. The
 * Python plan cache uses it so that a pooling plan built eagerly is never baked into a graph (and vice versa). */
int smos_stream_capture_id(void* stream, uint64_t* id_host);

/* ------------------------------------------------------------------------- */
/* (A1) VoxelMaxPool — scatter-max of point features into a dense 2-D grid.   */
/* Replaces point_deep.cuda_kernel.voxel_maxpooling_forward/backward          */
/* (deep_point/src/point_deep_cuda.cpp:22-62,                                  */
/*  deep_point/src/point_deep_cuda_kernel.cu:24-186) and the allocation +     */
/* metadata upload in deep_point/__init__.py:17-44.                            */
/* ------------------------------------------------------------------------- */

/* Bytes of scratch a pooling plan needs for B x N points on an (H, W) grid. */
int64_t smos_pool_plan_bytes(int64_t B, int64_t N, int32_t H, int32_t W);

/* Bytes of scratch one forward call needs (piece-maxima rows; touched sparsely). */
int64_t smos_pool_workspace_bytes(int64_t B, int64_t C, int64_t N);

/* Build the pooling plan for one coordinate array:
 *   cell(b,n) = trunc(float(ind[b,n,0]) * scale_h) * W + trunc(ind[b,n,1] * scale_w)
 *   valid iff 0 <= idx_d < size_d for both d (C cast = truncation toward zero,
 *   point_deep_cuda_kernel.cu:40-46); invalid points get cell = -1.
 * Then the valid points are counting-sorted by cell (per-cell count / start / sorted list), which
 * both the forward reduction and the dense output writer consume.
 *   pcds_ind       : (B, N, 2) float32, element strides ind_sb / ind_sn / ind_sd
 *   voxel_max_idx  : optional (B, N) int64 out — the reference's side product:
 *                    b*C*H*W + h*W + w (flat NCHW offset of the c=0 plane) or -1
 *                    (point_deep_cuda_kernel.cu:36,50; deep_point/__init__.py:27).
 *                    `idx_batch_stride` = C*H*W. May be NULL.
 *   plan           : scratch of smos_pool_plan_bytes(), 16-byte aligned.
 */
int smos_pool_plan_build(const float* pcds_ind, int64_t B, int64_t N,
                         int64_t ind_sb, int64_t ind_sn, int64_t ind_sd,
                         int32_t H, int32_t W, float scale_h, float scale_w,
                         int64_t* voxel_max_idx, int64_t idx_batch_stride,
                         void* plan, void* stream);

/* Several plans in one go (four kernel launches for all of them instead of four each): a scan
 * needs five — BEV at scales 1, 1/2, 1/4 and range view at 1/2, 1/4. Same semantics per entry as
 * smos_pool_plan_build. n <= 8. If the plan buffers are carved out of one allocation they are
 * cleared by the first of the four kernels (no memset nodes). */
typedef struct smos_pool_plan_desc {
  const float* pcds_ind;
  int64_t B, N, ind_sb, ind_sn, ind_sd;
  int32_t H, W;
  float scale_h, scale_w;
  int64_t* voxel_max_idx; /* may be NULL */
  int64_t idx_batch_stride;
  void* plan;
  void* gather_taps; /* may be NULL; else smos_gather_taps_bytes(B, N) bytes, 16-byte aligned: the BilinearSample
                        sampling state of every point for THIS grid and scale (networks/backbone.py:458-475), one
                        48-byte record per slot of the plan's cell order, for smos_bilinear_gather_forward_taps */
  const float* scale_dev; /* may be NULL; else two float32 ON THE DEVICE that override scale_h / scale_w: the
                             reference hands `scale_rate` to point_deep.cuda_kernel as a device tensor
                             (deep_point/__init__.py:31, point_deep_cuda.cpp:22-37) — reading it on the device
                             avoids a device->host copy and its synchronisation */
} smos_pool_plan_desc;

int smos_pool_plan_build_multi(const smos_pool_plan_desc* descs_host, int32_t n, void* stream);

/* Forward: voxel_out[b,c,h,w] = max over points of the cell, 0 for empty cells
 * (true max even if negative: point_deep_cuda_kernel.cu:56-99).
 *   pcds_feat : (B, C, N) float32, element strides f_sb / f_sc / f_sn
 *               (channel-major N-fastest and point-major C-fastest both run
 *               at full coalescing).
 *   workspace : scratch of smos_pool_workspace_bytes(B, C, N), 256-byte aligned.
 *   voxel_out : (B, C, H, W) float32 NCHW-contiguous; EVERY element is written
 *               (no pre-zeroing needed).
 */
int smos_voxel_maxpool_forward(const float* pcds_feat, int64_t B, int64_t C, int64_t N,
                               int64_t f_sb, int64_t f_sc, int64_t f_sn,
                               int32_t H, int32_t W, const void* plan, void* workspace,
                               float* voxel_out, void* stream);

/* Same call restricted to some of its stages (profiling / benchmarking of a single kernel; the
 * stages must have run once in order before a later stage is launched alone):
 *   REDUCE  [permute +] piece reduction   COMBINE  fold multi-piece cells   WRITE  dense output */
#define SMOS_POOL_STAGE_REDUCE 1
#define SMOS_POOL_STAGE_COMBINE 2
#define SMOS_POOL_STAGE_WRITE 4
#define SMOS_POOL_STAGE_ALL 7
#define SMOS_POOL_STAGE_NO_REDUCE 8 /* with REDUCE on a channel-major input: run the permute only (profiling) */
int smos_voxel_maxpool_forward_stages(const float* pcds_feat, int64_t B, int64_t C, int64_t N,
                                      int64_t f_sb, int64_t f_sc, int64_t f_sn,
                                      int32_t H, int32_t W, const void* plan, void* workspace,
                                      float* voxel_out, int32_t stages, void* stream);

/* Backward: grad_feat[b,c,n] = grad_out[b,c,cell] if voxel_out[b,c,cell] == feat[b,c,n]
 * else 0 — every tied point receives the gradient (point_deep_cuda_kernel.cu:109-132).
 *   grad_feat : (B, C, N) float32 with strides g_sb / g_sc / g_sn; every element written.
 */
int smos_voxel_maxpool_backward(const float* pcds_feat, int64_t B, int64_t C, int64_t N,
                                int64_t f_sb, int64_t f_sc, int64_t f_sn,
                                int32_t H, int32_t W, const void* plan,
                                const float* voxel_out, const float* grad_voxel_out,
                                float* grad_feat, int64_t g_sb, int64_t g_sc, int64_t g_sn,
                                void* stream);

/* ------------------------------------------------------------------------- */
/* (A3) BilinearSample — bilinear gather of grid features back to points.     */
/* Replaces networks/backbone.py:458-475 (normalise, F.grid_sample bilinear,   */
/* zeros padding, align_corners=True). The float sequence of the reference is  */
/* replayed: g = 2*c*s/(size-1) - 1 ; pix = ((g+1)/2)*(size-1).               */
/* ------------------------------------------------------------------------- */

/*   grid  : (B, C, H, W) float32, element strides gr_sb / gr_sc / gr_sh / gr_sw
 *           (NCHW and channels-last both supported)
 *   coord : (B, N, 2) float32, strides co_sb / co_sn / co_sd ; d=0 -> row (H), d=1 -> col (W)
 *   out   : (B, C, N) float32, strides o_sb / o_sc / o_sn
 */
int smos_bilinear_gather_forward(const float* grid, int64_t B, int64_t C, int32_t H, int32_t W,
                                 int64_t gr_sb, int64_t gr_sc, int64_t gr_sh, int64_t gr_sw,
                                 const float* coord, int64_t N,
                                 int64_t co_sb, int64_t co_sn, int64_t co_sd,
                                 float scale_h, float scale_w,
                                 float* out, int64_t o_sb, int64_t o_sc, int64_t o_sn,
                                 void* stream);

/* Same result, points visited in the cell order of a pooling plan built from the SAME coordinate array
 * (B, N as here; any grid plan_H x plan_W and scale): `plan` lists the points grouped by cell, out-of-grid
 * points at the tail, so a warp's taps share one or two cache lines per channel plane instead of one line
 * per point (a BEV arc in scan order crosses image rows at every point). Used when `out` is point-major
 * (o_sc == 1); otherwise identical to smos_bilinear_gather_forward. The reference model pools and gathers
 * with the same coordinates (networks/multi_view_encoder.py:395-417), so the plan exists anyway. */
int smos_bilinear_gather_forward_ordered(const float* grid, int64_t B, int64_t C, int32_t H, int32_t W,
                                         int64_t gr_sb, int64_t gr_sc, int64_t gr_sh, int64_t gr_sw,
                                         const float* coord, int64_t N,
                                         int64_t co_sb, int64_t co_sn, int64_t co_sd,
                                         float scale_h, float scale_w,
                                         float* out, int64_t o_sb, int64_t o_sc, int64_t o_sn,
                                         const void* plan, int32_t plan_H, int32_t plan_W, void* stream);

/* Same result again from sampling records that a pooling plan build emitted as a by-product
 * (smos_pool_plan_desc.gather_taps: same coordinates, H, W and scale as this gather): the kernel's prologue is one
 * 48-byte load per point instead of order entry -> coordinates -> replayed pixel arithmetic. Points are visited in
 * the plan's cell order; `out` must be point-major (o_sc == 1). N = points per batch entry of the plan. */
int64_t smos_gather_taps_bytes(int64_t B, int64_t N);
int smos_bilinear_gather_forward_taps(const float* grid, int64_t B, int64_t C, int32_t H, int32_t W,
                                      int64_t gr_sb, int64_t gr_sc, int64_t gr_sh, int64_t gr_sw,
                                      const void* taps, int64_t N,
                                      float* out, int64_t o_sb, int64_t o_sc, int64_t o_sn, void* stream);

/* Gradient wrt the grid (grid_sampler backward). grad_grid (B, C, H, W) NCHW-contiguous; the function zero-fills
 * it itself (no fill by the caller) and accumulates the contributions with fp32 atomics, so sums are exact up to
 * fp32 addition order like the reference's. Two kernels. */
int smos_bilinear_gather_backward(const float* grad_out, int64_t B, int64_t C, int64_t N,
                                  int64_t go_sb, int64_t go_sc, int64_t go_sn,
                                  const float* coord,
                                  int64_t co_sb, int64_t co_sn, int64_t co_sd,
                                  float scale_h, float scale_w,
                                  int32_t H, int32_t W, float* grad_grid, void* stream);

/* ------------------------------------------------------------------------- */
/* (B) MSDeformAttn sampling core.                                            */
/* Replaces MultiScaleDeformableAttention.ms_deform_attn_forward/backward      */
/* (deformattn/src/vision.cpp:13-16, src/ms_deform_attn.h:20-61,               */
/*  src/cuda/ms_deform_attn_cuda.cu:20-153, src/cuda/ms_deform_im2col_cuda.cuh */
/*  :237-299 forward, :301-403 backward).                                      */
/*   value            : (B, S, M, D) contiguous                                */
/*   spatial_shapes   : (L, 2) int64 DEVICE [(H_l, W_l)]                       */
/*   level_start_index: (L,) int64 DEVICE                                      */
/*   sampling_loc     : (B, Q, M, L, P, 2) contiguous, (x, y) normalised       */
/*   attn_weight      : (B, Q, M, L, P) contiguous                             */
/*   output           : (B, Q, M*D); every element written                     */
/* dtype: 0 = float32, 1 = float64 (the reference dispatches both).            */
/* ------------------------------------------------------------------------- */
#define SMOS_F32 0
#define SMOS_F64 1

int smos_ms_deform_attn_forward(int32_t dtype, const void* value,
                                const int64_t* spatial_shapes,
                                const int64_t* level_start_index,
                                const void* sampling_loc, const void* attn_weight,
                                int32_t B, int32_t S, int32_t M, int32_t D,
                                int32_t L, int32_t Q, int32_t P,
                                void* output, void* stream);

/* (next: SURVEY 8f rank 4, fusing) The same sampling core started one step    */
/* earlier, at what the attention module computes in front of it               */
/* (deformattn/modules/ms_deform_attn.py:96-108):                              */
/*   attn_weight  = softmax(attn_logits over the L*P samples of a (b, q, m))   */
/*   sampling_loc = reference_points[:, :, None, :, None, :]                    */
/*                  + sampling_offsets / (W_l, H_l)                 (ref_dim 2) */
/*                = ref[..., :2] + sampling_offsets / P * ref[..., 2:] * 0.5    */
/*                                                                  (ref_dim 4) */
/*   sampling_offsets : (B, Q, M, L, P, 2)   attn_logits : (B, Q, M, L*P)       */
/*   reference_points : (B, Q, L, ref_dim)   all contiguous, value's dtype      */
int smos_ms_deform_attn_fused_forward(int32_t dtype, const void* value,
                                      const int64_t* spatial_shapes,
                                      const int64_t* level_start_index,
                                      const void* sampling_offsets, const void* attn_logits,
                                      const void* reference_points, int32_t ref_dim,
                                      int32_t B, int32_t S, int32_t M, int32_t D,
                                      int32_t L, int32_t Q, int32_t P,
                                      void* output, void* stream);

/* grad_value must be ZERO-FILLED by the caller (atomic accumulation);
 * grad_sampling_loc and grad_attn_weight are fully written. */
int smos_ms_deform_attn_backward(int32_t dtype, const void* value,
                                 const int64_t* spatial_shapes,
                                 const int64_t* level_start_index,
                                 const void* sampling_loc, const void* attn_weight,
                                 const void* grad_output,
                                 int32_t B, int32_t S, int32_t M, int32_t D,
                                 int32_t L, int32_t Q, int32_t P,
                                 void* grad_value, void* grad_sampling_loc,
                                 void* grad_attn_weight, void* stream);

/* ------------------------------------------------------------------------- */
/* (D, next: SURVEY 8f rank 4) PointNet stem in front of VoxelMaxPool #1.      */
/* Replaces networks/backbone.py:199-250 PointNetStacker(Cin, 64, pre_bn=True, */
/* stack_num=2) in eval mode (models/StreamMOS.py:77,101):                     */
/*   h = relu(bn1(W1 . bn0(x))) ; y = relu(bn2(W2 . h))                         */
/* with every eval BatchNorm given as the per-channel affine torch applies:     */
/*   alpha = weight / sqrt(running_var + eps), beta = bias - running_mean*alpha */
/* ------------------------------------------------------------------------- */

/*   x  : (B, Cin, N) float32, strides x_sb / x_sc / x_sn ; Cin <= 16
 *   w1 : (C1, Cin) row major ; w2 : (C2, C1) row major ; C1 == C2 == 64
 *   bn0_alpha / bn0_beta may both be NULL (no pre-BN)
 *   y  : (B, C2, N) float32, element strides y_sb / y_sc / y_sn (channel-major y_sn == 1, or point-major y_sc == 1: each
 *        point's C2 features contiguous, what VoxelMaxPool reads without its permute stage); every element written.
 * Layer 2 runs on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split, fp32 accumulation in TMEM): within 1e-5 of
 * the fp32 reference. The environment variable SMOS_STEM_UMMA=0 selects the CUDA-core kernel (fixed FMA order, y_sn == 1). */
int smos_point_stem_forward(const float* x, int64_t B, int32_t Cin, int64_t N,
                            int64_t x_sb, int64_t x_sc, int64_t x_sn,
                            const float* bn0_alpha, const float* bn0_beta, const float* w1,
                            const float* bn1_alpha, const float* bn1_beta, const float* w2,
                            const float* bn2_alpha, const float* bn2_beta, int32_t C1, int32_t C2,
                            float* y, int64_t y_sb, int64_t y_sc, int64_t y_sn, void* stream);

/* ------------------------------------------------------------------------- */
/* (E, next: SURVEY 8f rank 2, exact part) model input tensors from raw scans. */
/* Replaces utils.Quantize (datasets/utils.py:151-169) + make_point_feat        */
/* (datasets/data_StreamMOS.py:25-50) inside the loader's form_batch (:471-493) */
/* and the TTA flips of form_batch_tta (:495-513). SphereQuantize is a separate, */
/* floating-point entry below (numpy's arctan2 / arcsin cannot be matched bit    */
/* for bit).                                                                     */
/* ------------------------------------------------------------------------- */

/*   points     : (T*N, row_stride>=4) float32 raw x, y, z, intensity of T frames (range filtered, padded)
 *   x_sign/y_sign : TTA flip factors (+1 / -1)
 *   min_*, d*  : Quantize origin and cell size, d = float32((range_hi - range_lo) / size)
 *   pcds_xyzi  : (T, 7, N) float32 out: x, y, z, intensity, dist, diff_x, diff_y
 *   pcds_coord : (T, N, 3) float32 out: x_quan, y_quan, z_quan
 * Bit-exact with the reference's numpy float32 arithmetic. */
int smos_form_batch(const float* points, int64_t T, int64_t N, int64_t row_stride, float x_sign, float y_sign,
                    float min_x, float min_y, float min_z, float dx, float dy, float dz,
                    float* pcds_xyzi, float* pcds_coord, void* stream);

/* (next: SURVEY 8f rank 2) utils.SphereQuantize (datasets/utils.py:172-192), the range-view coordinates of form_batch
 * (datasets/data_StreamMOS.py:481-484): d = sqrt(x*x + y*y + z*z) + 1e-12, phi = phi_hi - arctan2(x, y),
 * theta = theta_hi - arcsin(z / d), out = (theta / dtheta, phi / dphi), numpy float32 throughout.
 *   points       : (total, row_stride >= 3) float32 rows x, y, z, ...
 *   x_sign/y_sign: TTA flip factors (+1 / -1)
 *   phi_hi, theta_hi : float32(range[1] * pi / 180); dphi, dtheta : float32((hi_rad - lo_rad) / W or H)
 *   sphere_coord : (total, 2) float32 out (theta_quan, phi_quan), 8-byte aligned
 * FLOATING POINT, not bit-exact: numpy's float32 arctan2 / arcsin are not correctly rounded and differ between hosts;
 * here the two angles are evaluated in float64 and rounded to float32 once, all other operations replayed in float32.
 * Against the reference: |d phi_quan| <= 5e-4 cells, |d theta_quan| <= 5e-5 cells (1 ulp of the angle), identical on
 * most points; about one point per 120 k scan changes its range-view cell. */
int smos_sphere_quantize(const float* points, int64_t total, int64_t row_stride, float x_sign, float y_sign,
                         float phi_hi, float theta_hi, float dphi, float dtheta, float* sphere_coord, void* stream);

/* (next: SURVEY 8f rank 2) Scan ingestion: the loader steps in front of form_batch, per frame of the T-frame window
 * (datasets/data_StreamMOS.py:515-574): utils.Trans (datasets/utils.py:116-126: float64 pose_diff . (x, y, z, 1) ->
 * float32), utils.filter_pcds_mask (:107-113: lo <= p < hi on the aligned point), order-preserving compaction and
 * padding to n_out rows with (pad_xy, pad_xy, pad_z, pad_xy) = (-1000, -1000, -4000, -1000). Bit-exact against those
 * functions. Lets a stream keep the RAW scans of its window in HBM: only the new scan and the poses cross PCIe.
 *   frame.points  : (n_cap, row_floats >= 4) float32 raw scan (x, y, z, intensity) in its own sensor frame
 *   frame.n_dev   : DEVICE int32: rows of `points` that hold points (<= n_cap)        } read by the kernels, so the
 *   frame.pose_dev: DEVICE 12 float64: rows 0..2 of pose_diff, row major; NULL = none } launches are graph-replayable
 *   out_points    : (T, n_out, 4) float32, 16-byte aligned; out_src: (T, n_out) int32 raw row of each output row
 *                   (-1 = padding) or NULL; out_count: (T) int32 points that passed the filter or NULL. The loader
 *                   asserts count < n_out; here a frame that does not fit is truncated and out_count tells. T <= 8. */
typedef struct smos_ingest_frame {
  const float* points;
  const int32_t* n_dev;
  const double* pose_dev;
  int64_t n_cap;
} smos_ingest_frame;

int64_t smos_ingest_workspace_bytes(int32_t T, int64_t n_cap_max, int64_t n_out);
int smos_ingest_frames(const smos_ingest_frame* frames_host, int32_t T, int64_t row_floats,
                       float x_lo, float x_hi, float y_lo, float y_hi, float z_lo, float z_hi,
                       int64_t n_out, float pad_xy, float pad_z, void* workspace,
                       float* out_points, int32_t* out_src, int32_t* out_count, void* stream);

/* smos_form_batch followed by smos_point_stem_forward as ONE kernel (the (T, 7, N) tensor never exists): raw points
 * in, pcds_coord (T, N, 3) and the 64-channel features y (T, C2, N) out; results bit-identical to the two calls. */
int smos_point_stem_forward_raw(const float* points, int64_t T, int64_t N, int64_t row_stride,
                                float x_sign, float y_sign, float min_x, float min_y, float min_z,
                                float dx, float dy, float dz,
                                const float* bn0_alpha, const float* bn0_beta, const float* w1,
                                const float* bn1_alpha, const float* bn1_beta, const float* w2,
                                const float* bn2_alpha, const float* bn2_beta, int32_t C1, int32_t C2,
                                float* pcds_coord, float* y, int64_t y_sb, int64_t y_sc, int64_t y_sn, void* stream);

/* The same call with a cap on the kernel's persistent CTAs (one per SM; 0 = all SMs). A scheduling hint, results are
 * identical: the stem is latency bound and holds its SM (165 KB of shared memory) against every other kernel, so a
 * stream that keeps several scans in flight runs it on ~60 % of the SMs and lets the HBM-bound kernels of the
 * neighbouring scans have the rest (measured: +3-4 % end to end); a caller that runs one kernel at a time wants 0. */
int smos_point_stem_forward_raw_capped(const float* points, int64_t T, int64_t N, int64_t row_stride,
                                float x_sign, float y_sign, float min_x, float min_y, float min_z,
                                float dx, float dy, float dz,
                                const float* bn0_alpha, const float* bn0_beta, const float* w1,
                                const float* bn1_alpha, const float* bn1_beta, const float* w2,
                                const float* bn2_alpha, const float* bn2_beta, int32_t C1, int32_t C2,
                                float* pcds_coord, float* y, int64_t y_sb, int64_t y_sc, int64_t y_sn, int32_t max_ctas, void* stream);

/* ------------------------------------------------------------------------- */
/* (C) Long-term-memory voting.                                               */
/* ------------------------------------------------------------------------- */

/* Quantize (voxel_voting.py:77-91; voxel_instance_voting.py:117-135):
 *   out[p,d] = (pcds[p,d] - min_d) / delta_d   in float32, IEEE division.
 *   pcds : (P, row_stride>=3) float32 ; out : (P, 3) float32 contiguous. */
int smos_quantize(const float* pcds, int64_t P, int64_t row_stride,
                  float min_x, float min_y, float min_z,
                  float dx, float dy, float dz, float* out, void* stream);

/* The same Quantize with the arithmetic torch applies ON A CUDA DEVICE — where the reference's scripts evaluate it
 * (voxel_voting.py:218-240): tensor / Python scalar = tensor * float32(1 / scalar) (ATen BinaryDivTrueKernel.cu,
 * is_cpu_scalar branch). Bit-identical with `(pcds[:, d] - min_d) / delta_d` run by torch on the GPU; differs from
 * smos_quantize (= numpy / torch-CPU, IEEE division) in the last bit of some quotients. Same arguments. */
int smos_quantize_rcp(const float* pcds, int64_t P, int64_t row_stride,
                  float min_x, float min_y, float min_z,
                  float dx, float dy, float dz, float* out, void* stream);

/* Bytes of scratch for smos_vote_voxel_labels / smos_vote_fused / smos_vote_stream. */
int64_t smos_vote_workspace_bytes(int64_t P, int32_t X, int32_t Y, int32_t Z, int32_t num_classes);

/* determine_voxel_labels (voxel_voting.py:55-75): per-voxel class histogram over
 * the local map, argmax with ties -> lowest class, empty voxel -> 0.
 *   voxel_coords : (P, 3) int64 contiguous ; semantic_labels : (P,) int64
 *   voxel_labels : (X, Y, Z) int64 contiguous; every element written.
 * Points whose coordinates fall outside the grid or whose label is outside
 * [0, num_classes) are ignored (the reference assumes pre-cropped input). */
int smos_vote_voxel_labels(const int64_t* voxel_coords, const int64_t* semantic_labels,
                           int64_t P, int32_t X, int32_t Y, int32_t Z, int32_t num_classes,
                           void* workspace, int64_t* voxel_labels, void* stream);

/* get_point_labels_from_voxel_labels (voxel_voting.py:38-53): bounds mask +
 * gather; out-of-range points get 0.  (sx, sy, sz) are the label grid dims. */
int smos_vote_point_labels(const int64_t* new_voxel_coords, int64_t Pc,
                           const int64_t* voxel_labels, int32_t X, int32_t Y, int32_t Z,
                           int64_t* point_labels, void* stream);

/* Fused long-term voting for the streaming path (same results
 * as Quantize -> .to(int64) -> determine_voxel_labels -> get_point_labels...):
 *   points (P, row_stride) float32, labels (P,) uint8, the LAST `Pc` rows are the
 *   current scan. Writes voxel_labels_u8 (X*Y*Z) and point_labels (Pc,) int64. */
int smos_vote_fused(const float* points, int64_t P, int64_t row_stride,
                    const uint8_t* labels, int64_t Pc,
                    float min_x, float min_y, float min_z, float dx, float dy, float dz,
                    int32_t X, int32_t Y, int32_t Z, int32_t num_classes,
                    void* workspace, uint8_t* voxel_labels_u8, int64_t* point_labels,
                    void* stream);

/* Streaming long-term memory (replaces the per-frame loop body of voxel_voting.py:176-244):
 * the last scans stay resident in HBM; one call per frame pose-aligns the history into the current frame
 * (datasets/utils.py:116-126 Trans: float64 pose_diff x (x,y,z,1) -> float32), crops history and current to
 * the open box crop_lo < p < crop_hi (utils/transforms.py:151-161; thresholds already include eps), quantises
 * (voxel_voting.py:77-91), votes per voxel, and labels the current scan: a point inside the crop takes its
 * voxel's majority label, a point outside keeps its own prediction (voxel_voting.py:243-244).
 *   scans_host[j] : device points (n, row_stride>=3) f32, device labels (n,) u8, pose_diff = rows 0..2 of
 *                   inv(pose_current) . pose_j (row major, float64), transform = 0 for the current scan
 *   workspace     : smos_vote_workspace_bytes(sum n, X, Y, Z, num_classes) bytes
 *   voxel_labels_u8 (X*Y*Z) and point_labels (n of scan `current`) are fully written. n_scans <= 16. */
typedef struct smos_vote_stream_scan {
  const float* points;
  const uint8_t* labels;
  int64_t n;
  double pose_diff[12];
  int32_t transform;
} smos_vote_stream_scan;

int smos_vote_stream(const smos_vote_stream_scan* scans_host, int32_t n_scans, int32_t current,
                     int64_t row_stride, const float* crop_lo_host, const float* crop_hi_host,
                     float min_x, float min_y, float min_z, float dx, float dy, float dz,
                     int32_t X, int32_t Y, int32_t Z, int32_t num_classes,
                     void* workspace, uint8_t* voxel_labels_u8, int64_t* point_labels, void* stream);

/* Long-term memory ring insert (the bookkeeping around voxel_voting.py:176-194 once the scans stay in HBM):
 * the scan held in the "current" slot moves to its history slot (skipped when hist_* are NULL) and the new scan
 * takes the current slot. points (n, row_floats) f32, pred (n,) u8; slots have the same shapes. One kernel. */
int smos_memory_push(const float* points, const uint8_t* pred, int64_t n, int64_t row_floats,
                     float* cur_points, uint8_t* cur_pred, float* hist_points, uint8_t* hist_pred, void* stream);

/* Staging for the reference's int64 voting API, one kernel (voxel_voting.py:234-241): for every point of the
 * long-term memory ring, q = Quantize(point) (float32, as smos_quantize), coords = q.to(int64) (truncation) and
 * labels = pred.to(int64) — the three tensors the script builds with Quantize and two `.to(torch.int64)` casts.
 * Optionally performs smos_memory_push on the way (new_points / new_pred non-NULL): the scan in slot `cur_slot`
 * moves to `hist_slot` (skipped if hist_slot < 0) and the new scan takes `cur_slot`, before anything is quantised.
 *   ring_points (n_slots, n, row_floats) f32, ring_pred (n_slots, n) u8, slot major, updated in place when pushing
 *   q_out (n_slots*n, 3) f32 or NULL ; coords_out (n_slots*n, 3) int64 ; labels_out (n_slots*n) int64
 *   crop_lo_host / crop_hi_host: 3 floats each or NULL. The script crops the local map to the open box
 *   lo < p < hi (transforms.Crop, utils/transforms.py:151-161, thresholds fov -/+ eps compared in float32) BEFORE it
 *   quantises (voxel_voting.py:225-231); tensors cannot shrink inside a captured launch, so a point outside the box
 *   keeps its slot and gets coords (-1, -1, -1): it casts no vote and reads no voxel label, exactly as if it had
 *   been removed (q keeps the plain quotient). */
int smos_vote_stage(float* ring_points, uint8_t* ring_pred, int32_t n_slots, int64_t n, int64_t row_floats,
                    const float* new_points, const uint8_t* new_pred, int32_t cur_slot, int32_t hist_slot,
                    float min_x, float min_y, float min_z, float dx, float dy, float dz,
                    const float* crop_lo_host, const float* crop_hi_host,
                    float* q_out, int64_t* coords_out, int64_t* labels_out, void* stream);

/* Per-instance vote count (voxel_instance_voting.py:169-187, in_hull :62-76):
 * for each of K axis-aligned boxes count local-map points inside (inclusive
 * lo <= p <= hi) with prediction 1 (weight 1) and prediction 2 (weight 2).
 *   points : (P, row_stride>=3) float32 ; pred : (P,) int64
 *   box_lo, box_hi : (K, 3) float32
 *   sums   : (K, 2) int64 ZERO-FILLED by the caller -> [static_sum, dynamic_sum]
 * The cluster label is 2 if dynamic_sum > static_sum else 1 (:184-187). */
int smos_instance_vote(const float* points, int64_t P, int64_t row_stride,
                       const int64_t* pred, const float* box_lo, const float* box_hi,
                       int32_t K, int64_t* sums, void* stream);

/* Same, with the number of boxes read on the device: K = min(*K_dev, K_cap). For boxes produced by
 * smos_cluster_boxes (K_dev = counts + 2), so that clustering, vote and write-back need no host read in between.
 * sums (K_cap, 2) int64 zero-filled by the caller. */
int smos_instance_vote_counted(const float* points, int64_t P, int64_t row_stride,
                               const int64_t* pred, const float* box_lo, const float* box_hi,
                               int32_t K_cap, const int32_t* K_dev, int64_t* sums, void* stream);

/* Same vote without a zero-filled output: the CTAs accumulate into `workspace` (smos_instance_vote_workspace_bytes(K_cap)
 * bytes, 8-byte aligned, ZERO before its first use; every call leaves it zero again, so a stream can own one for its
 * lifetime) and the last CTA to finish writes sums (K_cap, 2) — every element, no pre-fill. K_dev may be NULL (K = K_cap).
 * Calls that share a workspace must be ordered on one stream. */
int64_t smos_instance_vote_workspace_bytes(int32_t K);
int smos_instance_vote_ws(const float* points, int64_t P, int64_t row_stride,
                          const int64_t* pred, const float* box_lo, const float* box_hi,
                          int32_t K_cap, const int32_t* K_dev, void* workspace, int64_t* sums, void* stream);

/* Instance clustering — cluster() of voxel_instance_voting.py:144-175 up to the vote block:
 * foreground = points with pred_bf == 2 (:145), DBSCAN(eps, min_samples) over their xyz (:150-153; scikit-learn's
 * labels, reproduced by an order-free formulation — see csrc/cluster.cu), clusters with more than
 * min_cluster_points points kept in label order (:160-166), one axis-aligned box per kept cluster with the floor
 * lifted by z_lift in float32 (:169-175). The boxes feed smos_instance_vote[_counted]; smos_cluster_apply then writes the
 * voted label to every point of a kept cluster (:184-191). No host synchronisation; nine kernels.
 *   points (n, row_stride>=3) f32 ; pred_bf (n,) int32
 *   workspace : smos_cluster_workspace_bytes(n) bytes, 16-byte aligned, kept until smos_cluster_apply
 *   fg_index  (n,) int32 : indices of the foreground points, ascending (first M entries)
 *   fg_label  (n,) int32 : DBSCAN label of foreground point i (first M entries; -1 = noise)
 *   counts    (3,) int32 : M, number of clusters, number of kept clusters K
 *   box_lo, box_hi (Kcap, 3) f32, kept_label (Kcap,) int32, Kcap >= n / (min_cluster_points + 1) + 1 */
int64_t smos_cluster_workspace_bytes(int64_t n);
int smos_cluster_boxes(const float* points, int64_t n, int64_t row_stride, const int32_t* pred_bf, double eps,
                       int32_t min_samples, int32_t min_cluster_points, float z_lift, void* workspace,
                       int32_t* fg_index, int32_t* fg_label, int32_t* counts, float* box_lo, float* box_hi,
                       int32_t* kept_label, void* stream);
/* sums (K, 2) int64 from smos_instance_vote over the K kept boxes; pred (n,) int64 is updated in place. */
int smos_cluster_apply(int64_t n, void* workspace, const int32_t* fg_index, const int32_t* fg_label,
                       const int32_t* counts, const int64_t* sums, int64_t* pred, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STREAMMOS_B200_H_ */
