/*
 * sis3d.h -- C ABI of libsis3d.so: the B200 (sm_100a) dense-voxel inference hot path of 3D-SIS.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless named h_*;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream, allocate
 *     nothing and never call exit(); caller owns all memory (workspace sizes via *_workspace_bytes);
 *   - return 0 on success, a negative SIS3D_E* code otherwise (see sis3d_strerror);
 *   - "VC layout" = channels-last voxel tensor [X][Y][Z][C] fp32 (C contiguous); "NCDHW" = the
 *     reference's [1][C][X][Y][Z] layout.
 *
 * Each entry cites the reference interface it replaces (paths relative to the reference root).
 */
#ifndef SIS3D_H
#define SIS3D_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SIS3D_OK 0
#define SIS3D_EINVAL (-1)   /* bad argument (shape, alignment, null pointer) */
#define SIS3D_ELAUNCH (-2)  /* CUDA launch / runtime error (cudaGetLastError) */
#define SIS3D_EWORKSPACE (-3)
#define SIS3D_EUNSUPPORTED (-4)

const char *sis3d_strerror(int code);
int sis3d_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t sis3d_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * 3D greedy NMS.  Replaces  int gpu_nms(THLongTensor* keep, THLongTensor* num_out, THCudaTensor*
 * boxes, float thresh)  (lib/layer_utils/nms/src/nms_cuda.h:1, src/nms_cuda.c:10-67) and the kernel
 * ABI  void _nms(int, float*, unsigned long long*, float)  (src/cuda/nms_kernel.h:14).
 * boxes [n,6] fp32 sorted by descending score; keep int64[n] (device), num_out int32[1] (device).
 * IoU arithmetic is bit-identical to the reference CUDA build (inclusive +1 extents, one FFMA).
 * workspace: sis3d_nms_workspace_bytes(n) bytes, 8-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
size_t sis3d_nms_workspace_bytes(int n);
int sis3d_nms(const float *boxes, int n, float thresh, int64_t *keep, int32_t *num_out,
              void *workspace, void *stream);

/* ------------------------------------------------------------------------------------------------
 * 3D RoI max pooling forward.  Replaces  int ROIPoolForwardLaucher(const float* bottom, float
 * scale, int num_rois, int width, int height, int length, int channels, int pw, int ph, int pl,
 * const float* rois, float* top, int* argmax, cudaStream_t)
 * (lib/layer_utils/roi_pooling/src/cuda/roi_pooling_kernel.h:8-12) and roi_pooling_forward_cuda
 * (src/roi_pooling_cuda.h:1-2).  feat_layout: 0 = NCDHW (reference), 1 = VC.
 * rois [n,6]; top [n,C,pw,ph,pl]; argmax int32 same shape or NULL (indices in the NCDHW convention).
 * sis3d_roi_pool_levels is the pyramid form used by the forward (lib/nets/network.py:503-534): VC
 * features of up to three levels, level_ids int32[n] (1-based, 0 = padded row -> zeros) pick the map.
 * ---------------------------------------------------------------------------------------------- */
int sis3d_roi_pool_fwd(const float *feat, int feat_layout, float spatial_scale, int num_rois,
                       int width, int height, int length, int channels, int pw, int ph, int pl,
                       const float *rois, float *top, int32_t *argmax, void *stream);
int sis3d_roi_pool_levels(const float *feat1, const float *feat2, const float *feat3,
                          const int32_t *level_ids, float spatial_scale, int num_rois, int width,
                          int height, int length, int channels, int pw, int ph, int pl,
                          const float *rois, float *top, int32_t *argmax, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Back-projection.  Replaces ProjectionHelper.compute_projection (lib/layer_utils/projection.py:
 * 52-121) and Projection.forward + the running max of lib/nets/network.py:220-239.
 *
 * views: float[n_views][40] per view = world_to_camera(16) | grid_to_world(16) | bounds_min(3) |
 *        bounds_max(3) | pad(2)   (row-major 4x4; bounds already clamped to the volume)
 * depth: float[n_views][img_h][img_w];   fx, fy, cx, cy = pinhole intrinsics of the depth image
 * sis3d_project_map     : pix int16[n_views][Z*Y*X] = v*img_w+u of the pixel a voxel projects to or
 *                         -1; index = z*X*Y + y*X + x (the reference's linear index); counts
 *                         int32[n_views] = number of valid voxels per view (zeroed by the call).
 * sis3d_project_compact : ordered compaction of one view's map into the reference's (lin3d, lin2d)
 *                         int64 lists of length 1+Z*Y*X with element 0 = count.
 * sis3d_backproject_max : fused gather + cross-view max into a VC volume [X][Y][Z][C]; reproduces
 *                         the reference's pairing of feature maps with surviving index lists when
 *                         views are dropped (pairs from sis3d_backproject_pairs), see DESIGN.md.  feats [n_views][C][h][w]
 *                         (reference layout) are first transposed to [n_views][h*w][C] into
 *                         `feats_t` (workspace of the same size).
 * ---------------------------------------------------------------------------------------------- */
/* Host helper (CPU code): packs the per-view constants views[n][40] from the poses, world2grid and their inverses
 * (the caller inverts the 4x4s, e.g. torch.inverse as the reference does); frustum bounds per projection.py:27-60. */
int sis3d_view_params_host(const float *h_poses, const float *h_w2g, const float *h_inv_poses, const float *h_inv_w2g,
                           int n, int n_w2g, double fx, double fy, double cx, double cy, int img_w, int img_h,
                           double depth_min, double depth_max, int X, int Y, int Z, float *h_out);
int sis3d_project_map(const float *views, const float *depth, int n_views, int img_w, int img_h,
                      float fx, float fy, float cx, float cy, float depth_min, float depth_max, float voxel_size,
                      int X, int Y, int Z, int16_t *pix, int32_t *counts, void *stream);
int sis3d_project_compact(const int16_t *pix, int X, int Y, int Z, int64_t *lin3d, int64_t *lin2d,
                          void *workspace, size_t workspace_bytes, void *stream);
size_t sis3d_project_compact_workspace_bytes(int X, int Y, int Z);
/* pairs int32[3*n_views] (2*n_pairs used), n_pairs int32[1]: (feature map, index map) pairing derived
 * on the device from counts -- no host round trip. */
int sis3d_backproject_pairs(const int32_t *counts, int n_views, int32_t *pairs, int32_t *n_pairs, void *stream);
/* reference-style index lists [n_lists][1+Z*Y*X] (lib/layer_utils/projection.py:110-121) -> dense maps */
int sis3d_project_scatter_lists(const int64_t *lin3d, const int64_t *lin2d, int n_lists, int X, int Y, int Z,
                                int16_t *pix, void *stream);
int sis3d_backproject_max(const float *feats, float *feats_t, const int16_t *pix,
                          const int32_t *pairs, const int32_t *n_pairs, int n_views, int C,
                          int img_w, int img_h, int X, int Y, int Z, float *volume_vc, void *stream);

/* Back-projection fused with the first colour convolution (2x2x2, stride 2, no bias, ReLU;
 * lib/nets/backbones.py:203 / :137): out = relu(conv3d(max-over-views back-projection, W, stride 2))
 * WITHOUT materialising the (>= 95 % zero) feature volume: covered voxels are listed per tap, multiplied by
 * that tap's weight slice and combined per output voxel in tap order (csrc/sparse.cu).  Arguments as for
 * sis3d_backproject_max; w_packed from sis3d_pack_conv_weight(ks=2) ([8*C][cout]); out is VC
 * [X/2][Y/2][Z/2][cout] with row stride out_ld / channel offset out_coff.  C % 16 == 0, cout in {32,64}. */
size_t sis3d_backproject_conv_k2s2_workspace_bytes(int X, int Y, int Z, int cout);
int sis3d_backproject_conv_k2s2(const float *feats, float *feats_t, const int16_t *pix, const int32_t *pairs,
                                const int32_t *n_pairs, int n_views, int C, int img_w, int img_h, int X,
                                int Y, int Z, const float *w_packed, int cout, float *out, int out_ld,
                                int out_coff, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * 3D convolution (implicit GEMM) over a table of regions.  Replaces the nn.Conv3d call sites of
 * lib/nets/backbones.py:20-22,126-169,188-231,241-272 and lib/nets/network.py:40-52 (cuDNN in the
 * reference).  One launch processes `n_regions` independent volumes (1 for the backbone, one per
 * RoI crop for the mask head, lib/nets/network.py:303-317).
 * Weights are pre-packed by sis3d_pack_conv_weight into [ks^3*cin][cout] (k = tap*cin + c,
 * tap = (kx*ks+ky)*ks+kz).  Output is VC with row stride out_ld and channel offset out_coff (lets a
 * producer write straight into a concatenated tensor); optional bias[cout], residual (VC, res_ld /
 * res_coff), act: 0 none, 1 relu, 2 sigmoid.
 * ---------------------------------------------------------------------------------------------- */
typedef struct sis3d_region {
    int64_t in_off;   /* element offset of the region's input voxel (0,0,0) */
    int64_t out_off;  /* element offset of the region's output voxel (0,0,0), before out_coff */
    int64_t res_off;
    int32_t in_dim[3];     /* input extent (x,y,z): taps outside read as zero */
    int32_t out_dim[3];    /* output extent */
    int64_t in_stride[3];  /* input voxel strides in elements (x,y,z) */
    int64_t out_stride[3]; /* output voxel strides in elements; all zero = dense rows (m * out_ld) */
    int32_t tile_begin;    /* first M-tile of this region (exclusive prefix sum of tile counts) */
    int32_t pad_;
} sis3d_region;

#define SIS3D_CONV_TILE_M 64
int sis3d_pack_conv_weight(const float *w_oidhw, int cout, int cin, int ks, float *w_packed, void *stream);
int sis3d_conv3d(const float *in, int64_t in_chan_stride, const float *w_packed, const float *bias,
                 const float *residual, int res_ld, int res_coff, float *out, int out_ld, int out_coff,
                 const sis3d_region *regions, int n_regions, int n_tiles, int cin, int cout, int ks,
                 int stride, int pad, int act, void *stream);
/* Fully connected layer y[M][N] = act(x[M][K] . W + b) for the RoI classifier MLP (lib/nets/backbones.py:
 * 92-96,225-231): split-K over `splits` CTAs-per-tile with a deterministic second-pass reduction.
 * w_packed from sis3d_pack_conv_weight(ks=1).  workspace >= sis3d_linear_workspace_bytes(M,N,K). */
size_t sis3d_linear_workspace_bytes(int M, int N, int K);
int sis3d_linear(const float *x, const float *w_packed, const float *bias, float *y, int M, int K, int N,
                 int act, void *workspace, size_t workspace_bytes, void *stream);

/* Same layer on the tensor cores (tcgen05 TF32, fp32 accumulate, split-K + deterministic reduce): w_nk is the
 * nn.Linear weight as stored, [N][K].  K % 32 == 0, N in {32, 64, 128k}. */
int sis3d_linear_tc_supported(int K, int N);
size_t sis3d_linear_tc_workspace_bytes(int M, int N, int K);
int sis3d_linear_tc(const float *x, const float *w_nk, const float *bias, float *y, int M, int K, int N, int act,
                    void *workspace, size_t workspace_bytes, void *stream);

/* Layers 2-3 of the classifier MLP and both heads in one launch (activations stay in shared memory):
 * x1 [R][d1] -> relu(W2) [d2] -> relu(W3) [d3] -> cls_score [R][nc] and bbox_pred [R][nb].  All weights packed by
 * sis3d_pack_conv_weight(ks=1); d1,d2,d3 <= 256.  (lib/nets/backbones.py:225-231, lib/nets/network.py:55-57) */
int sis3d_mlp_tail(const float *x1, int R, int d1, const float *w2, const float *b2, int d2, const float *w3,
                   const float *b3, int d3, const float *wc, const float *bc, int nc, const float *wb,
                   const float *bb, int nb, float *cls_score, float *bbox_pred, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core path: ks = 3 (stride 1, pad 1), ks = 1, or ks = 2 (the 2x2x2 / stride-2 / pad-0 downsampling convs,
 * lib/nets/backbones.py:182-205; X,Y,Z are then the INPUT extents, the output is [X/2][Y/2][Z/2][cout], tiles must be NULL;
 * the A operand is fetched by TMA boxes with element strides {1,2,2,2}) -- same call sites as sis3d_conv3d:
 * tcgen05.mma kind::tf32 with fp32 accumulation in TMEM, operands staged by 4-D/2-D TMA boxes
 * (csrc/conv_tc.cu).  `in` is a dense VC tensor [X][Y][Z][cin]; w_tc comes from
 * sis3d_pack_conv_weight_tc ([cout][ks^3*cin]).  tiles == NULL covers the whole volume with 8x2x8 (x,y,z)
 * bricks; otherwise (brick shape from sis3d_conv3d_tc_brick) tiles int32[n_tiles][8] = {x0,y0,z0,x1,y1,z1,0,0} lists brick origins and the
 * exclusive end of the voxels to be written (ragged RoI crops packed on one canvas).
 * Requires cin % 32 == 0 and cout in {32, 64, 128k}; returns SIS3D_EUNSUPPORTED otherwise.
 * ---------------------------------------------------------------------------------------------- */
int sis3d_pack_conv_weight_tc(const float *w_oidhw, int cout, int cin, int ks, float *w_tc, void *stream);
int sis3d_conv3d_k3_tc_supported(int cin, int cout);
/* brick shape (x,y,z) the kernel uses for a layer: 8x2x8 for whole volumes, 4x4x8 for explicit tile lists of 64->64 layers */
void sis3d_conv3d_tc_brick(int with_tile_list, int ks, int cin, int cout, int32_t *bx, int32_t *by, int32_t *bz);
int sis3d_conv3d_k3_tc(const float *in, const float *w_tc, const float *bias, const float *residual,
                       int res_ld, int res_coff, float *out, int out_ld, int out_coff, int X, int Y,
                       int Z, int cin, int cout, int ks, const int32_t *tiles, int n_tiles, int act, void *stream);

/* Bottleneck tail in ONE kernel (lib/nets/backbones.py:28-40: conv2 3x3x3 (+bias2) -> ReLU -> conv3 1x1 (+bias3) -> + residual -> act;
 * biases may be NULL):
 * the cmid-wide intermediate stays in shared memory / TMEM.  w2_tc / w3_tc from sis3d_pack_conv_weight_tc (ks 3 / ks 1).
 * Supported: cin % 32 == 0 and (cmid, cout) in {(32,32), (32,64), (64,128)}; results equal the two-kernel TF32 path
 * bit for bit (same operand values, same accumulation order). */
int sis3d_conv3d_k3_tc_fused_supported(int cin, int cmid, int cout);
int sis3d_conv3d_k3_tc_fused(const float *in, const float *w2_tc, const float *bias2, const float *w3_tc, const float *bias3,
                             const float *residual, int res_ld, int res_coff, float *out, int out_ld, int out_coff, int X,
                             int Y, int Z, int cin, int cmid, int cout, int act, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Voxel block of a `.scene` / `.chunk` file -> network input, on the device (replaces the host numpy passes of
 * lib/datasets/dataset.py:50-68 + the height crop :192-205).  sdf_xfast: the raw float32 block as stored in the file
 * (index = x + X*(y + Y*z)); data: [2][X][min(Y, y_keep)][Z] (z fastest): channel 0 = |clip(sdf, -truncation, truncation)|,
 * channel 1 = (sdf > -1) as 0/1.  Exact (no rounding involved).
 * ---------------------------------------------------------------------------------------------- */
int sis3d_chunk_decode(const float *sdf_xfast, int X, int Y, int Z, int y_keep, float truncation, float *data, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Error-compensated 3xTF32 on the tensor cores ("x3"): same call sites and arguments as sis3d_conv3d_k3_tc /
 * _fused / sis3d_linear_tc, fp32-class accuracy.  Every fp32 operand is split v = hi + lo (hi = tf32(v), round to
 * nearest; lo = tf32(v - hi)) and the product accumulated as A_lo.B_hi + A_hi.B_lo + A_hi.B_hi in the fp32 TMEM
 * accumulator.  Weights are split once (sis3d_pack_conv_weight_tc_x3 -> [2][cout][ks^3*cin]: hi rows, then lo rows; a
 * [N][K] nn.Linear weight is packed as a ks = 1 conv); activations are split in shared memory after the TMA load, so
 * they cross L2->SM once.  This is the mode in which the detector's integer outputs (top-N order, NMS keep lists, class
 * argmax, crop bounds) match the fp32 reference (lib/layer_utils/proposal_layer.py:181-197, lib/nets/network.py:296-301)
 * while the convolutions still run on tcgen05.  Whole volumes only (tiles = NULL).
 * ---------------------------------------------------------------------------------------------- */
int sis3d_pack_conv_weight_tc_x3(const float *w_oidhw, int cout, int cin, int ks, float *w_x3, void *stream);
int sis3d_conv3d_k3_tc_x3(const float *in, const float *w_x3, const float *bias, const float *residual,
                          int res_ld, int res_coff, float *out, int out_ld, int out_coff, int X, int Y,
                          int Z, int cin, int cout, int ks, int act, void *stream);
int sis3d_conv3d_k3_tc_fused_x3(const float *in, const float *w2_x3, const float *bias2, const float *w3_x3,
                                const float *bias3, const float *residual, int res_ld, int res_coff, float *out,
                                int out_ld, int out_coff, int X, int Y, int Z, int cin, int cmid, int cout, int act,
                                void *stream);
int sis3d_linear_tc_x3(const float *x, const float *w_x3, const float *bias, float *y, int M, int K, int N, int act,
                       void *workspace, size_t workspace_bytes, void *stream);
/* The same compensation with an fp16 operand split ("h3"): v = hi + lo / 2048, hi = fp16(v), lo = fp16((v - hi) * 2048) -- the
 * same 11 + 11 significand bits as the TF32 split -- so the three products are tcgen05 kind::f16 MMAs (2.5x the TF32 rate
 * measured on B200) over 32-channel stages.  Activations stay fp32 in HBM (split in shared memory after the TMA load, |v|
 * range-limited to fp16: |v| < 65504); weights from sis3d_pack_conv_weight_tc_h3: fp16 [cout][ks^3*cin/32][64], each 128-byte
 * row = the 32 hi halves of a 32-channel K slice followed by its 32 lo halves (2*cout*ks^3*cin halves in total), so one TMA
 * row request carries both parts.  Same arguments as _x3. */
int sis3d_pack_conv_weight_tc_h3(const float *w_oidhw, int cout, int cin, int ks, uint16_t *w_h3, void *stream);
int sis3d_conv3d_k3_tc_h3(const float *in, const uint16_t *w_h3, const float *bias, const float *residual,
                          int res_ld, int res_coff, float *out, int out_ld, int out_coff, int X, int Y,
                          int Z, int cin, int cout, int ks, int act, void *stream);
int sis3d_conv3d_k3_tc_fused_h3(const float *in, const uint16_t *w2_h3, const float *bias2, const uint16_t *w3_h3,
                                const float *bias3, const float *residual, int res_ld, int res_coff, float *out,
                                int out_ld, int out_coff, int X, int Y, int Z, int cin, int cmid, int cout, int act,
                                void *stream);
int sis3d_linear_tc_h3(const float *x, const uint16_t *w_h3, const float *bias, float *y, int M, int K, int N, int act,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * fp16-operand tensor-core path (tcgen05 kind::f16, fp32 accumulate): activations and weights are STORED as fp16 (same
 * 11-bit significand TF32 keeps, half the bytes through L2 -- the kernel is bound by operand feed).  in16 is a dense VC
 * fp16 tensor; w16 from sis3d_pack_conv_weight_tc_f16; outputs: out32 (fp32, may be NULL) and/or out16 (fp16 twin with
 * the same row stride / channel offset, may be NULL).  cin % 64 == 0 uses 128-B rows, cin == 32 64-B rows.
 * sis3d_cast_f16 makes the fp16 copy of an fp32 VC tensor; sis3d_conv3d_ex / sis3d_backproject_conv_k2s2_ex are the
 * fp32 kernels with an optional fp16 twin output.
 * ---------------------------------------------------------------------------------------------- */
int sis3d_pack_conv_weight_tc_f16(const float *w_oidhw, int cout, int cin, int ks, uint16_t *w16, void *stream);
int sis3d_cast_f16(const float *in, int in_ld, int in_coff, int64_t rows, int C, uint16_t *out, void *stream);
int sis3d_conv3d_tc_f16(const uint16_t *in16, const uint16_t *w16, const float *bias, const float *residual,
                        int res_ld, int res_coff, float *out32, uint16_t *out16, int out_ld, int out_coff, int X,
                        int Y, int Z, int cin, int cout, int ks, const int32_t *tiles, int n_tiles, int act, void *stream);
int sis3d_conv3d_ex(const float *in, int64_t in_chan_stride, const float *w_packed, const float *bias,
                    const float *residual, int res_ld, int res_coff, float *out, uint16_t *out16, int out_ld,
                    int out_coff, const sis3d_region *regions, int n_regions, int n_tiles, int cin, int cout, int ks,
                    int stride, int pad, int act, void *stream);
int sis3d_backproject_conv_k2s2_ex(const float *feats, float *feats_t, const int16_t *pix, const int32_t *pairs,
                                   const int32_t *n_pairs, int n_views, int C, int img_w, int img_h, int X, int Y,
                                   int Z, const float *w_packed, int cout, float *out, uint16_t *out16, int out_ld,
                                   int out_coff, void *workspace, size_t workspace_bytes, void *stream);

/* MaxPool3d(3,1,1) on a VC tensor (lib/nets/backbones.py:207,212,220); output row stride out_ld and
 * channel offset out_coff as for the convolution. */
int sis3d_maxpool3(const float *in, float *out, int out_ld, int out_coff, int X, int Y, int Z, int C, void *stream);
/* VC [X][Y][Z][C] -> NCDHW [C][X][Y][Z] (to hand tensors back in the reference layout). */
int sis3d_vc_to_ncdhw(const float *in, float *out, int64_t nvox, int C, void *stream);

/* ------------------------------------------------------------------------------------------------
 * RPN proposals ("rpn3d").  Replaces proposal_layer (lib/layer_utils/proposal_layer.py:11-204),
 * the softmax of lib/nets/network.py:546, generate_anchors (lib/layer_utils/generate_anchors.py:
 * 58-119), bbox_transform_inv / clip_boxes (lib/utils/bbox_transform.py:59-99,4-21) and the nms call
 * (proposal_layer.py:190) in three launches with no host round trip.
 * Per level l: cls [N_l][2*A_l] VC logits (channel = cls*A+a), deltas [N_l][6*A_l] VC,
 * anchor_sizes float[A_l][3], grid (gx,gy,gz), N_l = gx*gy*gz.
 * Outputs: rois [post_top_n][6], scores [post_top_n], level_ids int32[post_top_n] (1-based),
 * num_out int32[1]; debug_order int32[pre_top_n] (flat anchor index of each pre-NMS box, or NULL).
 * Tie rule: equal scores are ordered by ascending flat index (level, x, y, z, a).
 * ---------------------------------------------------------------------------------------------- */
typedef struct sis3d_rpn_level {
    const float *cls;
    const float *deltas;
    const float *anchor_sizes;
    int32_t grid[3];
    int32_t num_anchors;
    int32_t cls_mode; /* 0: cls = logits [N][2A]; 1: cls = foreground probability [N][A] */
    int32_t cls_ld;    /* row stride of cls in elements (0 = dense) -- lets both heads share one conv output */
    int32_t deltas_ld; /* row stride of deltas in elements (0 = dense 6A) */
    int32_t pad_;
} sis3d_rpn_level;
size_t sis3d_rpn_workspace_bytes(const sis3d_rpn_level *h_levels, int n_levels, int pre_top_n);
int sis3d_rpn_proposals(const sis3d_rpn_level *h_levels, int n_levels, int feat_stride,
                        int scene_x, int scene_y, int scene_z, int allow_border, int pre_top_n,
                        int post_top_n, float nms_thresh, float *rois, float *scores,
                        int32_t *level_ids, int32_t *num_out, int32_t *debug_order,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Detection decode for the mask branch.  Replaces lib/nets/network.py:285-301 (== lib/model/
 * trainval.py:825-858): argmax class, softmax confidence, class-specific bbox_transform_inv,
 * clip, conf > class_thresh and round()-degenerate filter (Python-3 half-to-even).
 * det float[n][16] = pred_box(6) | conf | cls | keep | crop lo(3) hi(3) | pad   (all as float)
 * n is read from device num_rois (int32[1]) and clamped to max_rois.
 * ---------------------------------------------------------------------------------------------- */
int sis3d_detect_decode(const float *rois, const int32_t *num_rois, int max_rois,
                        const float *cls_score, const float *bbox_pred, int num_classes,
                        int scene_x, int scene_y, int scene_z, float class_thresh,
                        float *cls_prob, int64_t *cls_pred, float *det, void *stream);

/* Host-side planner of the ragged mask stage (CPU code, no CUDA calls): from the decoded detection table
 * h_det [n][16] (sis3d_detect_decode layout) builds every table the six mask-head launches need into one host blob
 * (ship it with a single pinned H2D copy).  use_canvas = 1: crops packed along x on a zero canvas of size plan.canvas
 * (tensor-core path, brick list at off_rest); 0: compact per-crop buffers (region table of the middle layers at
 * off_rest).  Returns SIS3D_EWORKSPACE with plan->bytes set when `capacity` is too small. */
typedef struct sis3d_mask_plan {
    int32_t n_kept, canvas[3], n_tiles_tc, tiles_first, tiles_last, tiles_mid;
    int64_t total_voxels;
    int64_t off_first, off_last, off_rest, off_offs, off_cls, off_kept, off_sizes, bytes;
} sis3d_mask_plan;
int sis3d_mask_plan_build(const float *h_det, int n, int X, int Y, int Z, int ncls, int use_canvas, void *h_blob,
                          size_t capacity, sis3d_mask_plan *plan);

/* One host call for every launch of a scene's ragged mask stage: H2D of the planner's tables, canvas zeroing, the six
 * layers (math 0: fp32 CUDA-core kernel on compact crops; 1: TF32 tcgen05 on the canvas; 2: fp16-operand tcgen05), the
 * predicted-class select and (bits_host != NULL) the D2H of the thresholded masks.  w_first / w_last: sis3d_pack_conv_weight
 * of geometry.0 / geometry.10; w_mid[i]: geometry.{2,4,6,8} packed for `math` (pack_conv_weight / _tc / _tc_f16).
 * canvas_bytes >= 2 layers x voxels x 64 channels in the operand type; SIS3D_EWORKSPACE otherwise. */
typedef struct sis3d_mask_stage {
    const float *scene;            /* NCDHW [2][X][Y][Z] */
    int32_t X, Y, Z, ncls, math, reserved;
    const float *w_first, *w_last;
    const void *w_mid[4];
    void *tables;                  /* device, >= plan->bytes */
    void *canvas;
    size_t canvas_bytes;
    float *canvas32;               /* math 2 only: fp32 output of the last 3x3x3 layer */
    size_t canvas32_bytes;
    float *masks;                  /* [total_voxels][ncls] sigmoid outputs */
    uint8_t *bits;                 /* [total_voxels] thresholded predicted-class masks, or NULL */
    uint8_t *bits_host;            /* pinned host destination for `bits`, or NULL */
    float thresh;
    int32_t reserved2;
} sis3d_mask_stage;
int sis3d_mask_stage_launch(const sis3d_mask_plan *plan, const void *h_blob, const sis3d_mask_stage *args, void *stream);
/* cudaMemcpyAsync behind the C ABI (kind 1 H2D, 2 D2H, 3 D2D) for hosts that stage their own pinned buffers. */
int sis3d_memcpy_async(void *dst, const void *src, size_t bytes, int kind, void *stream);

/* Predicted-class mask channel of every kept RoI, packed back to back, optionally thresholded to bits
 * (lib/model/trainval.py:900-908).  masks [total][ncls]; offs int64[n_crops+1] voxel offsets; cls int32[n_crops]. */
int sis3d_mask_select(const float *masks, const int64_t *offs, const int32_t *cls, int n_crops, int ncls,
                      int64_t total_hint, float thresh, float *out, uint8_t *bits, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SIS3D_H */
