// roi.cu -- 3D RoI max pooling (forward) and the detection decode feeding the mask branch.
//
//   roi_pool_ncdhw_kernel  drop-in for ROIPoolForward (lib/layer_utils/roi_pooling/src/cuda/
//                          roi_pooling_kernel.cu:15-109) on the reference's NCDHW features
//   roi_pool_vc_kernel     same arithmetic on VC features, one CTA per RoI, channel-coalesced reads,
//                          picks the pyramid level per RoI (lib/nets/network.py:503-534)
//   detect_decode_kernel   lib/nets/network.py:285-301 / lib/model/trainval.py:825-858
#include <float.h>
#include "common.cuh"

namespace sis3d {

struct RoiBins {
    int sw, sh, sl;
    float bw, bh, bl;
};
__device__ __forceinline__ RoiBins roi_bins(const float *roi, float scale, int pw, int ph, int pl) {
    RoiBins b;
    b.sw = (int)floorf(__fmul_rn(roi[0], scale));
    b.sh = (int)floorf(__fmul_rn(roi[1], scale));
    b.sl = (int)floorf(__fmul_rn(roi[2], scale));
    const int ew = (int)ceilf(__fmul_rn(roi[3], scale)), eh = (int)ceilf(__fmul_rn(roi[4], scale));
    const int el = (int)ceilf(__fmul_rn(roi[5], scale));
    const int rw = max(ew - b.sw, 1), rh = max(eh - b.sh, 1), rl = max(el - b.sl, 1);  // malformed RoIs -> 1x1x1
    b.bw = __fdiv_rn((float)rw, (float)pw);
    b.bh = __fdiv_rn((float)rh, (float)ph);
    b.bl = __fdiv_rn((float)rl, (float)pl);
    return b;
}
__device__ __forceinline__ void bin_range(int p, float bin, int start, int dim, int &lo, int &hi) {
    lo = (int)floorf(__fmul_rn((float)p, bin));
    hi = (int)ceilf(__fmul_rn((float)(p + 1), bin));
    lo = min(max(lo + start, 0), dim);
    hi = min(max(hi + start, 0), dim);
}

__global__ void __launch_bounds__(256) roi_pool_ncdhw_kernel(int nthreads, const float *feat, float scale, int W, int H, int L,
                                                             int C, int pw_, int ph_, int pl_, const float *rois, float *top,
                                                             int *argmax) {
    for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < nthreads; index += blockDim.x * gridDim.x) {
        int n = index;
        const int pl = n % pl_; n /= pl_;
        const int ph = n % ph_; n /= ph_;
        const int pw = n % pw_; n /= pw_;
        const int c = n % C; n /= C;
        const RoiBins b = roi_bins(rois + n * 6, scale, pw_, ph_, pl_);
        int ws, we, hs, he, ls, le;
        bin_range(pw, b.bw, b.sw, W, ws, we);
        bin_range(ph, b.bh, b.sh, H, hs, he);
        bin_range(pl, b.bl, b.sl, L, ls, le);
        const bool empty = (he <= hs) || (we <= ws) || (le <= ls);
        float best = empty ? 0.f : -FLT_MAX;
        int besti = -1;
        for (int w = ws; w < we; ++w)
            for (int h = hs; h < he; ++h)
                for (int l = ls; l < le; ++l) {
                    const int idx = (c * W + w) * H * L + h * L + l;
                    const float v = __ldg(feat + idx);
                    if (v > best) { best = v; besti = idx; }
                }
        top[index] = best;
        if (argmax) argmax[index] = besti;
    }
}

constexpr int kRoiMaxBins = 64;
// grid = RoIs, block = 128 threads (channels); smem staging so the [C][bins] rows are written coalesced
__global__ void __launch_bounds__(128) roi_pool_vc_kernel(const float *feat1, const float *feat2, const float *feat3,
                                                          const int32_t *level_ids, float scale, int W, int H, int L, int C,
                                                          int pw_, int ph_, int pl_, const float *rois, float *top, int *argmax) {
    extern __shared__ float s_out[];  // [blockDim.x][nb+1] values, then ints for argmax
    const int r = blockIdx.x;
    const int nb = pw_ * ph_ * pl_;
    int *s_arg = reinterpret_cast<int *>(s_out + blockDim.x * (nb + 1));
    const int lvl = level_ids ? level_ids[r] : 1;
    const float *feat = lvl == 1 ? feat1 : (lvl == 2 ? feat2 : (lvl == 3 ? feat3 : nullptr));
    const RoiBins b = roi_bins(rois + r * 6, scale, pw_, ph_, pl_);
    for (int c0 = 0; c0 < C; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        if (c < C) {
            int bin = 0;
            for (int pw = 0; pw < pw_; ++pw) {
                int ws, we;
                bin_range(pw, b.bw, b.sw, W, ws, we);
                for (int ph = 0; ph < ph_; ++ph) {
                    int hs, he;
                    bin_range(ph, b.bh, b.sh, H, hs, he);
                    for (int pl = 0; pl < pl_; ++pl, ++bin) {
                        int ls, le;
                        bin_range(pl, b.bl, b.sl, L, ls, le);
                        const bool empty = (he <= hs) || (we <= ws) || (le <= ls) || !feat;
                        float best = empty ? 0.f : -FLT_MAX;
                        int besti = -1;
                        if (!empty)
                            for (int w = ws; w < we; ++w)
                                for (int h = hs; h < he; ++h) {
                                    const float *row = feat + ((int64_t)(w * H + h) * L) * C + c;
                                    for (int l = ls; l < le; ++l) {
                                        const float v = __ldg(row + (int64_t)l * C);
                                        if (v > best) { best = v; besti = (c * W + w) * H * L + h * L + l; }
                                    }
                                }
                        s_out[threadIdx.x * (nb + 1) + bin] = best;
                        s_arg[threadIdx.x * (nb + 1) + bin] = besti;
                    }
                }
            }
        }
        __syncthreads();
        const int cn = min((int)blockDim.x, C - c0);
        for (int i = threadIdx.x; i < cn * nb; i += blockDim.x) {
            const int cc = i / nb, bb = i - cc * nb;
            const int64_t o = ((int64_t)r * C + c0 + cc) * nb + bb;
            top[o] = s_out[cc * (nb + 1) + bb];
            if (argmax) argmax[o] = s_arg[cc * (nb + 1) + bb];
        }
        __syncthreads();
    }
}

// Tail of the RoI classifier (lib/nets/backbones.py:225-231 layers 2,4 + lib/nets/network.py:55-57 heads) in one
// launch: 8 RoI rows per CTA, activations stay in shared memory, weights ([K][ldw] packed) stream from L2.
constexpr int kMlpRows = 8;
__global__ void __launch_bounds__(256) mlp_tail_kernel(const float *x1, int R, int d1, const float *w2, const float *b2, int d2,
                                                       const float *w3, const float *b3, int d3, const float *wc,
                                                       const float *bc, int nc, const float *wb, const float *bb, int nb,
                                                       float *cls_score, float *bbox_pred) {
    __shared__ float sa[kMlpRows][256];
    __shared__ float sb[kMlpRows][256];
    const int r0 = blockIdx.x * kMlpRows, t = threadIdx.x;
    for (int i = t; i < kMlpRows * d1; i += 256) {
        const int r = i / d1, k = i - r * d1;
        sa[r][k] = (r0 + r < R) ? x1[(int64_t)(r0 + r) * d1 + k] : 0.f;
    }
    __syncthreads();
    float acc[kMlpRows];
    if (t < d2) {  // layer 2: d1 -> d2, ReLU
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) acc[r] = 0.f;
        const int ld = (d2 + 3) & ~3;
        for (int k = 0; k < d1; ++k) {
            const float w = __ldg(w2 + (int64_t)k * ld + t);
#pragma unroll
            for (int r = 0; r < kMlpRows; ++r) acc[r] = fmaf(sa[r][k], w, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) sb[r][t] = fmaxf(acc[r] + b2[t], 0.f);
    }
    __syncthreads();
    if (t < d3) {  // layer 3: d2 -> d3, ReLU
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) acc[r] = 0.f;
        const int ld = (d3 + 3) & ~3;
        for (int k = 0; k < d2; ++k) {
            const float w = __ldg(w3 + (int64_t)k * ld + t);
#pragma unroll
            for (int r = 0; r < kMlpRows; ++r) acc[r] = fmaf(sb[r][k], w, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) sa[r][t] = fmaxf(acc[r] + b3[t], 0.f);
    }
    __syncthreads();
    for (int o = t; o < nc + nb; o += 256) {  // heads: d3 -> nc (class scores) | nb (box deltas)
        const bool is_cls = o < nc;
        const int n = is_cls ? o : o - nc;
        const float *w = is_cls ? wc : wb;
        const int ld = ((is_cls ? nc : nb) + 3) & ~3;
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) acc[r] = 0.f;
        for (int k = 0; k < d3; ++k) {
            const float wv = __ldg(w + (int64_t)k * ld + n);
#pragma unroll
            for (int r = 0; r < kMlpRows; ++r) acc[r] = fmaf(sa[r][k], wv, acc[r]);
        }
        const float bias = is_cls ? bc[n] : bb[n];
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r)
            if (r0 + r < R) (is_cls ? cls_score : bbox_pred)[(int64_t)(r0 + r) * (is_cls ? nc : nb) + n] = acc[r] + bias;
    }
}

// one thread per RoI row
__global__ void detect_decode_kernel(const float *rois, const int32_t *num_rois, int max_rois, const float *cls_score,
                                     const float *bbox_pred, int nc, int sx, int sy, int sz, float thresh, float *cls_prob,
                                     int64_t *cls_pred, float *det) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= max_rois) return;
    const int n = min(*num_rois, max_rois);
    float *d = det + i * 16;
    if (i >= n) {
        for (int k = 0; k < 16; ++k) d[k] = 0.f;
        cls_pred[i] = 0;
        for (int k = 0; k < nc; ++k) cls_prob[i * nc + k] = 0.f;
        return;
    }
    const float *s = cls_score + i * nc;
    int best = 0;
    float m = s[0];
    for (int k = 1; k < nc; ++k)
        if (s[k] > m) { m = s[k]; best = k; }  // first maximum (torch.max on CPU)
    float sum = 0.f;
    for (int k = 0; k < nc; ++k) sum += expf(s[k] - m);
    for (int k = 0; k < nc; ++k) cls_prob[i * nc + k] = __fdiv_rn(expf(s[k] - m), sum);
    const float conf = cls_prob[i * nc + best];
    cls_pred[i] = best;
    const float *dl = bbox_pred + (int64_t)i * nc * 6 + best * 6;
    const float dims[3] = {(float)sx, (float)sy, (float)sz};
    float box[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float lo = rois[i * 6 + k], hi = rois[i * 6 + 3 + k];
        const float w = __fsub_rn(hi, lo);
        const float ctr = __fadd_rn(lo, __fmul_rn(0.5f, w));
        const float pc = __fadd_rn(__fmul_rn(dl[k], w), ctr);
        const float pw = __fmul_rn(expf(dl[3 + k]), w);
        const float hw = __fmul_rn(0.5f, pw);
        box[k] = fminf(fmaxf(__fsub_rn(pc, hw), 0.f), dims[k]);
        box[3 + k] = fminf(fmaxf(__fadd_rn(pc, hw), 0.f), dims[k]);
    }
    bool keep = conf > thresh;
    float crop[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) crop[k] = rintf(box[k]);  // Python-3 round(): half to even
    if (crop[0] >= crop[3] || crop[1] >= crop[4] || crop[2] >= crop[5]) keep = false;
    for (int k = 0; k < 6; ++k) d[k] = box[k];
    d[6] = conf; d[7] = (float)best; d[8] = keep ? 1.f : 0.f;
    for (int k = 0; k < 6; ++k) d[9 + k] = crop[k];
    d[15] = 0.f;
}

}  // namespace sis3d
using namespace sis3d;

extern "C" int sis3d_roi_pool_fwd(const float *feat, int feat_layout, float spatial_scale, int num_rois, int width,
                                  int height, int length, int channels, int pw, int ph, int pl, const float *rois,
                                  float *top, int32_t *argmax, void *stream) {
    if (!feat || !rois || !top || num_rois < 0 || pw <= 0 || ph <= 0 || pl <= 0 || channels <= 0) return SIS3D_EINVAL;
    if (num_rois == 0) return SIS3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (feat_layout == 0) {
        const int total = num_rois * channels * pw * ph * pl;
        roi_pool_ncdhw_kernel<<<min(cdiv(total, 256), kNumSMs * 8), 256, 0, s>>>(total, feat, spatial_scale, width, height,
                                                                                length, channels, pw, ph, pl, rois, top,
                                                                                argmax);
    } else {
        const int nb = pw * ph * pl;
        const size_t smem = (size_t)128 * (nb + 1) * 8;
        if (smem > 200 * 1024) return SIS3D_EUNSUPPORTED;
        if (smem > 48 * 1024) cudaFuncSetAttribute(roi_pool_vc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        roi_pool_vc_kernel<<<num_rois, 128, smem, s>>>(feat, nullptr, nullptr, nullptr, spatial_scale, width, height, length,
                                                     channels, pw, ph, pl, rois, top, argmax);
    }
    return finish_launch();
}

extern "C" int sis3d_roi_pool_levels(const float *feat1, const float *feat2, const float *feat3, const int32_t *level_ids,
                                     float spatial_scale, int num_rois, int width, int height, int length, int channels,
                                     int pw, int ph, int pl, const float *rois, float *top, int32_t *argmax, void *stream) {
    if (!feat1 || !level_ids || !rois || !top || num_rois <= 0) return SIS3D_EINVAL;
    const int nb = pw * ph * pl;
    const size_t smem = (size_t)128 * (nb + 1) * 8;
    if (smem > 200 * 1024) return SIS3D_EUNSUPPORTED;
    if (smem > 48 * 1024) cudaFuncSetAttribute(roi_pool_vc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    roi_pool_vc_kernel<<<num_rois, 128, smem, (cudaStream_t)stream>>>(feat1, feat2, feat3, level_ids, spatial_scale, width,
                                                                    height, length, channels, pw, ph, pl, rois, top, argmax);
    return finish_launch();
}

extern "C" int sis3d_detect_decode(const float *rois, const int32_t *num_rois, int max_rois, const float *cls_score,
                                   const float *bbox_pred, int num_classes, int scene_x, int scene_y, int scene_z,
                                   float class_thresh, float *cls_prob, int64_t *cls_pred, float *det, void *stream) {
    if (!rois || !num_rois || !cls_score || !bbox_pred || !cls_prob || !cls_pred || !det || max_rois <= 0) return SIS3D_EINVAL;
    detect_decode_kernel<<<cdiv(max_rois, 128), 128, 0, (cudaStream_t)stream>>>(rois, num_rois, max_rois, cls_score, bbox_pred,
                                                                              num_classes, scene_x, scene_y, scene_z,
                                                                              class_thresh, cls_prob, cls_pred, det);
    return finish_launch();
}

extern "C" int sis3d_mlp_tail(const float *x1, int R, int d1, const float *w2, const float *b2, int d2, const float *w3,
                              const float *b3, int d3, const float *wc, const float *bc, int nc, const float *wb,
                              const float *bb, int nb, float *cls_score, float *bbox_pred, void *stream) {
    if (!x1 || !w2 || !b2 || !w3 || !b3 || !wc || !bc || !wb || !bb || !cls_score || !bbox_pred || R <= 0) return SIS3D_EINVAL;
    if (d1 > 256 || d2 > 256 || d3 > 256 || d1 <= 0 || d2 <= 0 || d3 <= 0 || nc <= 0 || nb <= 0) return SIS3D_EUNSUPPORTED;
    mlp_tail_kernel<<<cdiv(R, kMlpRows), 256, 0, (cudaStream_t)stream>>>(x1, R, d1, w2, b2, d2, w3, b3, d3, wc, bc, nc, wb, bb, nb,
                                                                        cls_score, bbox_pred);
    return finish_launch();
}
