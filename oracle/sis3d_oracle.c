/*
 * oracle/sis3d_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
 *
 * Plain-C restatement of the two native operators of the 3D-SIS inference hot path and of the
 * RPN proposal post-process, used as the CPU checker for the sm_100a kernels.  Each function
 * cites the reference lines it follows (paths relative to the reference root).
 *
 * Build: make -C oracle   ->  oracle/liboracle.so
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * 3D IoU with the inclusive "+1" extent convention.
 *   numpy path : lib/layer_utils/nms/pth_nms.py:22,31-40   (no FMA contraction)
 *   CUDA path  : lib/layer_utils/nms/src/cuda/nms_kernel.cu:11-31 (nvcc contracts two FMAs)
 * `fma_mode` = 0 restates the numpy arithmetic; 1 restates the FMA pattern nvcc emits for devIoU
 * (t = fma(wb*hb, lb, Sa), see DESIGN.md "NMS arithmetic"); both are float32 throughout.
 * ------------------------------------------------------------------------------------------- */
static float iou3d(const float *a, const float *b, int fma_mode)
{
    float left = fmaxf(a[0], b[0]), top = fmaxf(a[1], b[1]), front = fmaxf(a[2], b[2]);
    float right = fminf(a[3], b[3]), bottom = fminf(a[4], b[4]), back = fminf(a[5], b[5]);
    volatile float w = fmaxf(right - left + 1.0f, 0.f);
    volatile float h = fmaxf(bottom - top + 1.0f, 0.f);
    volatile float l = fmaxf(back - front + 1.0f, 0.f);
    volatile float wa = a[3] - a[0] + 1.0f, ha = a[4] - a[1] + 1.0f, la = a[5] - a[2] + 1.0f;
    volatile float wb = b[3] - b[0] + 1.0f, hb = b[4] - b[1] + 1.0f, lb = b[5] - b[2] + 1.0f;
    volatile float wh = w * h;
    volatile float inter = wh * l;
    volatile float pa = wa * ha, pb = wb * hb;
    if (!fma_mode) {
        volatile float Sa = pa * la, Sb = pb * lb;
        volatile float s = Sa + Sb;
        volatile float u = s - inter;
        return inter / u;
    } else {
        /* nvcc default (-fmad=true) SASS for `Sa + Sb - interS` (a = row box, loop-invariant Sa kept as
         * two FMULs): t = FFMA(wb*hb, lb, Sa); u = FADD(t, -inter); IEEE div.  Verified in the sm_100a
         * SASS of the unmodified reference kernel ("NMS arithmetic"). */
        volatile float Sa = pa * la;
        volatile float t = fmaf(pb, lb, Sa);
        volatile float u = t - inter;
        return inter / u;
    }
}

/* Greedy NMS on score-sorted boxes; suppress when IoU > thresh (CPU keeps `ovr <= thresh`,
 * lib/layer_utils/nms/pth_nms.py:26-43; GPU tests `> thresh`, nms_kernel.cu:70 + host reduce
 * lib/layer_utils/nms/src/nms_cuda.c:41-59).  Returns number kept; keep[] holds indices. */
int oracle_nms3d(const float *boxes, int n, float thresh, int fma_mode, int64_t *keep)
{
    unsigned char *dead = (unsigned char *)calloc((size_t)n + 1, 1);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        keep[nk++] = i;
        for (int j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            if (iou3d(boxes + 6 * i, boxes + 6 * j, fma_mode) > thresh) dead[j] = 1;
        }
    }
    free(dead);
    return nk;
}

/* Pairwise IoU table (n x n, row-major) for tests that need to know how close to the
 * threshold a pair is. */
void oracle_iou_matrix(const float *boxes, int n, int fma_mode, float *out)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) out[(size_t)i * n + j] = iou3d(boxes + 6 * i, boxes + 6 * j, fma_mode);
}

/* ---------------------------------------------------------------------------------------------
 * 3D RoI max pooling forward with argmax.
 *   CUDA : lib/layer_utils/roi_pooling/src/cuda/roi_pooling_kernel.cu:15-109
 *   CPU  : lib/layer_utils/roi_pooling/src/roi_pooling.c:6-124 (no argmax)
 * features [C,W,H,L] (batch 1), rois [n,6], out [n,C,pw,ph,pl], argmax int32 same shape (may be NULL).
 * ------------------------------------------------------------------------------------------- */
int oracle_roi_pool3d(const float *feat, int C, int W, int H, int L, const float *rois, int n,
                      int pw_, int ph_, int pl_, float scale, float *out, int32_t *argmax)
{
    for (int r = 0; r < n; ++r) {
        const float *roi = rois + 6 * r;
        int sw = (int)floor(roi[0] * scale), sh = (int)floor(roi[1] * scale), sl = (int)floor(roi[2] * scale);
        int ew = (int)ceil(roi[3] * scale), eh = (int)ceil(roi[4] * scale), el = (int)ceil(roi[5] * scale);
        int rw = (int)fmaxf((float)(ew - sw), 1.f), rh = (int)fmaxf((float)(eh - sh), 1.f),
            rl = (int)fmaxf((float)(el - sl), 1.f);
        float bw = (float)rw / (float)pw_, bh = (float)rh / (float)ph_, bl = (float)rl / (float)pl_;
        for (int c = 0; c < C; ++c)
            for (int pw = 0; pw < pw_; ++pw)
                for (int ph = 0; ph < ph_; ++ph)
                    for (int pl = 0; pl < pl_; ++pl) {
                        int ws = (int)floorf((float)pw * bw), hs = (int)floorf((float)ph * bh),
                            ls = (int)floorf((float)pl * bl);
                        int we = (int)ceilf((float)(pw + 1) * bw), he = (int)ceilf((float)(ph + 1) * bh),
                            le = (int)ceilf((float)(pl + 1) * bl);
                        ws = (int)fminf(fmaxf((float)(ws + sw), 0.f), (float)W);
                        hs = (int)fminf(fmaxf((float)(hs + sh), 0.f), (float)H);
                        ls = (int)fminf(fmaxf((float)(ls + sl), 0.f), (float)L);
                        we = (int)fminf(fmaxf((float)(we + sw), 0.f), (float)W);
                        he = (int)fminf(fmaxf((float)(he + sh), 0.f), (float)H);
                        le = (int)fminf(fmaxf((float)(le + sl), 0.f), (float)L);
                        int empty = (he <= hs) || (we <= ws) || (le <= ls);
                        float best = empty ? 0.f : -FLT_MAX;
                        int besti = -1;
                        for (int w = ws; w < we; ++w)
                            for (int h = hs; h < he; ++h)
                                for (int l = ls; l < le; ++l) {
                                    int idx = (c * W + w) * H * L + h * L + l;
                                    if (feat[idx] > best) { best = feat[idx]; besti = idx; }
                                }
                        size_t o = ((((size_t)r * C + c) * pw_ + pw) * ph_ + ph) * pl_ + pl;
                        out[o] = best;
                        if (argmax) argmax[o] = besti;
                    }
    }
    return 1;
}
