// mask_plan_dev.cu -- WORK IN PROGRESS (round 2): the ragged mask stage's planner as a single-CTA kernel, so that the
// stage can join the CUDA graph of the static stage (no mid-scene D2H of the detection table, no host planning, no separate
// launches).  Same tables as the host planner sis3d_mask_plan_build (api.cu; canvas mode) but at FIXED offsets derived from
// the capacities (kcap kept RoIs, tcap bricks), which is what a captured graph needs: downstream kernels get constant
// pointers and read the live counts from the plan record.  cy/cz > 0 fix the canvas' y/z extents (static tensor maps);
// 0 = tight extents like the host planner.  Not compiled into libsis3d.so yet; not run on a GPU yet.  Its logic is
// checked bit for bit against the host planner under host emulation (tests/test_mask_plan_dev.py).
#ifdef SIS3D_HOST_EMU
#include "../emu_shims/host_emu.h"
#define SIS3D_LAUNCH(kernel, grid, block, stream, ...) emu_launch(kernel, grid, block, __VA_ARGS__)
#else
#include <cuda_runtime.h>
#define SIS3D_LAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif
#include <stdint.h>
#include "../../../include/sis3d.h"

typedef struct sis3d_mask_plan_dev {
    int32_t n_kept, canvas[3], n_tiles_tc, tiles_first, overflow, reserved;  // overflow: 1 = more kept RoIs / bricks than capacity
    int64_t total_voxels;
} sis3d_mask_plan_dev;

namespace {
constexpr int KMAX = 256;  // upper bound of kcap (smem arrays)

struct Layout { int64_t first, last, tiles, offs, cls, kept, sizes, bytes; };
__host__ __device__ inline Layout layout_of(int kcap, int tcap) {
    Layout l;
    l.first = 0;
    l.last = (int64_t)kcap * (int64_t)sizeof(sis3d_region);
    l.tiles = 2 * l.last;
    l.offs = l.tiles + (int64_t)tcap * 32;
    l.cls = l.offs + 8 * ((int64_t)kcap + 1);
    l.kept = l.cls + 4 * (int64_t)kcap;
    l.sizes = l.kept + 4 * (int64_t)kcap;
    l.bytes = l.sizes + 12 * (int64_t)kcap;
    return l;
}

__global__ void __launch_bounds__(256) mask_plan_kernel(const float *det, int n_max, int X, int Y, int Z, int ncls, int cy, int cz,
                                                       int kcap, int tcap, char *blob, sis3d_mask_plan_dev *plan) {
    __shared__ int s_kept[KMAX], s_xoff[KMAX], s_tb[KMAX], s_bb[KMAX];
    __shared__ long long s_voff[KMAX + 1];
    __shared__ int s_nk, s_ymax, s_zmax, s_overflow;
    const int t = threadIdx.x;
    if (t == 0) {
        int n = (int)det[15];  // RoI count travels in row 0, column 15 of the detection table (detect_decode_kernel)
        n = n < 0 ? 0 : (n > n_max ? n_max : n);
        int nk = 0, overflow = 0;
        for (int i = 0; i < n; ++i)
            if (det[i * 16 + 8] > 0.5f) {
                if (nk < kcap) s_kept[nk++] = i; else overflow = 1;
            }
        long long voff = 0, bricks = 0;
        int xoff = 0, tb = 0, ymax = 0, zmax = 0;
        for (int j = 0; j < nk; ++j) {
            const float *d = det + s_kept[j] * 16;
            const int w = (int)d[12] - (int)d[9], h = (int)d[13] - (int)d[10], l = (int)d[14] - (int)d[11];
            const long long vox = (long long)w * h * l;
            s_voff[j] = voff; s_xoff[j] = xoff; s_tb[j] = tb; s_bb[j] = (int)bricks;
            voff += vox;
            xoff += w + 1;
            tb += (int)((vox + SIS3D_CONV_TILE_M - 1) / SIS3D_CONV_TILE_M);
            bricks += (long long)((w + 3) / 4) * ((h + 3) / 4) * ((l + 7) / 8);
            ymax = h > ymax ? h : ymax;
            zmax = l > zmax ? l : zmax;
        }
        if (bricks > tcap) overflow = 1;
        s_voff[nk] = voff;
        s_nk = nk; s_overflow = overflow;
        s_ymax = cy > 0 ? cy : ymax;
        s_zmax = cz > 0 ? cz : zmax;
        plan->n_kept = nk;
        plan->canvas[0] = xoff; plan->canvas[1] = s_ymax; plan->canvas[2] = s_zmax;
        plan->n_tiles_tc = overflow ? 0 : (int)bricks;
        plan->tiles_first = tb;
        plan->overflow = overflow;
        plan->reserved = 0;
        plan->total_voxels = voff;
    }
    __syncthreads();
    const int nk = s_nk;
    const Layout L = layout_of(kcap, tcap);
    sis3d_region *first = (sis3d_region *)(blob + L.first), *last = (sis3d_region *)(blob + L.last);
    int32_t *tiles = (int32_t *)(blob + L.tiles);
    int64_t *offs = (int64_t *)(blob + L.offs);
    int32_t *cls = (int32_t *)(blob + L.cls), *kidx = (int32_t *)(blob + L.kept), *sizes = (int32_t *)(blob + L.sizes);
    const int64_t cs0 = (int64_t)s_ymax * s_zmax * 64, cs1 = (int64_t)s_zmax * 64, cs2 = 64;
    if (t == 0) offs[nk] = s_voff[nk];
    for (int j = t; j < nk; j += blockDim.x) {
        const float *d = det + s_kept[j] * 16;
        const int x0 = (int)d[9], y0 = (int)d[10], z0 = (int)d[11];
        const int w = (int)d[12] - x0, h = (int)d[13] - y0, l = (int)d[14] - z0;
        const int xoff = s_xoff[j];
        sis3d_region f = {}, q = {};
        f.in_off = ((int64_t)x0 * Y + y0) * Z + z0;
        f.in_dim[0] = f.out_dim[0] = w; f.in_dim[1] = f.out_dim[1] = h; f.in_dim[2] = f.out_dim[2] = l;
        f.in_stride[0] = (int64_t)Y * Z; f.in_stride[1] = Z; f.in_stride[2] = 1;
        f.tile_begin = s_tb[j];
        f.out_off = xoff * cs0;
        f.out_stride[0] = cs0; f.out_stride[1] = cs1; f.out_stride[2] = cs2;
        q.in_dim[0] = q.out_dim[0] = w; q.in_dim[1] = q.out_dim[1] = h; q.in_dim[2] = q.out_dim[2] = l;
        q.out_off = s_voff[j] * ncls;
        q.tile_begin = s_tb[j];
        q.in_off = xoff * cs0;
        q.in_stride[0] = cs0; q.in_stride[1] = cs1; q.in_stride[2] = cs2;
        first[j] = f;
        last[j] = q;
        offs[j] = s_voff[j];
        cls[j] = (int32_t)d[7];
        kidx[j] = s_kept[j];
        sizes[3 * j] = w; sizes[3 * j + 1] = h; sizes[3 * j + 2] = l;
        if (!s_overflow) {
            int tile_no = s_bb[j];
            for (int bx = 0; bx < w; bx += 4)
                for (int by = 0; by < h; by += 4)
                    for (int bz = 0; bz < l; bz += 8) {
                        int32_t *tt = tiles + (int64_t)tile_no * 8;
                        tt[0] = xoff + bx; tt[1] = by; tt[2] = bz;
                        tt[3] = xoff + w; tt[4] = h; tt[5] = l; tt[6] = tt[7] = 0;
                        ++tile_no;
                    }
        }
    }
}
}  // namespace

extern "C" size_t sis3d_mask_plan_device_bytes(int kcap, int tcap) { return (size_t)layout_of(kcap, tcap).bytes; }
extern "C" void sis3d_mask_plan_device_layout(int kcap, int tcap, int64_t *offsets7) {
    const Layout l = layout_of(kcap, tcap);
    offsets7[0] = l.first; offsets7[1] = l.last; offsets7[2] = l.tiles; offsets7[3] = l.offs; offsets7[4] = l.cls;
    offsets7[5] = l.kept; offsets7[6] = l.sizes;
}
extern "C" int sis3d_mask_plan_device(const float *det, int n_max, int X, int Y, int Z, int ncls, int cy, int cz, int kcap, int tcap,
                                      void *blob, sis3d_mask_plan_dev *plan, void *stream) {
    if (!det || !blob || !plan || n_max <= 0 || kcap <= 0 || kcap > KMAX || tcap <= 0) return SIS3D_EINVAL;
    SIS3D_LAUNCH(mask_plan_kernel, dim3(1), dim3(256), stream, det, n_max, X, Y, Z, ncls, cy, cz, kcap, tcap, (char *)blob, plan);
    return cudaGetLastError() == cudaSuccess ? SIS3D_OK : SIS3D_ELAUNCH;
}
