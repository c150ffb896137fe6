// mask_shell_zero.cu -- WORK IN PROGRESS (round 2), companion of mask_plan_dev.cu.  With a canvas of FIXED extents (so
// that its tensor maps can live in a captured graph) zeroing the whole canvas per scene would cost hundreds of MB; only
// the voxels a valid output can read outside its own crop have to be zero: the x-slabs just before and just after each
// crop and the planes y = h, z = l behind it (y = -1, z = -1 and everything beyond the canvas are zero-filled by TMA).
// One CTA per (crop, face); element type is opaque (bytes per voxel row = row_bytes).  Checked under host emulation
// against a brute-force "every out-of-crop neighbour of every crop voxel reads zero" test (tests/test_mask_plan_dev.py).
#ifdef SIS3D_HOST_EMU
#include "../emu_shims/host_emu.h"
#define SIS3D_LAUNCH(kernel, grid, block, stream, ...) emu_launch(kernel, grid, block, __VA_ARGS__)
#else
#include <cuda_runtime.h>
#define SIS3D_LAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif
#include <stdint.h>
#include "../../../include/sis3d.h"

namespace {
// n_kept lives in device memory (plan record); sizes int32[k][3]; xoff recomputed as the running sum of (w + 1)
__global__ void __launch_bounds__(256) mask_shell_zero_kernel(const int32_t *n_kept_ptr, const int32_t *sizes, int kcap, int Xc, int Yc,
                                                             int Zc, int row_bytes, char *canvas) {
    const int j = blockIdx.x, face = blockIdx.y;  // face 0: slab x = xoff-1, 1: slab x = xoff+w, 2: plane y = h, 3: plane z = l
    const int nk = min(*n_kept_ptr, kcap);
    if (j >= nk) return;
    int xoff = 0;
    for (int i = 0; i < j; ++i) xoff += sizes[3 * i] + 1;
    const int w = sizes[3 * j], h = sizes[3 * j + 1], l = sizes[3 * j + 2];
    const int hy = min(h + 1, Yc), lz = min(l + 1, Zc);
    int x0, x1, y0, y1, z0, z1;
    if (face == 0) { x0 = xoff - 1; x1 = xoff; y0 = 0; y1 = hy; z0 = 0; z1 = lz; }
    else if (face == 1) { x0 = xoff + w; x1 = x0 + 1; y0 = 0; y1 = hy; z0 = 0; z1 = lz; }
    else if (face == 2) { x0 = xoff; x1 = xoff + w; y0 = h; y1 = h + 1; z0 = 0; z1 = lz; }
    else { x0 = xoff; x1 = xoff + w; y0 = 0; y1 = h; z0 = l; z1 = l + 1; }
    if (x0 < 0 || x1 > Xc || y1 > Yc || z1 > Zc) {  // faces that fall outside the canvas are TMA's zero fill
        if (x0 < 0 || x0 >= Xc || y0 >= Yc || z0 >= Zc) return;
        x1 = min(x1, Xc); y1 = min(y1, Yc); z1 = min(z1, Zc);
    }
    const int nz = z1 - z0, ny = y1 - y0;
    const int64_t rows = (int64_t)(x1 - x0) * ny * nz;
    const int words = row_bytes / 16;  // 16-byte stores
    for (int64_t i = threadIdx.x; i < rows * words; i += blockDim.x) {
        const int64_t r = i / words;
        const int q = (int)(i - r * words);
        const int z = z0 + (int)(r % nz), y = y0 + (int)((r / nz) % ny), x = x0 + (int)(r / ((int64_t)nz * ny));
        float4 *p = reinterpret_cast<float4 *>(canvas + ((((int64_t)x * Yc + y) * Zc + z) * (int64_t)row_bytes)) + q;
        *p = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
}  // namespace

extern "C" int sis3d_mask_shell_zero(const int32_t *n_kept_dev, const int32_t *sizes_dev, int kcap, int Xc, int Yc, int Zc,
                                     int row_bytes, void *canvas, void *stream) {
    if (!n_kept_dev || !sizes_dev || !canvas || kcap <= 0 || Xc <= 0 || Yc <= 0 || Zc <= 0 || row_bytes <= 0 || (row_bytes & 15))
        return SIS3D_EINVAL;
    SIS3D_LAUNCH(mask_shell_zero_kernel, dim3(kcap, 4), dim3(256), stream, n_kept_dev, sizes_dev, kcap, Xc, Yc, Zc, row_bytes,
                 (char *)canvas);
    return cudaGetLastError() == cudaSuccess ? SIS3D_OK : SIS3D_ELAUNCH;
}
