// io.cu -- GPU-side decode of the voxel block of a `.scene` / `.chunk` file.
//
// The container stores the signed distance field x-fastest (index = x + X*(y + Y*z); written by
// datagen/SceneSampler/main.cpp:348-395); the reference reader reshapes it with order='F' and builds the network's two input
// channels on the host with four full-volume numpy passes (lib/datasets/dataset.py:50-68: clip to +-TRUNCATED, abs, `> -1`,
// concatenate) and crops the height (dataset.py:192-205: data[:, :, :maxHeight, :]).  Here the raw block goes to the device as
// it sits in the file and ONE kernel writes the network input [2][X][Yk][Z] (z fastest): a 32x32 (x,z) tile transpose through
// shared memory per y, both channels computed on the way.  Bit-identical to the reference's arithmetic (clip/abs/compare are
// exact operations).
#include "common.cuh"

namespace sis3d {

__global__ void __launch_bounds__(256) chunk_decode_kernel(const float *sdf, int X, int Y, int Z, int Yk, float trunc, float *out) {
    __shared__ float tile[32][33];
    const int y = blockIdx.z;
    const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {  // read: x fastest
        const int z = z0 + r, x = x0 + tx;
        tile[r][tx] = (z < Z && x < X) ? __ldg(sdf + ((int64_t)z * Y + y) * X + x) : 0.f;
    }
    __syncthreads();
    const int64_t plane = (int64_t)X * Yk * Z;
    for (int r = ty; r < 32; r += 8) {  // write: z fastest
        const int x = x0 + r, z = z0 + tx;
        if (x < X && z < Z) {
            const float v = tile[tx][r];
            const int64_t o = ((int64_t)x * Yk + y) * Z + z;
            out[o] = fabsf(fminf(fmaxf(v, -trunc), trunc));  // np.abs(np.clip(sdf, -T, T))
            out[plane + o] = v > -1.f ? 1.f : 0.f;           // np.greater(sdf, -1)
        }
    }
}

}  // namespace sis3d
using namespace sis3d;

extern "C" int sis3d_chunk_decode(const float *sdf_xfast, int X, int Y, int Z, int y_keep, float truncation, float *data, void *stream) {
    if (!sdf_xfast || !data || X <= 0 || Y <= 0 || Z <= 0 || y_keep <= 0 || !(truncation > 0.f)) return SIS3D_EINVAL;
    const int Yk = y_keep < Y ? y_keep : Y;
    dim3 grid(cdiv(X, 32), cdiv(Z, 32), Yk);
    if (grid.y > 65535u || grid.z > 65535u) return SIS3D_EUNSUPPORTED;
    chunk_decode_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sdf_xfast, X, Y, Z, Yk, truncation, data);
    return finish_launch();
}
