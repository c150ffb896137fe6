// common.cuh -- shared helpers for libsis3d (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sis3d.h"

namespace sis3d {

extern unsigned long long g_launch_count;  // host-side counter, bumped once per kernel launch

inline int finish_launch(int n_launches = 1) {
    g_launch_count += (unsigned long long)n_launches;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SIS3D_OK : SIS3D_ELAUNCH;
}

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace sis3d
