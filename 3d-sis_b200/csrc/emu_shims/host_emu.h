// host_emu.h -- TEST INFRASTRUCTURE: compiles a CUDA source for the HOST so that its kernels' index arithmetic, shared-memory
// staging and barriers can be exercised without a GPU (g++ -DSIS3D_HOST_EMU -x c++ file.cu).  One CUDA block = blockDim
// std::threads that meet at a std::barrier for __syncthreads(); blocks run one after another, so function-local
// `__shared__` arrays (mapped to `static`) behave like per-block shared memory.  No warp-level intrinsics are modelled:
// only kernels that do not rely on them may be emulated.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
inline int cudaGetLastError() { return cudaSuccess; }

inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
inline std::barrier<> *emu_barrier = nullptr;
inline void __syncthreads() { emu_barrier->arrive_and_wait(); }
template <class T> inline T __ldg(const T *p) { return *p; }
using std::max;
using std::min;

template <class Kernel, class... Args>
void emu_launch(Kernel kernel, dim3 grid, dim3 block, Args... args) {
    const unsigned nthreads = block.x * block.y * block.z;
    std::barrier<> bar((std::ptrdiff_t)nthreads);
    emu_barrier = &bar;
    blockDim = block;
    gridDim = grid;
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t)
        pool.emplace_back([=, &bar]() {
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        blockIdx = dim3(bx, by, bz);
                        kernel(args...);
                        bar.arrive_and_wait();  // the next block reuses the `static` shared arrays
                    }
        });
    for (auto &th : pool) th.join();
    emu_barrier = nullptr;
}
