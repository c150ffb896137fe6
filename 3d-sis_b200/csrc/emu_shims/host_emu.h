// host_emu.h -- TEST INFRASTRUCTURE: compiles a CUDA source for the HOST so that its kernels' index arithmetic, shared-memory
// staging and barriers can be exercised without a GPU (g++ -DSIS3D_HOST_EMU -x c++ file.cu).  One CUDA block = blockDim
// std::threads that meet at a std::barrier for __syncthreads(); blocks run one after another, so function-local
// `__shared__` arrays (mapped to `static`) behave like per-block shared memory.  Full-mask warp shuffles / ballots are
// modelled with a per-warp barrier and scratch row; TMA, tcgen05 and partial-mask collectives are not.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct uint2 { unsigned x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
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

// ---- warp collectives (full-mask use only): the 32 lanes of a warp exchange values through a per-warp scratch row and
// meet at a per-warp barrier; block sizes must be multiples of 32
#include <atomic>
#include <memory>
inline thread_local unsigned emu_lane = 0, emu_warp = 0;
inline std::vector<std::unique_ptr<std::barrier<>>> emu_warp_barriers;
inline std::vector<unsigned long long> emu_warp_scratch;  // [warps][32]
inline std::atomic<int> emu_count[2];
inline thread_local int emu_count_phase = 0;
template <class T> inline unsigned long long emu_bits(T v) { unsigned long long b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T emu_unbits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <class T> inline T emu_exchange(T v, int src_lane) {
    unsigned long long *row = emu_warp_scratch.data() + (size_t)emu_warp * 32;
    __atomic_store_n(&row[emu_lane], emu_bits(v), __ATOMIC_RELAXED);
    emu_warp_barriers[emu_warp]->arrive_and_wait();
    const T r = (src_lane >= 0 && src_lane < 32) ? emu_unbits<T>(__atomic_load_n(&row[src_lane], __ATOMIC_RELAXED)) : v;
    emu_warp_barriers[emu_warp]->arrive_and_wait();
    return r;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu_exchange(v, (int)(emu_lane ^ (unsigned)lane_mask)); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int delta) { return emu_exchange(v, (int)emu_lane - delta); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int delta) { return emu_exchange(v, (int)emu_lane + delta); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned long long *row = emu_warp_scratch.data() + (size_t)emu_warp * 32;
    __atomic_store_n(&row[emu_lane], (unsigned long long)(pred != 0), __ATOMIC_RELAXED);
    emu_warp_barriers[emu_warp]->arrive_and_wait();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (unsigned)__atomic_load_n(&row[i], __ATOMIC_RELAXED) << i;
    emu_warp_barriers[emu_warp]->arrive_and_wait();
    return m;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_barriers[emu_warp]->arrive_and_wait(); }
inline int __syncthreads_count(int pred) {
    const int ph = emu_count_phase;
    emu_count_phase ^= 1;
    if (pred) emu_count[ph].fetch_add(1, std::memory_order_relaxed);
    emu_barrier->arrive_and_wait();
    const int r = emu_count[ph].load(std::memory_order_relaxed);
    emu_barrier->arrive_and_wait();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) emu_count[ph].store(0, std::memory_order_relaxed);
    return r;
}
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
    const unsigned nwarps = (nthreads + 31) / 32;
    emu_warp_barriers.clear();
    for (unsigned w = 0; w < nwarps; ++w)
        emu_warp_barriers.emplace_back(std::make_unique<std::barrier<>>((std::ptrdiff_t)std::min(32u, nthreads - 32 * w)));
    emu_warp_scratch.assign((size_t)nwarps * 32, 0ull);
    emu_count[0] = 0;
    emu_count[1] = 0;
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t)
        pool.emplace_back([=, &bar]() {
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            emu_lane = t % 32;
            emu_warp = t / 32;
            emu_count_phase = 0;
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
