// emu_shims/cuda_runtime.h -- TEST INFRASTRUCTURE: stands in for <cuda_runtime.h> when a libsis3d source is compiled for
// the host (tools/cuda_host_emu.py); see host_emu.h.
#pragma once
#include <cstddef>
#include <cstring>
#include "host_emu.h"
typedef int cudaError_t;
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> inline int cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
inline int cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
// warp intrinsics are NOT modelled: kernels that use them must not be emulated (the shim only lets shared headers compile)
template <class T> inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
