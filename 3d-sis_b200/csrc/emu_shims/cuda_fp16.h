// emu_shims/cuda_fp16.h -- TEST INFRASTRUCTURE: __half for host emulation builds (IEEE binary16 through the compiler's _Float16)
#pragma once
struct __half { _Float16 v; };
inline __half __float2half_rn(float f) { return __half{(_Float16)f}; }
inline float __half2float(__half h) { return (float)h.v; }
struct __half2 { __half x, y; };
inline __half2 __floats2half2_rn(float a, float b) { return __half2{__float2half_rn(a), __float2half_rn(b)}; }
