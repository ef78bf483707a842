// gf_hd.h -- host/device function qualifier.  Under nvcc everything is __host__ __device__; under a
// plain C++ compiler (tests/cpu_emul, test-only) a minimal float2 / double2 stand-in is provided.
#pragma once
#include <stdint.h>
#include <math.h>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define GF_HD __host__ __device__ __forceinline__
#else
#define GF_HD inline
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif
