// Minimal CUDA intrinsics for compiling planar_kernels.cuh as host C++ (debug / CPU tests only).
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>
#define __host__
#define __device__
#define __forceinline__ inline
#define DEVI inline
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline double __longlong_as_double(long long v) { double f; memcpy(&f, &v, 8); return f; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
