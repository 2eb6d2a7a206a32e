// hd.h -- qualifiers and exactly-rounded arithmetic helpers shared by all kernels.
//
// Kernel logic lives in `MWB_DEV inline` functions.  Under nvcc they are __device__ code;
// the test-only host simulator (tests/hostsim, built with g++ -ffp-contract=off) compiles
// the very same functions for the CPU so kernel logic can be debugged without a GPU.
// libmwb.so itself never executes them on the host.
//
// Bit-exact parts of the pipeline (physics in float64, coverage / depth in float32) must
// not be contracted into FMAs by the compiler: they use the *_rn wrappers below.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef __CUDACC__
struct uint2 { unsigned int x, y; };
#endif

#ifdef __CUDACC__
#define MWB_DEV __device__ __forceinline__
#define MWB_DEVM __device__ __forceinline__
#define MWB_DEV_NOINLINE __device__ __noinline__
#define MWB_ALIGN16 __align__(16)
#define MWB_DEVCONST __device__ const
#else
#define MWB_DEV static inline
#define MWB_DEVM inline
#define MWB_DEV_NOINLINE static
#define MWB_ALIGN16 alignas(16)
#define MWB_DEVCONST static const
#endif

#ifdef __CUDA_ARCH__
MWB_DEV double d_mul(double a, double b) { return __dmul_rn(a, b); }
MWB_DEV double d_add(double a, double b) { return __dadd_rn(a, b); }
MWB_DEV double d_sub(double a, double b) { return __dsub_rn(a, b); }
MWB_DEV double d_div(double a, double b) { return __ddiv_rn(a, b); }
MWB_DEV double d_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
MWB_DEV double d_sqrt(double a) { return __dsqrt_rn(a); }
MWB_DEV float f_mul(float a, float b) { return __fmul_rn(a, b); }
MWB_DEV float f_add(float a, float b) { return __fadd_rn(a, b); }
MWB_DEV float f_sub(float a, float b) { return __fsub_rn(a, b); }
MWB_DEV float f_div(float a, float b) { return __fdiv_rn(a, b); }
MWB_DEV float f_sqrt(float a) { return __fsqrt_rn(a); }
MWB_DEV uint64_t umulhi64(uint64_t a, uint64_t b) { return __umul64hi(a, b); }
#else
// host build: compiled with -ffp-contract=off, so plain operators are single IEEE ops
MWB_DEV double d_mul(double a, double b) { return a * b; }
MWB_DEV double d_add(double a, double b) { return a + b; }
MWB_DEV double d_sub(double a, double b) { return a - b; }
MWB_DEV double d_div(double a, double b) { return a / b; }
MWB_DEV double d_fma(double a, double b, double c) { return fma(a, b, c); }
MWB_DEV double d_sqrt(double a) { return sqrt(a); }
MWB_DEV float f_mul(float a, float b) { return a * b; }
MWB_DEV float f_add(float a, float b) { return a + b; }
MWB_DEV float f_sub(float a, float b) { return a - b; }
MWB_DEV float f_div(float a, float b) { return a / b; }
MWB_DEV float f_sqrt(float a) { return sqrtf(a); }
MWB_DEV uint64_t umulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
#endif

MWB_DEV uint64_t d2bits(double d) {
  union { double d; uint64_t u; } c;
  c.d = d;
  return c.u;
}
