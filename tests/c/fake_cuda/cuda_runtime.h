// TEST INFRASTRUCTURE: stand-in for <cuda_runtime.h> so that g++ can compile the device code of lbm_b200/csrc/kernels.cuh for the CPU
// (tests/c/kernels_harness.cpp).  Qualifiers vanish, thread indices are variables a driver loop sets, IEEE intrinsics are the plain
// operators (the harness is compiled with -ffp-contract=off, so a + b is __dadd_rn(a, b)), cache-hinted loads / stores are plain ones.
// Kernels that need barriers or shuffles (k_step, k_residual) compile but are never called.
#pragma once
#include <cmath>
#include <cstdint>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static
#define __grid_constant__
#define __launch_bounds__(...)

struct FakeDim3 { unsigned x = 0, y = 0, z = 0; };
static FakeDim3 blockIdx, blockDim, threadIdx, gridDim;

static inline void __syncthreads() {}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }

static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
