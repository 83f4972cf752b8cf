// TEST INFRASTRUCTURE: stand-in for <cuda_runtime.h> so that g++ can compile the device code of lbm_b200/csrc/kernels.cuh for the CPU
// (tests/c/kernels_harness.cpp).  Qualifiers vanish, thread indices are variables a driver loop sets, IEEE intrinsics are the plain
// operators (the harness is compiled with -ffp-contract=off, so a + b is __dadd_rn(a, b)), cache-hinted loads / stores are plain ones.
// __syncthreads() is a real barrier over the OS threads the harness starts for one block (k_step runs that way, with its shared
// memory as function-local statics: one block at a time); k_residual (warp shuffles) compiles but is never called.
#pragma once
#include <pthread.h>

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

struct uint4 { unsigned x, y, z, w; };
struct FakeDim3 { unsigned x = 0, y = 0, z = 0; };
static FakeDim3 blockIdx, blockDim, gridDim;
static thread_local FakeDim3 threadIdx;

static pthread_barrier_t g_block_barrier;
static bool              g_block_barrier_on = false; // single-threaded driver loops: __syncthreads() has nothing to wait for
static inline void __syncthreads() {
  if(g_block_barrier_on) pthread_barrier_wait(&g_block_barrier);
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) {}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

static inline long long __double_as_longlong(double x) { long long b; __builtin_memcpy(&b, &x, 8); return b; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return static_cast<unsigned long long>((static_cast<unsigned __int128>(a) * b) >> 64); }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
