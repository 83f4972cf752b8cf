// solver_common.cuh -- what the translation units of liblbm_b200.so share: error reporting, device buffers, the solver interface
// behind the C ABI.  The fused solver is instantiated per lattice / precision in its own translation unit (fused_*.cu) so that the
// library builds in parallel.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../../include/lbm_b200.h"
#include "kernels.cuh"
#include "plan.hpp"
#include "nccl_dyn.hpp"

namespace lbm_impl {


extern thread_local std::string g_error; // defined in solver.cu

inline int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

#define NCCL_TRY(expr)                                                                                  \
  do {                                                                                                  \
    ncclResult_t r_ = (expr);                                                                           \
    if(r_ != ncclSuccess)                                                                               \
      return fail(LBM_B200_ECUDA, std::string(#expr) + ": " + lbm::nccl_api().GetErrorString(r_));      \
  } while(0)

#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if(e_ != cudaSuccess)                                                                               \
      return fail(LBM_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));                  \
  } while(0)

template <class T>
struct DevBuf {
  T*     p = nullptr;
  size_t n = 0;
  ~DevBuf() { if(p) cudaFree(p); }
  cudaError_t alloc(size_t count) {
    if(p) cudaFree(p);
    p = nullptr;
    n = count;
    if(count == 0) return cudaSuccess;
    return cudaMalloc(&p, count * sizeof(T));
  }
  cudaError_t upload(const std::vector<T>& h) {
    cudaError_t e = alloc(h.size());
    if(e != cudaSuccess || h.empty()) return e;
    return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  }
  size_t bytes() const { return n * sizeof(T); }
};

struct SolverBase {
  virtual ~SolverBase() = default;
  virtual int init()                                                     = 0;
  virtual int step(int64_t n, float* ms_total, float* ms_main)           = 0;
  virtual int sync()                                                     = 0;
  virtual int residual(double* out, int32_t* diverged)                   = 0;
  virtual int get_populations(double* f, double* fold)                   = 0;
  virtual int set_populations(const double* f, const double* fold)       = 0;
  virtual int get_vars(double* vars, double* varsold)                    = 0;
  virtual int get_moments(double* m)                                     = 0;
  virtual void stats(lbm_b200_stats* st) const                           = 0;
  virtual int64_t owned() const                                          = 0;
  virtual int encode_output(const uint8_t* keep, char* text, int64_t capacity, int64_t* offsets) {
    (void)keep; (void)text; (void)capacity; (void)offsets;
    return fail(LBM_B200_EUNSUP, "no device-side output encoder for this solver kind (use lbm_b200_get_moments and a host writer)");
  }
  virtual int p2p_export(void* blob) { (void)blob; return fail(LBM_B200_EUNSUP, "peer-to-peer halo: only the fused solver is partitioned"); }
  virtual int p2p_import(int32_t nranks, const void* blobs) { (void)nranks; (void)blobs; return fail(LBM_B200_EUNSUP, "peer-to-peer halo: only the fused solver is partitioned"); }
  virtual int debug_plan(lbm_b200_plan_view* out) { (void)out; return fail(LBM_B200_EUNSUP, "no device plan for this solver kind"); }
  lbm_b200_config cfg{};
  lbm::PlanInput  in;
  cudaStream_t    stream = nullptr;
  ncclComm_t      comm   = nullptr;
  int             comm_rank = 0, comm_size = 1;
  bool            inited = false;
  int64_t         t      = 0;
};

// factories of the fused solver, one translation unit each (fused_*.cu)
SolverBase* make_fused_d2q9_f64();
SolverBase* make_fused_d2q9_f32();
SolverBase* make_fused_d3q19_f64();
SolverBase* make_fused_d3q19_f32();
SolverBase* make_fused_d3q27_f64();
SolverBase* make_fused_d3q27_f32();

} // namespace lbm_impl
