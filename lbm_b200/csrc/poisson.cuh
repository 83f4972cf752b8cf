// poisson.cuh -- the reference's Poisson equation types on the GPU.
//
// LBEquationType::Poisson of the reference (/root/reference/src/lbm/solver.cpp:283-293 initial condition, :540-543 potential,
// :562-566 + equilibrium_func.h:149-163 equilibrium, :589-612 collision with the source term; lattices D1Q3 / D2Q5 / D2Q9,
// src/lbm/constants.h:248-310) with its boundary conditions LBMBnd_DirichletNEEM (src/lbm/bnd/bnd_dirichlet.h:250-368) and
// LBMBnd_NeumannNEEM (src/lbm/bnd/bnd_neumann.h:15-66).  These conditions rewrite whole cells from the post-streaming state of an
// extrapolation cell and from m_vars as earlier conditions of the same pass left it, so -- like the wet-node walls
// (sequential.cuh) -- the reference's passes run as kernels in the reference's order, explicit m_f / m_fold / m_feq / m_vars /
// m_varsold in the reference's array-of-structures layout, one variable per cell (the potential).  fp64, the reference's operation
// order with intrinsics the compiler never contracts: bit-identical to the reference (pinned for the CPU oracle on the five Poisson
// cases of the reference's test/run.sh, tests/test_oracle_golden.py).  The reference's cases have 256 .. 65 536 cells: this path is
// launch-latency bound, not a bandwidth kernel.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lbm {
namespace poisson {

struct Lat {
  int    D, Q;
  double w[9];    // LBMethod<>::m_weights
  double pw[9];   // LBMethod<>::m_poissonWeights
  double inv_1mw; // 1.0 / (1.0 - w[Q-1]), moments.h:89
};

struct State {
  double *f, *fold, *feq, *vars, *varsold;
  const int32_t* pull;  // [n*(Q-1)] inverse of the push table (highest source wins)
  const int64_t* nghbr; // [n*(Q-1)] push table
  int64_t        n;
  double         omega, om1;
  double         dt_diff; // m_dt * diffusivity, solver.cpp:606-609
  double         rate2;   // poisson_D * poisson_D
};

struct Bc {
  int            neumann;
  int64_t        n;
  const int64_t* cells;  // [n]
  const int64_t* ext;    // [n] extrapolation cell
  const int64_t* ext2;   // [n] Neumann: neighbour of ext in the extrapolation direction
  double*        values; // [n] m_value (Neumann rewrites it every step)
  double         grad;
};

// moments.h:83-91 / solver.cpp:540-543: 1/(1 - w_rest) * (sum of the moving populations, ascending, from 0.0)
__device__ __forceinline__ double potential_of(const Lat& L, const double* __restrict__ fo) {
  double acc = 0.0; // std::accumulate(..., 0.0): starting from +0.0 also fixes the sign of a zero sum
  for(int i = 0; i < L.Q - 1; ++i) acc = __dadd_rn(acc, fo[i]);
  return __dmul_rn(L.inv_1mw, acc);
}

// passes 2-4 (solver.cpp:513-613): potential, equilibrium, BGK collision + source term
__global__ void k_cell(State s, Lat L) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(c >= s.n) return;
  const int     Q   = L.Q;
  const double* fo  = &s.fold[c * Q];
  const double  phi = potential_of(L, fo);
  s.vars[c]         = phi;
  const double rhs  = __dmul_rn(s.rate2, phi);
  for(int i = 0; i < Q; ++i) {
    const double fe = i < Q - 1 ? __dmul_rn(L.w[i], phi) : __dmul_rn(__dsub_rn(L.w[Q - 1], 1.0), phi);
    s.feq[c * Q + i] = fe;
    double f = __dadd_rn(__dmul_rn(s.om1, fo[i]), __dmul_rn(s.omega, fe));
    if(i != Q - 1) f = __dadd_rn(f, __dmul_rn(__dmul_rn(s.dt_diff, L.pw[i]), rhs));
    s.f[c * Q + i] = f;
  }
}

// output(): the potential of the current m_fold (solver.cpp:336)
__global__ void k_potential(State s, Lat L, double* out) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(c < s.n) out[c] = potential_of(L, &s.fold[c * L.Q]);
}

// pass 7 as a pull (solver.cpp:715-740)
__global__ void k_stream(State s, Lat L) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(t >= s.n) return;
  const int Q = L.Q;
  for(int j = 0; j < Q - 1; ++j) {
    const int32_t src = s.pull[t * (Q - 1) + j];
    if(src >= 0) s.fold[t * Q + j] = s.f[static_cast<int64_t>(src) * Q + j];
  }
  s.fold[t * Q + Q - 1] = s.f[t * Q + Q - 1];
}

// initialCondition (solver.cpp:283-293): f = fold = feq = w * potential, rest population (w_rest - 1) * potential
__global__ void k_init(State s, Lat L) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(c >= s.n) return;
  const int    Q   = L.Q;
  const double phi = s.vars[c];
  for(int i = 0; i < Q; ++i) {
    const double v = i < Q - 1 ? __dmul_rn(L.w[i], phi) : __dmul_rn(__dsub_rn(L.w[Q - 1], 1.0), phi);
    s.feq[c * Q + i] = s.f[c * Q + i] = s.fold[c * Q + i] = v;
  }
}

// LBMBnd_NeumannNEEM::apply, first loop (bnd_neumann.h:52-59): value = (4 phi(ext) - phi(ext2) + grad) / 3 with phi(ext2) recomputed
// from the post-streaming populations and phi(ext) as m_vars holds it at this point
__global__ void k_neumann_value(State s, Lat L, Bc b) {
  const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(k >= b.n) return;
  const int64_t e = b.ext[k], e2 = b.ext2[k];
  const double  p2 = potential_of(L, &s.fold[e2 * L.Q]);
  s.vars[e2]   = p2;
  b.values[k] = __ddiv_rn(__dadd_rn(__dsub_rn(__dmul_rn(4.0, s.vars[e]), p2), b.grad), 3.0);
}

// LBMBnd_DirichletNEEM::apply, calcDensity over the extrapolation cells (bnd_dirichlet.h:349-350)
__global__ void k_ext_potential(State s, Lat L, Bc b) {
  const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(k >= b.n) return;
  const int64_t e = b.ext[k];
  s.vars[e] = potential_of(L, &s.fold[e * L.Q]);
}

// LBMBnd_DirichletNEEM::apply, the entries (bnd_dirichlet.h:352-365)
__global__ void k_dirichlet(State s, Lat L, Bc b) {
  const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(k >= b.n) return;
  const int     Q = L.Q;
  const int64_t c = b.cells[k], e = b.ext[k];
  const double  val = b.values[k], pe = s.vars[e];
  s.vars[c] = val;
  for(int i = 0; i < Q; ++i) {
    const double w = i < Q - 1 ? L.w[i] : __dsub_rn(L.w[Q - 1], 1.0);
    s.fold[c * Q + i] = __dsub_rn(__dadd_rn(__dmul_rn(w, val), s.fold[e * Q + i]), __dmul_rn(w, pe));
  }
}

// sumAbsDiff (solver.cpp:809-815) over one variable, fixed-shape partial sums
__global__ void k_residual(const double* __restrict__ v, const double* __restrict__ vo, int64_t n, double* __restrict__ partial) {
  __shared__ double sh[256];
  double acc = 0;
  for(int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; c < n; c += static_cast<int64_t>(gridDim.x) * blockDim.x)
    acc += fabs(v[c] - vo[c]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for(int o = blockDim.x / 2; o > 0; o >>= 1) {
    if(static_cast<int>(threadIdx.x) < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if(threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

} // namespace poisson
} // namespace lbm
