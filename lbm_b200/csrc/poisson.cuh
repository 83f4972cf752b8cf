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
#ifndef LBM_POISSON_HOST_HARNESS // tests/c/poisson_harness.cpp compiles the kernel bodies for the CPU with its own stand-ins
#include <cuda_runtime.h>
#endif

#include <cstdint>
#include <string>
#include <vector>

#include "plan.hpp"

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

// ---- host side: everything lbm_b200_init derives from the caller's tables before the first launch (no CUDA call in here, so the CPU
// harness of the tests runs exactly this code)
struct HostBc {
  int                  neumann = 0;
  std::vector<int64_t> cells, ext, ext2;
  std::vector<double>  values;
  double               grad = 0;
};
struct HostSetup {
  Lat                  lat{};
  std::vector<int32_t> pull;   // [n*(Q-1)]
  std::vector<double>  vars0;  // [n] initial potential: the Dirichlet values (initCnd, bnd_dirichlet.h:335-341), 0 elsewhere
  std::vector<HostBc>  bcs;
  double omega = 1, om1 = 0, dt_diff = 0, rate2 = 0;
  int    code = 0;             // 0, or the LBM_B200_E* code (-1 EINVAL, -5 EUNSUP)
  std::string error;
};

inline bool prepare(const PlanInput& in, double omega, HostSetup& S) {
  const LatticeRT& LR = in.L;
  const int     Q = LR.Q, QM = Q - 1, D = LR.D;
  const int64_t N = in.n;
  auto bad = [&](int code, const char* msg) { S.code = code; S.error = msg; return false; };
  if(!((D == 1 && Q == 3) || (D == 2 && Q == 5) || (D == 2 && Q == 9))) return bad(-1, "Unsupported model"); // m_canPoisson / solverExe.h:37-90
  S.lat.D = D;
  S.lat.Q = Q;
  for(int i = 0; i < Q; ++i) {
    S.lat.w[i]  = LR.w[i];
    S.lat.pw[i] = i == QM ? 0.0 : (Q == 3 ? 0.5 : (Q == 5 ? 0.25 : 1.0 / 8.0)); // constants.h:265,286,306
  }
  S.lat.inv_1mw = 1.0 / (1.0 - LR.w[QM]);
  S.omega = omega;
  S.om1   = 1 - omega;
  // solver.cpp:606: diffusivity = m_poissonAlpha * pow(m_latticeVelocity = 1, 2) * (0.5 - m_relaxTime) * m_dt, m_relaxTime = 1 / omega
  const double alpha       = (D == 2 && Q == 5) ? 1.0 / 2.0 : 1.0 / 3.0;
  const double relax_time  = 1.0 / omega;
  const double diffusivity = alpha * 1.0 * (0.5 - relax_time) * in.poisson_dt;
  S.dt_diff = in.poisson_dt * diffusivity;
  S.rate2   = in.poisson_rate * in.poisson_rate;
  // neighbour in any of the grid's directions: the lattice's own columns, or (D2Q5 corners) the grid table's diagonal columns
  const int NW = in.nghbr_wide.empty() ? QM : 8;
  auto NB = [&](int64_t c, int j) -> int64_t {
    if(j < QM) return in.nghbr[static_cast<size_t>(c) * QM + j];
    return j < NW ? in.nghbr_wide[static_cast<size_t>(c) * 8 + j] : -1;
  };
  S.pull.assign(static_cast<size_t>(N) * QM, -1);
  for(int64_t c = 0; c < N; ++c)
    for(int j = 0; j < QM; ++j) {
      const int64_t t = NB(c, j);
      if(t >= 0) S.pull[static_cast<size_t>(t) * QM + j] = static_cast<int32_t>(c); // the highest source wins
    }
  S.vars0.assign(static_cast<size_t>(N), 0.0);
  static const int opp8[8] = {1, 0, 3, 2, 6, 7, 4, 5}; // cartesian::oppositeDir incl. the 2D diagonals
  for(const BcInput& bc : in.bcs) {
    if(bc.kind != BC_POISSON_DIRICHLET && bc.kind != BC_POISSON_NEUMANN)
      return bad(-1, "this boundary condition does not exist for the Poisson equation types");
    const int64_t n = static_cast<int64_t>(bc.cells.size());
    HostBc b;
    b.neumann = bc.kind == BC_POISSON_NEUMANN;
    b.cells   = bc.cells;
    b.values  = bc.values;
    b.grad    = bc.grad;
    b.ext.assign(static_cast<size_t>(n), -1);
    b.ext2.assign(static_cast<size_t>(n), -1);
    std::vector<char> is_cell(static_cast<size_t>(N), 0), is_ext(static_cast<size_t>(N), 0);
    for(int64_t k = 0; k < n; ++k) {
      // LBMBnd_DirichletNEEM constructor, bnd_dirichlet.h:287-317: opposite of the first missing axis neighbour, the diagonal
      // neighbour at a 2D corner
      const int64_t c = bc.cells[k];
      int ed = -1;
      for(int dist = 0; dist < 2 * D; ++dist) {
        if(NB(c, dist) != -1) continue;
        if(ed < 0) ed = dist;
        else {
          if(ed == 0 && dist == 2) ed = 6;
          if(ed == 0 && dist == 3) ed = 7;
          if(ed == 1 && dist == 3) ed = 4;
          if(ed == 1 && dist == 2) ed = 5;
        }
      }
      if(ed < 0) return bad(-1, "No valid extrapolation cellId");
      const int edir = opp8[ed];
      if(edir >= NW) return bad(-1, "No valid extrapolation cellId (corner: pass the grid's 8-column table, stride >= 8)");
      b.ext[k] = NB(c, edir);
      if(b.ext[k] < 0) return bad(-1, "No valid extrapolation cellId");
      if(b.neumann) {
        b.ext2[k] = NB(b.ext[k], edir);
        if(b.ext2[k] < 0) return bad(-1, "Neumann boundary: no second extrapolation cell");
      }
      if(is_cell[c]) return bad(-5, "Poisson boundary: a cell is listed twice in one surface (order-dependent in the reference)");
      is_cell[c]       = 1;
      is_ext[b.ext[k]] = 1;
    }
    for(int64_t k = 0; k < n; ++k) {
      if(is_cell[b.ext[k]]) return bad(-5, "Poisson boundary: extrapolation cell lies on the same surface (order-dependent in the reference)");
      if(b.neumann && is_ext[b.ext2[k]])
        return bad(-5, "Neumann boundary: a second extrapolation cell is another entry's first one (order-dependent in the reference)");
    }
    if(!b.neumann) // initCnd, bnd_dirichlet.h:335-341 (the Neumann condition has none, bnd_neumann.h:41)
      for(int64_t k = 0; k < n; ++k) S.vars0[bc.cells[k]] = bc.values[k];
    S.bcs.push_back(std::move(b));
  }
  return true;
}

#ifndef LBM_POISSON_HOST_HARNESS
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
#endif

} // namespace poisson
} // namespace lbm
