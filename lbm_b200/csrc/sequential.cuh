// sequential.cuh -- reference-order GPU pipeline for boundary conditions whose result depends on the application order.
//
// The wet-node wall family of the reference (equilibrium, NEEM, NEBB: /root/reference/src/lbm/bnd/bnd_wall.h:101-479,
// bnd_dirichlet.h:134-248, bnd_wetnode.h:12-72, moments.h:120-154) rewrites whole cells from the post-streaming state and from
// m_vars as earlier boundary conditions of the same pass left them (e.g. a corner cell keeps the density the pressure BC
// wrote a moment before; the pressure BC extrapolates from the velocity a wall BC just imposed on its neighbours).  The fused
// kernel resolves "last writer wins" per population slot once, on the host, which cannot express such chains.  Configurations
// that contain a wet-node wall therefore run here: the reference's eight passes as CUDA kernels in the reference's order
// (src/lbm/solver.cpp:307-320), explicit m_f / m_fold / m_feq / m_vars / m_varsold in the reference's layout, boundary
// conditions as phase kernels in LBMBndManager order (bnd.h:48-65).  Every phase is data-parallel where the reference's
// serial loop is order-independent and falls back to one thread per cell where it is not (duplicate surface entries).
// fp64, strict arithmetic: bit-identical to the reference.  These are the reference's own small 2D cases; the bandwidth-
// optimised path is the fused kernel (kernels.cuh).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "kernels.cuh"
#include "plan.hpp"

namespace lbm {
namespace seq {

struct BcDev {
  int            kind = 0;
  int64_t        n = 0;
  const int64_t* cells = nullptr;    // [n]
  const double*  normals = nullptr;  // [n*D]
  double         value[3] = {0, 0, 0};
  double         pressure = 0;
  int            has_pressure = 0, has_velocity = 0;
  const double*  wallval = nullptr;  // tangential BB: [n*Q]
  const int64_t* link = nullptr;     // periodic: [n*Q]
  const int*     linkdist = nullptr; // [n*Q]
  const int*     nset = nullptr;     // [n]
  const int*     lim_n = nullptr;    // wet node: [n]
  const int*     lim_dist = nullptr; // [n*Q]
  const double*  lim_const = nullptr;
  const int64_t* cell2bnd = nullptr; // [n]
  const int64_t* ext = nullptr;      // NEEM
  const int64_t* first_of_cell = nullptr; // [n]: 1 if this entry is the first entry of its cell in the list
  const int64_t* n1 = nullptr;       // pressure: inward neighbours
  const int64_t* n2 = nullptr;
};

struct State {
  double *f, *fold, *feq, *vars, *varsold;
  const int32_t* pull;  // [n*(Q-1)] inverse of the push table
  const int64_t* nghbr; // [n*stride] push table
  int            stride;
  int64_t        n;
  double         omega, om1, omega_minus;
  double         rates[27];
};

template <class L>
__device__ __forceinline__ bool dev_in_direction(const double* nrm, int dist) {
  double dot = 0;
#pragma unroll
  for(int d = 0; d < L::D; ++d) dot += nrm[d] * static_cast<double>(L::c(dist, d));
  return dot >= 2.220446049250313e-16;
}

// passes 2-4: moments, equilibrium, collision (per cell, no ordering issue)
template <class L, int COLL>
__global__ void k_cell(State s) {
  using P = Phys<L, double, true>;
  constexpr int Q = L::Q, D = L::D, NV = D + 1;
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(c >= s.n) return;
  double fo[Q], fe[Q], f[Q], rho, u[D];
#pragma unroll
  for(int i = 0; i < Q; ++i) fo[i] = s.fold[c * Q + i];
  P::moments(fo, rho, u);
  P::equilibrium(rho, u, fe);
  DevParams<double> p{};
  p.omega = s.omega;
  p.om1 = s.om1;
  p.omega_minus = s.omega_minus;
#pragma unroll
  for(int i = 0; i < 27; ++i) p.rates[i] = s.rates[i];
  P::template collide<COLL>(p, fo, fe, f);
#pragma unroll
  for(int i = 0; i < Q; ++i) {
    s.feq[c * Q + i] = fe[i];
    s.f[c * Q + i]   = f[i];
  }
#pragma unroll
  for(int d = 0; d < D; ++d) s.vars[c * NV + d] = u[d];
  s.vars[c * NV + D] = rho;
}

// updateMacroscopicValues only (output(), solver.cpp:336)
template <class L>
__global__ void k_moments(State s, double* out) {
  using P = Phys<L, double, true>;
  constexpr int Q = L::Q, D = L::D, NV = D + 1;
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(c >= s.n) return;
  double fo[Q], rho, u[D];
#pragma unroll
  for(int i = 0; i < Q; ++i) fo[i] = s.fold[c * Q + i];
  P::moments(fo, rho, u);
#pragma unroll
  for(int d = 0; d < D; ++d) out[c * NV + d] = u[d];
  out[c * NV + D] = rho;
}

// pass 7 as a pull: fold[t,j] = f[source(t,j), j] where a source exists, rest population copied in place
template <class L>
__global__ void k_stream(State s) {
  constexpr int Q = L::Q;
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(t >= s.n) return;
#pragma unroll
  for(int j = 0; j < Q - 1; ++j) {
    const int32_t src = s.pull[t * (Q - 1) + j];
    if(src >= 0) s.fold[t * Q + j] = s.f[static_cast<int64_t>(src) * Q + j];
  }
  s.fold[t * Q + Q - 1] = s.f[t * Q + Q - 1];
}

// forcing(), solver.cpp:651-693 (pairs resolved on the host)
template <class L>
__global__ void k_forcing(State s, const ForceEntry* ent, int n) {
  using P = Phys<L, double, true>;
  using A = Ar<double, true>;
  constexpr int Q = L::Q, D = L::D, NV = D + 1;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n) return;
  const ForceEntry e = ent[k];
  double u[D];
#pragma unroll
  for(int d = 0; d < D; ++d) u[d] = s.vars[static_cast<int64_t>(e.val) * NV + d];
  const double vs = P::vsq(u);
#pragma unroll
  for(int i = 0; i < Q; ++i) {
    const double cuv = A::mul(u[0], static_cast<double>(L::c(i, 0)));
    s.f[static_cast<int64_t>(e.target) * Q + i] =
        A::sub(A::add(P::eq_one(L::w(i), e.p, cuv, vs), s.f[static_cast<int64_t>(e.val) * Q + i]), s.feq[static_cast<int64_t>(e.val) * Q + i]);
  }
}

// preApply of one boundary condition (bnd_pressure.h:43-50, bnd_periodic.h:100-118)
template <class L>
__global__ void k_pre_apply(State s, BcDev b) {
  using P = Phys<L, double, true>;
  using A = Ar<double, true>;
  constexpr int Q = L::Q, D = L::D, NV = D + 1;
  const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(k >= b.n) return;
  const int64_t c = b.cells[k];
  if(b.kind == BC_PRESSURE) {
    s.vars[c * NV + D] = b.pressure;
  } else if(b.kind == BC_PERIODIC) {
    const int64_t l0 = b.link[k * Q];
    if(b.has_pressure) {
      if(b.nset[k] < 0) return; // a later entry rewrites the same linked cell: the last writer wins (resolved on the host)
      double u[D];
#pragma unroll
      for(int d = 0; d < D; ++d) u[d] = s.vars[c * NV + d];
      const double vs = P::vsq(u);
      for(int i = 0; i < Q; ++i) {
        const double cuv = P::cu_rt(i, u);
        s.fold[l0 * Q + i] = A::sub(A::add(P::eq_one(L::w(i), b.pressure, cuv, vs), s.f[c * Q + i]), s.feq[c * Q + i]);
      }
      s.vars[l0 * NV + D] = b.pressure;
    } else {
      for(int id = 0; id < b.nset[k]; ++id) {
        const int dist = b.linkdist[k * Q + id];
        if(dist < 0) continue; // overwritten by a later entry (resolved on the host)
        s.fold[b.link[k * Q + id] * Q + dist] = s.f[c * Q + dist];
      }
      s.vars[l0 * NV + D] = 1.0;
    }
  }
}

// moments.h:120-154
template <class L>
__device__ __forceinline__ void density_limited(const State& s, const BcDev& b, int64_t c, int64_t idx, bool noslip) {
  using A = Ar<double, true>;
  constexpr int Q = L::Q, D = L::D, NV = D + 1;
  const int m = b.lim_n[idx];
  if(m == 0) return;
  double rho = 0;
  for(int q = 0; q < m; ++q) {
    const int dist = b.lim_dist[idx * Q + q];
    rho = A::add(rho, A::mul(b.lim_const[idx * Q + dist], s.fold[c * Q + dist]));
  }
  if(!noslip) {
    for(int d = 0; d < D; ++d) {
      const double nd = b.normals[idx * D + d];
      if(nd > 2.220446049250313e-16) rho = A::mul(rho, A::div(1.0, A::add(1.0, s.vars[c * NV + d])));
      else if(nd < 0) rho = A::mul(rho, A::div(1.0, A::sub(1.0, s.vars[c * NV + d])));
    }
  }
  s.vars[c * NV + D] = rho;
}

// LBMBnd_DirichletEQ::apply<VALZERO>, bnd_dirichlet.h:211-241, for entry k
template <class L>
__device__ __forceinline__ void wall_eq_entry(const State& s, const BcDev& b, int64_t k) {
  using P = Phys<L, double, true>;
  constexpr int Q = L::Q, D = L::D, NV = D + 1;
  const int64_t c = b.cells[k], idx = b.cell2bnd[k];
  const bool valzero = !b.has_velocity;
#pragma unroll
  for(int d = 0; d < D; ++d) s.vars[c * NV + d] = b.value[d];
  density_limited<L>(s, b, c, idx, valzero);
  const double rho = s.vars[c * NV + D];
  if(valzero) {
    for(int i = 0; i < Q; ++i) s.fold[c * Q + i] = P::eq_one(L::w(i), rho, 0.0, 0.0);
  } else {
    double u[D], fe[Q];
#pragma unroll
    for(int d = 0; d < D; ++d) u[d] = s.vars[c * NV + d];
    // eq::defaultEq<LBTYPE>(feq, rho, u): cu accumulated from 0 over all dimensions (equilibrium_func.h:68-84)
    const double vs = P::vsq(u);
    for(int i = 0; i < Q; ++i) fe[i] = P::eq_one(L::w(i), rho, P::cu_rt(i, u), vs);
    for(int i = 0; i < Q; ++i) s.fold[c * Q + i] = fe[i];
  }
}

// apply of one boundary condition, phase `phase` (bnd.h:60-65). One thread per entry; for the wet-node kinds one thread per
// CELL that walks the cell's entries in list order (a duplicated entry reads what the first one wrote).
template <class L>
__global__ void k_apply(State s, BcDev b, int phase) {
  using P = Phys<L, double, true>;
  using A = Ar<double, true>;
  constexpr int Q = L::Q, D = L::D, NV = D + 1;
  const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(k >= b.n) return;
  const int64_t c   = b.cells[k];
  const double* nrm = &b.normals[k * D];
  switch(b.kind) {
    case BC_WALL_BB:
    case BC_WALL_BB_TANGENTIAL:
    case BC_DIRICHLET_BB: { // bnd_dirichlet.h:79-121
      for(int i = 0; i < Q - 1; ++i) {
        if(s.nghbr[c * s.stride + i] != -1 || !dev_in_direction<L>(nrm, i)) continue;
        const int op = L::opp(i);
        double    v  = s.f[c * Q + i];
        if(b.kind == BC_WALL_BB_TANGENTIAL) {
          v = A::add(v, A::mul(A::mul(A::div(A::mul(1.0, 2.0), 1.0 / 3.0), L::w(op)), b.wallval[k * Q + op]));
        } else if(b.kind == BC_DIRICHLET_BB) {
          for(int d = 0; d < D; ++d)
            v = A::add(v, A::mul(A::mul(A::mul(A::div(A::mul(1.0, 2.0), 1.0 / 3.0), L::w(op)), static_cast<double>(L::c(op, d))), b.value[d]));
        }
        s.fold[c * Q + op] = v;
      }
      break;
    }
    case BC_PRESSURE: { // bnd_pressure.h:55-106
      const int64_t n1 = b.n1[k], n2 = b.n2[k];
      double ue[D];
#pragma unroll
      for(int d = 0; d < D; ++d) ue[d] = A::sub(A::mul(1.5, s.vars[n1 * NV + d]), A::mul(0.5, s.vars[n2 * NV + d]));
      if(phase == 0) return; // phase 0 only exists to order reads before writes (see launch_apply)
      s.vars[c * NV + D] = b.pressure;
#pragma unroll
      for(int d = 0; d < D; ++d) s.vars[c * NV + d] = ue[d];
      const double vs = P::vsq(ue);
      for(int i = 0; i < Q - 1; ++i) {
        if(s.nghbr[c * s.stride + i] != -1 || !dev_in_direction<L>(nrm, i)) continue;
        const int    op = L::opp(i);
        const double se = P::symm_eq_one(L::w(i), b.pressure, P::cu_rt(i, ue), vs);
        s.fold[c * Q + op] = A::add(-s.f[c * Q + i], A::mul(2.0, se));
      }
      break;
    }
    case BC_WALL_EQ:
    case BC_WALL_NEEM: {
      if(phase == 0) { // equilibrium part, cell by cell
        if(!b.first_of_cell[k]) return;
        for(int64_t j = k; j < b.n; ++j)
          if(b.cells[j] == c) wall_eq_entry<L>(s, b, j);
      } else if(phase == 1) { // calcDensity of the extrapolation cells (moments.h:68-76)
        const int64_t e = b.ext[k];
        double rho = s.fold[e * Q];
        for(int i = 1; i < Q; ++i) rho = A::add(rho, s.fold[e * Q + i]);
        s.vars[e * NV + D] = rho;
      } else if(phase == 2) { // calcVelocity (moments.h:17-35)
        const int64_t e = b.ext[k];
        for(int d = 0; d < D; ++d) {
          double v = 0;
          for(int i = 0; i < Q - 1; ++i) v = A::add(v, A::mul(static_cast<double>(L::c(i, d)), s.fold[e * Q + i]));
          s.vars[e * NV + d] = A::div(v, s.vars[e * NV + D]);
        }
      } else { // add the non-equilibrium part of the extrapolation cell (bnd_wall.h:282-296), cell by cell for duplicates
        if(!b.first_of_cell[k]) return;
        for(int64_t j = k; j < b.n; ++j) {
          if(b.cells[j] != c) continue;
          const int64_t e = b.ext[j];
          double u[D];
#pragma unroll
          for(int d = 0; d < D; ++d) u[d] = s.vars[e * NV + d];
          const double rho = s.vars[e * NV + D], vs = P::vsq(u);
          for(int i = 0; i < Q; ++i)
            s.fold[c * Q + i] = A::add(s.fold[c * Q + i], A::sub(s.fold[e * Q + i], P::eq_one(L::w(i), rho, P::cu_rt(i, u), vs)));
        }
      }
      break;
    }
    default: break;
  }
}

// LBMBnd_wallNEBB (D2Q9 only), bnd_wall.h:366-466; one thread per cell (first entry), phases as in the reference's loops
__global__ void k_nebb(State s, BcDev b, int phase) {
  using L = Lattice<2, 9>;
  using A = Ar<double, true>;
  constexpr int Q = 9, D = 2, NV = 3;
  const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(k >= b.n) return;
  const int64_t c = b.cells[k];
  double*       f = &s.fold[c * Q];
  const double* wv = b.value;
  if(!b.has_velocity) {
    if(phase == 0) {
      s.vars[c * NV + 0] = 0;
      s.vars[c * NV + 1] = 0;
    } else if(phase == 1) {
      if(!b.first_of_cell[k]) return;
      for(int64_t j = k; j < b.n; ++j)
        if(b.cells[j] == c) density_limited<L>(s, b, c, j, true);
    } else if(phase == 2) {
      if(!b.first_of_cell[k]) return;
      for(int dist = 0; dist < 4; ++dist)
        if(s.nghbr[c * s.stride + dist] == -1) f[dist ^ 1] = f[dist];
    } else {
      if(!b.first_of_cell[k]) return;
      const double* nrm = &b.normals[0]; // the reference never advances its entry index in this loop (bnd_wall.h:393-414)
      for(int64_t j = k; j < b.n; ++j) {
        if(b.cells[j] != c) continue;
        if(nrm[0] < 0) {
          f[6] = A::add(f[4], A::mul(0.5, A::sub(f[3], f[2])));
          f[7] = A::sub(f[5], A::mul(0.5, A::sub(f[3], f[2])));
        } else if(nrm[0] > 0) {
          f[4] = A::sub(f[6], A::mul(0.5, A::sub(f[3], f[2])));
          f[5] = A::add(f[7], A::mul(0.5, A::sub(f[3], f[2])));
        } else if(nrm[1] < 0) {
          f[5] = A::sub(f[7], A::mul(0.5, A::sub(f[1], f[0])));
          f[6] = A::add(f[4], A::mul(0.5, A::sub(f[1], f[0])));
        } else if(nrm[1] > 0) {
          f[7] = A::add(f[5], A::mul(0.5, A::sub(f[1], f[0])));
          f[4] = A::sub(f[6], A::mul(0.5, A::sub(f[1], f[0])));
        }
      }
    }
  } else {
    if(!b.first_of_cell[k]) return;
    if(phase == 0) {
      for(int64_t j = k; j < b.n; ++j) {
        if(b.cells[j] != c) continue;
        density_limited<L>(s, b, c, j, false);
        s.vars[c * NV + 0] = wv[0];
        s.vars[c * NV + 1] = wv[1];
      }
    } else if(phase == 1) {
      for(int64_t j = k; j < b.n; ++j) {
        if(b.cells[j] != c) continue;
        const double* nrm = &b.normals[j * D];
        for(int dist = 0; dist < 4; ++dist) {
          const int dir = dist / 2;
          if(s.nghbr[c * s.stride + dist] == -1 && fabs(nrm[dir]) > 0) {
            double v = f[dist];
            const double t = A::mul(A::mul(2.0 / 3.0, s.vars[c * NV + D]), wv[dir]);
            v = nrm[dir] < 0 ? A::sub(v, t) : A::add(v, t);
            f[dist ^ 1] = v;
          }
        }
      }
    } else if(phase == 2) {
      for(int64_t j = k; j < b.n; ++j) {
        if(b.cells[j] != c) continue;
        const double* nrm = &b.normals[j * D];
        const double  rho = s.vars[c * NV + D];
        // a +- 0.5*(fa - fb) -+ 0.5*rho*w_t -+ 1/6*rho*w_n, evaluated left to right like the reference
        auto t_half = [&](double wt) { return A::mul(A::mul(0.5, rho), wt); };
        auto t_sixth = [&](double wn) { return A::mul(A::mul(1.0 / 6.0, rho), wn); };
        if(nrm[0] > 0) {
          const double h = A::mul(0.5, A::sub(f[3], f[2]));
          f[6] = A::sub(A::sub(A::add(f[4], h), t_half(wv[1])), t_sixth(wv[0]));
          const double h2 = A::mul(0.5, A::sub(f[3], f[2]));
          f[7] = A::sub(A::add(A::sub(f[5], h2), t_half(wv[1])), t_sixth(wv[0]));
        } else if(nrm[0] < 0) {
          const double h = A::mul(0.5, A::sub(f[3], f[2]));
          f[4] = A::add(A::add(A::sub(f[6], h), t_half(wv[1])), t_sixth(wv[0]));
          const double h2 = A::mul(0.5, A::sub(f[3], f[2]));
          f[5] = A::add(A::sub(A::add(f[7], h2), t_half(wv[1])), t_sixth(wv[0]));
        } else if(nrm[1] > 0) {
          const double h = A::mul(0.5, A::sub(f[1], f[0]));
          f[5] = A::sub(A::add(A::sub(f[7], h), t_half(wv[0])), t_sixth(wv[1]));
          const double h2 = A::mul(0.5, A::sub(f[1], f[0]));
          f[6] = A::sub(A::sub(A::add(f[4], h2), t_half(wv[0])), t_sixth(wv[1]));
        } else if(nrm[1] < 0) {
          const double h = A::mul(0.5, A::sub(f[1], f[0]));
          f[7] = A::add(A::sub(A::add(f[5], h), t_half(wv[0])), t_sixth(wv[1]));
          const double h2 = A::mul(0.5, A::sub(f[1], f[0]));
          f[4] = A::add(A::add(A::sub(f[6], h2), t_half(wv[0])), t_sixth(wv[1]));
        }
      }
    }
  }
}

// sum_c |vars - varsold| per variable, AoS (solver.cpp:809-815); same fixed-shape two-pass reduction as k_residual
__global__ void k_residual_aos(const double* __restrict__ v, const double* __restrict__ vo, int64_t n, int nvar, double* __restrict__ partial) {
  __shared__ double sh[32];
  for(int var = 0; var < nvar; ++var) {
    double acc = 0;
    for(int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; c < n; c += static_cast<int64_t>(gridDim.x) * blockDim.x)
      acc += fabs(v[c * nvar + var] - vo[c * nvar + var]);
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if(threadIdx.x < 32) {
      double t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
      for(int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if(threadIdx.x == 0) partial[static_cast<size_t>(var) * gridDim.x + blockIdx.x] = t;
    }
    __syncthreads();
  }
}

// initialCondition(), solver.cpp:267-295
template <class L>
__global__ void k_init(State s) {
  using P = Phys<L, double, true>;
  constexpr int Q = L::Q, D = L::D, NV = D + 1;
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(c >= s.n) return;
  double u[D], fe[Q];
#pragma unroll
  for(int d = 0; d < D; ++d) u[d] = s.vars[c * NV + d];
  s.vars[c * NV + D] = 1.0;
  P::equilibrium(1.0, u, fe);
#pragma unroll
  for(int i = 0; i < Q; ++i) {
    s.feq[c * Q + i]  = fe[i];
    s.f[c * Q + i]    = fe[i];
    s.fold[c * Q + i] = fe[i];
  }
}

} // namespace seq
} // namespace lbm
