// kernels.cuh -- sm_100a device code of the fused lattice-Boltzmann step.
//
// One launch per time step does what the reference does in eight passes over array-of-structures data
// (/root/reference/src/lbm/solver.cpp:307-320): pull the populations that passes 6-8 of the previous reference
// step (preApply, push-propagation, apply; :699-755) left in m_fold, then moments (:513-553), equilibrium
// (:556-571, equilibrium_func.h:52-84) and collision (:601-613), and store the post-collision populations.
// Populations are SoA, fp64 (or fp32), double buffered (A -> B).
//
// Bandwidth design (B200: HBM3e bound, no tensor cores):
//   * every population is read once and written once per step: 2*Q*sizeof(real) bytes per cell;
//   * "fast" SFC chunks (8^3 / 32^2 cells, plan.hpp) need no per-cell index: the pull offsets inside a chunk are a
//     property of the curve, held once in shared memory (the template) plus 3^D neighbour-chunk bases per chunk;
//     a persistent CTA walks chunks in curve order so halo reads hit L2;
//   * stores are fully coalesced (consecutive threads, consecutive cells); loads are gathers inside a 4 KB
//     window per direction that L1 absorbs;
//   * cells next to boundaries go through 32-bit link codes (generic path), O(surface) of the domain.
//
// Arithmetic policies: STRICT reproduces the reference's operation order with IEEE intrinsics that the compiler
// never contracts to FMA (bit-identical to the reference in fp64); FAST lets the compiler contract and replaces the
// divisions by constants with multiplications.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "lattice.h"

namespace lbm {

enum { COLL_BGK = 0, COLL_TRT = 1, COLL_MRT = 2 };

template <class Real>
struct AddEntryT { Real v[3]; int32_t n; };
template <class Real>
struct AbbDev { int32_t cell, n1, n2; Real p; };
template <class Real>
struct ForceDev { int32_t target, val; Real p; };
template <class Real>
struct PerPDev { int32_t cell, vbase; Real p; };
template <class Real>
struct VarFixDev { int32_t cell, var, abb, comp; Real value; };
struct CopySrcDev { int32_t cell, dir; };

// what a non-pull slot needs; passed BY VALUE to the out-of-line slot evaluator so that the kernel parameter
// block never has to be copied to local memory
template <class Real>
struct SlotTables {
  const CopySrcDev*      copytab;
  const AddEntryT<Real>* addtab;
  const AbbDev<Real>*    abb;
  const Real*            uext;   // [n_abb][3]  extrapolated velocity of every pressure entry (current)
  const Real*            values; // static + dynamic slot values (current)
  int64_t                stride;
};

template <class Real>
struct DevParams {
  const Real* A;   // populations before the step (post-collision of the previous step), SoA [Q][stride]
  Real*       B;   // populations after the step
  int64_t     stride;
  // fast chunks
  const uint16_t* tmpl;     // [(Q-1)][CHUNK]
  const int32_t*  chunk_nb; // [n_fast_chunks][NSEL + 1]: neighbour chunk bases (-1 = wall), wall descriptor id
  const AddEntryT<Real>* wall_desc; // [n_wall_desc][NSEL][Q-1] bounce-back addends of wall chunks per (missing neighbour chunk,
                                    // direction); n < 0: anti-bounce-back slot
  const int32_t*  chunk_abb_base;   // [n_fast_chunks] row of chunk_abb for chunks on a pressure face, -1 otherwise
  const int32_t*  chunk_abb;        // [rows][CHUNK] pressure entry of the cell at that offset
  int32_t         n_fast_chunks;    // fast chunks [chunk_off, chunk_off + n_fast_chunks) are updated by this launch
  int32_t         chunk_off;
  int32_t         n_fast_blocks;
  unsigned long long* ticket;      // global chunk ticket counter of this launch class
  unsigned long long  ticket_base; // value of the counter when this launch starts
  // generic range
  int32_t        gen_begin, n_gen, n_gen_blocks; // generic cells [gen_off, gen_off + n_gen) of the generic range
  int32_t        gen_off;
  int64_t        gen_stride;
  const int32_t* codes; // [(Q-1)][gen_stride]
  SlotTables<Real> tabs;
  // collision
  Real omega, om1, omega_minus;
  Real rates[27];
  // options
  Real*   vars_out; // SoA [NVAR][stride] or nullptr
  int32_t first;    // 1: step 0 -- m_fold is the initial condition itself, not a streamed state
};

// ------------------------------------------------------------------------------------------- arithmetic
template <class Real, bool STRICT>
struct Ar;
template <>
struct Ar<double, true> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <>
struct Ar<float, true> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <class Real>
struct Ar<Real, false> {
  static __device__ __forceinline__ Real add(Real a, Real b) { return a + b; }
  static __device__ __forceinline__ Real sub(Real a, Real b) { return a - b; }
  static __device__ __forceinline__ Real mul(Real a, Real b) { return a * b; }
  static __device__ __forceinline__ Real div(Real a, Real b) { return a / b; }
};

// c * x for c in {-1,0,1} is exact; adding an exact zero never changes a sum, so zero terms are skipped.
template <class L, class Real, bool STRICT>
struct Phys {
  using A = Ar<Real, STRICT>;
  static constexpr int D = L::D, Q = L::Q;

  // lbm_cssq = 1.0/3.0 as a double (constants.h:27); the derived constants are rounded like the reference's
  static __device__ __forceinline__ Real cssq() { return static_cast<Real>(1.0 / 3.0); }
  static __device__ __forceinline__ Real c2() { return static_cast<Real>(2.0 * (1.0 / 3.0) * (1.0 / 3.0)); }
  static __device__ __forceinline__ Real c3() { return static_cast<Real>(2.0 * (1.0 / 3.0)); }

  // solver.cpp:527-535: rho = sum ascending from 0.0; u_d = (sum_{i<Q-1} c_id f_i) / rho
  static __device__ __forceinline__ void moments(const Real (&f)[Q], Real& rho, Real (&u)[D]) {
    Real r = f[0];
#pragma unroll
    for(int i = 1; i < Q; ++i) r = A::add(r, f[i]);
    rho = r;
#pragma unroll
    for(int d = 0; d < D; ++d) {
      Real v     = 0;
      bool first = true;
#pragma unroll
      for(int i = 0; i < Q - 1; ++i) {
        if(L::c(i, d) == 0) continue;
        const Real t = L::c(i, d) > 0 ? f[i] : -f[i];
        v     = first ? t : A::add(v, t);
        first = false;
      }
      u[d] = A::div(v, r);
    }
  }

  static __device__ __forceinline__ Real vsq(const Real (&u)[D]) {
    Real v = A::mul(u[0], u[0]);
#pragma unroll
    for(int d = 1; d < D; ++d) v = A::add(v, A::mul(u[d], u[d]));
    return v;
  }

  template <int I>
  static __device__ __forceinline__ Real cu(const Real (&u)[D]) {
    Real v     = 0;
    bool first = true;
#pragma unroll
    for(int d = 0; d < D; ++d) {
      if(L::c(I, d) == 0) continue;
      const Real t = L::c(I, d) > 0 ? u[d] : -u[d];
      v     = first ? t : A::add(v, t);
      first = false;
    }
    return v;
  }
  static __device__ __forceinline__ Real cu_rt(int i, const Real (&u)[D]) {
    Real v = 0;
#pragma unroll
    for(int d = 0; d < D; ++d) v = A::add(v, A::mul(u[d], static_cast<Real>(L::c(i, d))));
    return v;
  }

  // equilibrium_func.h:52-54: w*rho*(1.0 + cu/cssq + cu*cu/(2.0*cssq*cssq) - vsq/(2.0*cssq))
  static __device__ __forceinline__ Real eq_one(Real w, Real rho, Real cuv, Real vs) {
    if constexpr(STRICT) {
      const Real t = A::sub(A::add(A::add(Real(1), A::div(cuv, cssq())), A::div(A::mul(cuv, cuv), c2())), A::div(vs, c3()));
      return A::mul(A::mul(w, rho), t);
    } else {
      return w * rho * (Real(1) + Real(3) * cuv + Real(4.5) * cuv * cuv - Real(1.5) * vs);
    }
  }
  // equilibrium_func.h:109-111
  static __device__ __forceinline__ Real symm_eq_one(Real w, Real rho, Real cuv, Real vs) {
    if constexpr(STRICT) {
      const Real t = A::sub(A::add(Real(1), A::div(A::mul(cuv, cuv), c2())), A::div(vs, c3()));
      return A::mul(A::mul(w, rho), t);
    } else {
      return w * rho * (Real(1) + Real(4.5) * cuv * cuv - Real(1.5) * vs);
    }
  }

  // all Q equilibria. Opposite directions share |cu|: x/c and (x*x)/c are sign-symmetric in IEEE arithmetic, so the
  // pair is computed from one set of divisions and is still bit-identical to the reference's per-direction form.
  static __device__ __forceinline__ void equilibrium(Real rho, const Real (&u)[D], Real (&feq)[Q]) {
    const Real vs = vsq(u);
    if constexpr(STRICT) {
      const Real vterm = A::div(vs, c3());
      static_for_eq<0>(rho, u, vterm, feq);
    } else {
      const Real base = Real(1) - Real(1.5) * vs;
      static_for_eq_fast<0>(rho, u, base, feq);
    }
  }
  template <int I>
  static __device__ __forceinline__ void static_for_eq(Real rho, const Real (&u)[D], Real vterm, Real (&feq)[Q]) {
    if constexpr(I < Q) {
      constexpr int J = L::opp(I);
      if constexpr(J >= I) {
        const Real wr = A::mul(static_cast<Real>(L::w(I)), rho);
        if constexpr(J == I) {
          // rest population: cu = 0 -> 1.0 + 0/cssq + 0/c2 - vterm
          feq[I] = A::mul(wr, A::sub(Real(1), vterm));
        } else {
          const Real cuv = cu<I>(u);
          const Real a   = A::div(cuv, cssq());
          const Real b   = A::div(A::mul(cuv, cuv), c2());
          feq[I]         = A::mul(wr, A::sub(A::add(A::add(Real(1), a), b), vterm));
          feq[J]         = A::mul(wr, A::sub(A::add(A::add(Real(1), -a), b), vterm));
        }
      }
      static_for_eq<I + 1>(rho, u, vterm, feq);
    }
  }
  template <int I>
  static __device__ __forceinline__ void static_for_eq_fast(Real rho, const Real (&u)[D], Real base, Real (&feq)[Q]) {
    if constexpr(I < Q) {
      constexpr int J = L::opp(I);
      if constexpr(J >= I) {
        const Real wr = static_cast<Real>(L::w(I)) * rho;
        if constexpr(J == I) {
          feq[I] = wr * base;
        } else {
          const Real cuv = cu<I>(u);
          const Real e   = base + Real(4.5) * cuv * cuv;
          feq[I]         = wr * (e + Real(3) * cuv);
          feq[J]         = wr * (e - Real(3) * cuv);
        }
      }
      static_for_eq_fast<I + 1>(rho, u, base, feq);
    }
  }

  // collision: BGK is the reference's formula (solver.cpp:603); TRT / MRT are extensions (see oracle/lbm_oracle.c)
  template <int COLL>
  static __device__ __forceinline__ void collide(const DevParams<Real>& p, const Real (&fo)[Q], const Real (&fe)[Q], Real (&f)[Q]) {
    if constexpr(COLL == COLL_BGK) {
#pragma unroll
      for(int i = 0; i < Q; ++i) {
        if constexpr(STRICT) f[i] = A::add(A::mul(p.om1, fo[i]), A::mul(p.omega, fe[i]));
        else f[i] = fo[i] + p.omega * (fe[i] - fo[i]);
      }
    } else {
#pragma unroll
      for(int i = 0; i < Q; ++i) {
        const int j  = L::opp(i);
        Real      wp = p.omega, wm = p.omega_minus;
        if constexpr(COLL == COLL_MRT) {
          wp = p.rates[i < j ? i : j];
          wm = p.rates[i < j ? j : i];
        }
        const Real fp  = A::mul(Real(0.5), A::add(fo[i], fo[j]));
        const Real fm  = A::mul(Real(0.5), A::sub(fo[i], fo[j]));
        const Real fep = A::mul(Real(0.5), A::add(fe[i], fe[j]));
        const Real fem = A::mul(Real(0.5), A::sub(fe[i], fe[j]));
        f[i]           = A::sub(A::sub(fo[i], A::mul(wp, A::sub(fp, fep))), A::mul(wm, A::sub(fm, fem)));
      }
    }
  }
};

// ------------------------------------------------------------------------------------------- gathers
// value of a non-pull slot of device cell `cell`, direction J (generic path)
template <class L, class Real, bool STRICT>
__device__ __noinline__ Real special_slot(const SlotTables<Real> p, const Real* __restrict__ Abuf, int32_t code, int32_t cell, int j) {
  using P = Phys<L, Real, STRICT>;
  using A = Ar<Real, STRICT>;
  const int     kind = link_kind(code);
  const int32_t pl   = link_payload(code);
  const int     oj   = L::opp(j);
  switch(kind) {
    case LK_COPY: {
      const CopySrcDev cs = p.copytab[pl];
      return Abuf[static_cast<size_t>(cs.dir) * p.stride + cs.cell];
    }
    case LK_BB: return Abuf[static_cast<size_t>(oj) * p.stride + cell];
    case LK_BB_ADD: {
      // bnd_dirichlet.h:92,111-117: fold = f; then one += per addend
      Real                  v = Abuf[static_cast<size_t>(oj) * p.stride + cell];
      const AddEntryT<Real> e = p.addtab[pl];
      for(int t = 0; t < e.n; ++t) v = A::add(v, e.v[t]);
      return v;
    }
    case LK_ABB: {
      // bnd_pressure.h:100: fold[c,opp] = -f[c,dist] + 2 * symmEq(dist, p, u_ext)
      const AbbDev<Real> e = p.abb[pl];
      Real               u[L::D];
#pragma unroll
      for(int d = 0; d < L::D; ++d) u[d] = p.uext[static_cast<size_t>(pl) * 3 + d];
      const Real vs  = P::vsq(u);
      const Real cuv = P::cu_rt(oj, u);
      const Real se  = P::symm_eq_one(static_cast<Real>(L::w(oj)), e.p, cuv, vs);
      return A::add(-Abuf[static_cast<size_t>(oj) * p.stride + cell], A::mul(Real(2), se));
    }
    default: return p.values[pl];
  }
}

// m_fold of a cell of the generic range
template <class L, class Real, bool STRICT>
__device__ __forceinline__ void gather_generic(const DevParams<Real>& p, const Real* __restrict__ Abuf, int32_t cell, Real (&fold)[L::Q]) {
  constexpr int Q = L::Q;
  const int64_t g = cell - p.gen_begin;
  if(p.first) {
#pragma unroll
    for(int j = 0; j < Q; ++j) fold[j] = Abuf[static_cast<size_t>(j) * p.stride + cell];
    return;
  }
  int32_t code[Q - 1];
#pragma unroll
  for(int j = 0; j < Q - 1; ++j) code[j] = __ldg(&p.codes[static_cast<size_t>(j) * p.gen_stride + g]);
#pragma unroll
  for(int j = 0; j < Q - 1; ++j) {
    if(code[j] >= 0) fold[j] = Abuf[static_cast<size_t>(j) * p.stride + code[j]];
  }
  fold[Q - 1] = Abuf[static_cast<size_t>(Q - 1) * p.stride + cell];
#pragma unroll
  for(int j = 0; j < Q - 1; ++j) {
    if(code[j] < 0) fold[j] = special_slot<L, Real, STRICT>(p.tabs, Abuf, code[j], cell, j);
  }
}

// m_fold of a cell of a fast chunk, template taken from global memory (used by the small auxiliary kernels)
// one slot of a fast chunk: pull at the template offset, or bounce back (+ addends) when the neighbour chunk is a wall
template <class L, class Real, bool STRICT, int J>
__device__ __forceinline__ Real fast_slot(const DevParams<Real>& p, const Real* __restrict__ Abuf, int32_t cell, int32_t nbv, uint32_t off,
                                          const AddEntryT<Real>* __restrict__ wall_of_chunk, uint32_t sel) {
  using A = Ar<Real, STRICT>;
  if(nbv >= 0) return Abuf[static_cast<size_t>(J) * p.stride + nbv + static_cast<int32_t>(off)];
  // the descriptor of this (missing neighbour chunk, direction): on an edge of the domain the same direction bounces off different walls
  const AddEntryT<Real>* __restrict__ wall = wall_of_chunk + sel * (L::Q - 1);
  const int n = wall[J].n;
  if(n < 0) {
    // chunk on a pressure in-/outlet face: anti-bounce-back with the cell's own pressure entry, evaluated by the same
    // out-of-line routine as on the generic path (bnd_pressure.h:100)
    const int32_t entry = p.chunk_abb[static_cast<size_t>(p.chunk_abb_base[cell / L::CHUNK]) * L::CHUNK + cell % L::CHUNK];
    return special_slot<L, Real, STRICT>(p.tabs, Abuf, link_code(LK_ABB, entry), cell, J);
  }
  Real v = Abuf[static_cast<size_t>(L::opp(J)) * p.stride + cell]; // bnd_dirichlet.h:92
  for(int t = 0; t < n; ++t) v = A::add(v, wall[J].v[t]);            // bnd_dirichlet.h:111-117
  return v;
}

template <class L, class Real, bool STRICT, int J>
__device__ __forceinline__ void gather_fast_global_rec(const DevParams<Real>& p, const Real* __restrict__ Abuf, int32_t cell, int chunk, int o,
                                                       const AddEntryT<Real>* __restrict__ wall, Real (&fold)[L::Q]) {
  if constexpr(J < L::Q - 1) {
    const uint32_t t   = p.tmpl[J * L::CHUNK + o];
    const int32_t  nbv = p.chunk_nb[static_cast<size_t>(chunk) * (L::NSEL + 1) + (t >> 10)];
    fold[J]            = fast_slot<L, Real, STRICT, J>(p, Abuf, cell, nbv, t & 1023u, wall, t >> 10);
    gather_fast_global_rec<L, Real, STRICT, J + 1>(p, Abuf, cell, chunk, o, wall, fold);
  }
}

// m_fold of a cell of a fast chunk, template taken from global memory (used by the small auxiliary kernels)
template <class L, class Real, bool STRICT>
__device__ __forceinline__ void gather_fast_global(const DevParams<Real>& p, const Real* __restrict__ Abuf, int32_t cell, Real (&fold)[L::Q]) {
  constexpr int Q = L::Q, CH = L::CHUNK;
  if(p.first) {
#pragma unroll
    for(int j = 0; j < Q; ++j) fold[j] = Abuf[static_cast<size_t>(j) * p.stride + cell];
    return;
  }
  const int chunk = cell / CH, o = cell % CH;
  const int32_t wid = p.chunk_nb[static_cast<size_t>(chunk) * (L::NSEL + 1) + L::NSEL];
  const AddEntryT<Real>* wall = p.wall_desc + static_cast<size_t>(wid < 0 ? 0 : wid) * (Q - 1) * L::NSEL;
  gather_fast_global_rec<L, Real, STRICT, 0>(p, Abuf, cell, chunk, o, wall, fold);
  fold[Q - 1] = Abuf[static_cast<size_t>(Q - 1) * p.stride + cell];
}

template <class L, class Real, bool STRICT>
__device__ __forceinline__ void gather_any(const DevParams<Real>& p, const Real* __restrict__ Abuf, int32_t cell, Real (&fold)[L::Q]) {
  if(cell < p.gen_begin) gather_fast_global<L, Real, STRICT>(p, Abuf, cell, fold);
  else gather_generic<L, Real, STRICT>(p, Abuf, cell, fold);
}

// ------------------------------------------------------------------------------------------- main kernel
template <class L, class Real, bool STRICT, int COLL>
__device__ __forceinline__ void update_and_store(const DevParams<Real>& p, int32_t cell, const Real (&fold)[L::Q]) {
  using P = Phys<L, Real, STRICT>;
  constexpr int Q = L::Q, D = L::D;
  Real rho, u[D], feq[Q], f[Q];
  P::moments(fold, rho, u);
  P::equilibrium(rho, u, feq);
  P::template collide<COLL>(p, fold, feq, f);
#pragma unroll
  for(int j = 0; j < Q; ++j) {
#if defined(LBM_STORE_CS)
    __stcs(&p.B[static_cast<size_t>(j) * p.stride + cell], f[j]);
#elif defined(LBM_STORE_CG)
    __stcg(&p.B[static_cast<size_t>(j) * p.stride + cell], f[j]);
#else
    p.B[static_cast<size_t>(j) * p.stride + cell] = f[j];
#endif
  }
  if(p.vars_out != nullptr) {
#pragma unroll
    for(int d = 0; d < D; ++d) p.vars_out[static_cast<size_t>(d) * p.stride + cell] = u[d];
    p.vars_out[static_cast<size_t>(D) * p.stride + cell] = rho;
  }
}

#ifndef LBM_THREADS
#define LBM_THREADS 256
#endif
#ifndef LBM_MINBLOCKS
#define LBM_MINBLOCKS 2
#endif
constexpr int kThreads = LBM_THREADS;

template <class L, class Real, bool STRICT, int J>
__device__ __forceinline__ void fast_wall_gather(const DevParams<Real>& p, const Real* __restrict__ Abuf, int32_t cell, int o,
                                                 const uint16_t* __restrict__ s_tmpl, const int32_t* __restrict__ nb,
                                                 const AddEntryT<Real>* __restrict__ wall, Real (&fold)[L::Q]) {
  if constexpr(J < L::Q - 1) {
    const uint32_t t = s_tmpl[J * L::CHUNK + o];
    fold[J]          = fast_slot<L, Real, STRICT, J>(p, Abuf, cell, nb[t >> 10], t & 1023u, wall, t >> 10);
    fast_wall_gather<L, Real, STRICT, J + 1>(p, Abuf, cell, o, s_tmpl, nb, wall, fold);
  }
}

template <class L, class Real, bool STRICT, int COLL>
__global__ void __launch_bounds__(kThreads, LBM_MINBLOCKS) k_step(const __grid_constant__ DevParams<Real> p) {
  constexpr int Q = L::Q, QM = Q - 1, CH = L::CHUNK, NSEL = L::NSEL;
  __shared__ uint16_t s_tmpl[QM * CH];
  __shared__ int32_t  s_nb[2][NSEL + 1];
  __shared__ int32_t  s_ticket[2];
  const Real* __restrict__ Abuf = p.A;

  if(static_cast<int>(blockIdx.x) < p.n_gen_blocks) {
    // ---- generic path: one thread per cell, per-slot link codes. Scheduled first: these blocks are the slow ones.
    const int32_t gl = blockIdx.x * kThreads + threadIdx.x;
    if(gl >= p.n_gen) return;
    const int32_t cell = p.gen_begin + p.gen_off + gl;
    Real          fold[Q];
    gather_generic<L, Real, STRICT>(p, Abuf, cell, fold);
    update_and_store<L, Real, STRICT, COLL>(p, cell, fold);
    return;
  }

  // ---- fast path: persistent CTA over SFC chunks, template in shared memory, no per-cell index traffic
  // Chunks are handed out dynamically, in curve order, from a global ticket counter: a CTA that becomes resident late
  // (another kernel holds its slot) simply takes fewer chunks, and there is no tail of unevenly loaded CTAs.  The counter
  // only ever grows: a launch over n chunks with B CTAs advances it by exactly n + B (every CTA draws one ticket past the
  // end), so the host knows the first ticket of every launch and never has to reset anything.
  for(int t = threadIdx.x; t < QM * CH; t += kThreads) s_tmpl[t] = p.tmpl[t];
  if(threadIdx.x == 0) s_ticket[0] = static_cast<int32_t>(atomicAdd(p.ticket, 1ull) - p.ticket_base);
  __syncthreads();
  int32_t ticket = s_ticket[0];
  int     buf    = 0;
  while(ticket < p.n_fast_chunks) {
    const int chunk = p.chunk_off + ticket;
    // neighbour-chunk bases and the next ticket are double buffered: one barrier per chunk is enough (a thread can run at
    // most one chunk ahead of the slowest one, and then it touches the other buffer)
    if(threadIdx.x < NSEL + 1) s_nb[buf][threadIdx.x] = p.chunk_nb[static_cast<size_t>(chunk) * (NSEL + 1) + threadIdx.x];
    if(threadIdx.x == 0) s_ticket[buf ^ 1] = static_cast<int32_t>(atomicAdd(p.ticket, 1ull) - p.ticket_base);
    __syncthreads();
    ticket = s_ticket[buf ^ 1];
    const int32_t* nb  = s_nb[buf];
    buf ^= 1;
    const int32_t  wid = nb[NSEL]; // wall descriptor of this chunk, -1: interior chunk
    const AddEntryT<Real>* wall = p.wall_desc + static_cast<size_t>(wid < 0 ? 0 : wid) * QM * NSEL;
    const int32_t base = chunk * CH;
#pragma unroll 1
    for(int o = threadIdx.x; o < CH; o += kThreads) {
      const int32_t cell = base + o;
      Real          fold[Q];
      if(p.first) {
#pragma unroll
        for(int j = 0; j < Q; ++j) fold[j] = Abuf[static_cast<size_t>(j) * p.stride + cell];
      } else if(wid < 0) {
        // interior chunk: pure pulls
#pragma unroll
        for(int j = 0; j < QM; ++j) {
          const uint32_t t   = s_tmpl[j * CH + o];
          const int32_t  src = nb[t >> 10] + static_cast<int32_t>(t & 1023u);
#if defined(LBM_LOAD_NC)
          fold[j]            = __ldg(&Abuf[static_cast<size_t>(j) * p.stride + src]);
#elif defined(LBM_LOAD_CG)
          fold[j]            = __ldcg(&Abuf[static_cast<size_t>(j) * p.stride + src]);
#else
          fold[j]            = Abuf[static_cast<size_t>(j) * p.stride + src];
#endif
        }
        fold[QM] = Abuf[static_cast<size_t>(QM) * p.stride + cell];
      } else {
        // wall chunk: slots whose neighbour chunk is a wall bounce back
        fast_wall_gather<L, Real, STRICT, 0>(p, Abuf, cell, o, s_tmpl, nb, wall, fold);
        fold[QM] = Abuf[static_cast<size_t>(QM) * p.stride + cell];
      }
      update_and_store<L, Real, STRICT, COLL>(p, cell, fold);
    }
  }
}

// ------------------------------------------------------------------------------------------- auxiliary kernels
// They run after the main kernel of a step, on the few cells that the reference touches with forcing(),
// preApply() and the pressure boundary condition.  Abuf still holds the populations the step started from, so the
// step's m_fold / m_vars / m_feq of any cell can be rebuilt on the fly.

// forcing(): f[target,i] = defaultEq(w_i, p, u_x(val)*c_ix, |u(val)|^2) + f[val,i] - feq[val,i]   (solver.cpp:664-669)
template <class L, class Real, bool STRICT>
__global__ void k_forcing(const __grid_constant__ DevParams<Real> p, const ForceDev<Real>* __restrict__ ent, int n) {
  using P = Phys<L, Real, STRICT>;
  using A = Ar<Real, STRICT>;
  constexpr int Q = L::Q, D = L::D;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n) return;
  const ForceDev<Real> e = ent[k];
  Real fold[Q], rho, u[D], feq[Q];
  gather_any<L, Real, STRICT>(p, p.A, e.val, fold);
  P::moments(fold, rho, u);
  P::equilibrium(rho, u, feq);
  const Real vs = P::vsq(u);
#pragma unroll
  for(int i = 0; i < Q; ++i) {
    const Real cuv = A::mul(u[0], static_cast<Real>(L::c(i, 0)));
    const Real fv  = p.B[static_cast<size_t>(i) * p.stride + e.val];
    p.B[static_cast<size_t>(i) * p.stride + e.target] = A::sub(A::add(P::eq_one(static_cast<Real>(L::w(i)), e.p, cuv, vs), fv), feq[i]);
  }
}

// periodic boundary with pressure, preApply: value = defaultEq(i, p, u(c)) + f[c,i] - feq[c,i]  (bnd_periodic.h:101-108)
template <class L, class Real, bool STRICT>
__global__ void k_periodic_pressure(const __grid_constant__ DevParams<Real> p, const PerPDev<Real>* __restrict__ ent, int n,
                                    Real* __restrict__ values_next) {
  using P = Phys<L, Real, STRICT>;
  using A = Ar<Real, STRICT>;
  constexpr int Q = L::Q, D = L::D;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n) return;
  const PerPDev<Real> e = ent[k];
  Real fold[Q], rho, u[D], feq[Q];
  gather_any<L, Real, STRICT>(p, p.A, e.cell, fold);
  P::moments(fold, rho, u);
  P::equilibrium(rho, u, feq);
  const Real vs = P::vsq(u);
#pragma unroll
  for(int i = 0; i < Q; ++i) {
    const Real cuv = P::cu_rt(i, u);
    const Real fv  = p.B[static_cast<size_t>(i) * p.stride + e.cell];
    values_next[e.vbase + i] = A::sub(A::add(P::eq_one(static_cast<Real>(L::w(i)), e.p, cuv, vs), fv), feq[i]);
  }
}

// velocity (this step's m_vars) of an inward neighbour of a pressure cell: rebuilt from the populations when this rank owns
// the cell, taken from the received velocity halo when another rank does (n < 0: -(slot + 1), plan.hpp section 8)
template <class L, class Real, bool STRICT>
__device__ __forceinline__ void neighbour_velocity(const DevParams<Real>& p, int32_t n, const Real* __restrict__ vrecv, Real (&u)[L::D]) {
  if(n < 0) {
#pragma unroll
    for(int d = 0; d < L::D; ++d) u[d] = vrecv[static_cast<size_t>(-n - 1) * 3 + d];
    return;
  }
  Real fold[L::Q], rho;
  gather_any<L, Real, STRICT>(p, p.A, n, fold);
  Phys<L, Real, STRICT>::moments(fold, rho, u);
}

// pressure boundary: u_ext = 1.5 u(n1) - 0.5 u(n2) from this step's m_vars  (bnd_pressure.h:78-84)
template <class L, class Real, bool STRICT>
__global__ void k_pressure_extrapolate(const __grid_constant__ DevParams<Real> p, int n, Real* __restrict__ uext_next,
                                       const Real* __restrict__ vrecv) {
  using A = Ar<Real, STRICT>;
  constexpr int D = L::D;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n) return;
  const AbbDev<Real> e = p.tabs.abb[k];
  Real u1[D], u2[D];
  neighbour_velocity<L, Real, STRICT>(p, e.n1, vrecv, u1);
  neighbour_velocity<L, Real, STRICT>(p, e.n2, vrecv, u2);
#pragma unroll
  for(int d = 0; d < D; ++d) uext_next[static_cast<size_t>(k) * 3 + d] = A::sub(A::mul(Real(1.5), u1[d]), A::mul(Real(0.5), u2[d]));
}

// velocity halo, sending side: this step's m_vars velocity of the listed owned cells, 3 reals per item (wire order)
template <class L, class Real, bool STRICT>
__global__ void k_velocity_pack(const __grid_constant__ DevParams<Real> p, const int32_t* __restrict__ cells, int n, Real* __restrict__ out) {
  constexpr int D = L::D;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n) return;
  Real u[D];
  neighbour_velocity<L, Real, STRICT>(p, cells[k], nullptr, u);
#pragma unroll
  for(int d = 0; d < 3; ++d) out[static_cast<size_t>(k) * 3 + d] = d < D ? u[d < D ? d : 0] : Real(0);
}

// what the boundary conditions write into m_vars after the moments pass (residual bookkeeping only)
template <class Real>
__global__ void k_varfix(const VarFixDev<Real>* __restrict__ fix, int n, const Real* __restrict__ uext, Real* __restrict__ vars, int64_t stride) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= n) return;
  const VarFixDev<Real> v = fix[k];
  vars[static_cast<size_t>(v.var) * stride + v.cell] = v.abb >= 0 ? uext[static_cast<size_t>(v.abb) * 3 + v.comp] : v.value;
}

// m_fold (and optionally its moments) of every cell, for read-back / output()
template <class L, class Real, bool STRICT>
__global__ void k_gather_all(const __grid_constant__ DevParams<Real> p, int32_t ncells, Real* __restrict__ fold_out, Real* __restrict__ mom_out) {
  using P = Phys<L, Real, STRICT>;
  constexpr int Q = L::Q, D = L::D;
  const int32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
  if(cell >= ncells) return;
  Real fold[Q];
  gather_any<L, Real, STRICT>(p, p.A, cell, fold);
  if(fold_out != nullptr) {
#pragma unroll
    for(int j = 0; j < Q; ++j) fold_out[static_cast<size_t>(j) * p.stride + cell] = fold[j];
  }
  if(mom_out != nullptr) {
    Real rho, u[D];
    P::moments(fold, rho, u);
#pragma unroll
    for(int d = 0; d < D; ++d) mom_out[static_cast<size_t>(d) * p.stride + cell] = u[d];
    mom_out[static_cast<size_t>(D) * p.stride + cell] = rho;
  }
}

// host layout (reference: array of structures, double, reference cell order) <-> device layout (SoA, Real, plan order)
template <class Real>
__global__ void k_unpack_aos(const double* __restrict__ aos, const int32_t* __restrict__ ref2dev, int64_t n, int width, Real* __restrict__ soa, int64_t stride) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(c >= n) return;
  const int32_t dv = ref2dev[c];
  for(int j = 0; j < width; ++j) soa[static_cast<size_t>(j) * stride + dv] = static_cast<Real>(aos[c * width + j]);
}
template <class Real>
__global__ void k_pack_aos(const Real* __restrict__ soa, const int32_t* __restrict__ ref2dev, int64_t n, int width, double* __restrict__ aos, int64_t stride) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(c >= n) return;
  const int32_t dv = ref2dev[c];
  for(int j = 0; j < width; ++j) aos[c * width + j] = static_cast<double>(soa[static_cast<size_t>(j) * stride + dv]);
}

// halo exchange: gather the outgoing populations into one contiguous buffer / scatter the received ones into the ghosts
template <class Real>
__global__ void k_halo_pack(const Real* __restrict__ f, const int64_t* __restrict__ index, int64_t n, Real* __restrict__ out) {
  const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(k < n) out[k] = f[index[k]];
}
template <class Real>
__global__ void k_halo_unpack(Real* __restrict__ f, const int64_t* __restrict__ index, int64_t n, const Real* __restrict__ in) {
  const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(k < n) f[index[k]] = in[k];
}

// initialCondition(): rho = 1, u = preset, f = feq  (solver.cpp:267-295)
template <class L, class Real, bool STRICT>
__global__ void k_init(Real* __restrict__ f, const Real* __restrict__ vars0, int64_t stride, int32_t ncells) {
  using P = Phys<L, Real, STRICT>;
  constexpr int Q = L::Q, D = L::D;
  const int32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
  if(cell >= ncells) return;
  Real u[D], feq[Q];
#pragma unroll
  for(int d = 0; d < D; ++d) u[d] = vars0[static_cast<size_t>(d) * stride + cell];
  const Real rho = vars0[static_cast<size_t>(D) * stride + cell];
  P::equilibrium(rho, u, feq);
#pragma unroll
  for(int j = 0; j < Q; ++j) f[static_cast<size_t>(j) * stride + cell] = feq[j];
}

// residual: sum_c |vars - varsold| per variable (solver.cpp:809-815). Two-pass and fixed-shape, so it is
// deterministic run to run; the reference's serial sum differs from it by rounding only.
template <class Real>
__global__ void k_residual(const Real* __restrict__ v, const Real* __restrict__ vo, int64_t stride, int64_t n, int nvar, double* __restrict__ partial) {
  __shared__ double s[32];
  for(int var = 0; var < nvar; ++var) {
    double acc = 0;
    for(int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; c < n; c += static_cast<int64_t>(gridDim.x) * blockDim.x)
      acc += fabs(static_cast<double>(v[var * stride + c]) - static_cast<double>(vo[var * stride + c]));
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if(threadIdx.x < 32) {
      double t = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.0;
#pragma unroll
      for(int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      if(threadIdx.x == 0) partial[static_cast<size_t>(var) * gridDim.x + blockIdx.x] = t;
    }
    __syncthreads();
  }
}

} // namespace lbm
