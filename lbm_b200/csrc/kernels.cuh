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
#include <cstring>

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
  PermRange              pr;     // which device cells sit in chunk-shaped blocks (per-direction in-chunk layouts, lattice.h)
};

template <class Real>
struct DevParams {
  const Real* A;   // populations before the step (post-collision of the previous step), SoA [Q][stride]
  Real*       B;   // populations after the step
  int64_t     stride;
  PermRange   pr;  // device cells below perm_end and in [gb_begin, gb_end) are stored in the per-direction in-chunk layouts
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
  // peer-to-peer halo: the first n_outer_tiles tiles of the launch hold populations a peer needs; every such tile that has been
  // written out bumps *outer_done, which the communication stream watches (nullptr: not used)
  unsigned long long* outer_done;
  int32_t             n_outer_tiles;
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

// index of population (direction j, device cell) in the SoA arrays A / B
template <class L>
__device__ __forceinline__ size_t pop_index(int j, int32_t cell, int64_t stride, PermRange pr) {
  return static_cast<size_t>(j) * stride + pop_slot(layout_of<L>(j), cell, pr);
}

// ------------------------------------------------------------------------------------------- arithmetic
template <class Real, bool STRICT>
struct Ar;
template <>
struct Ar<double, true> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  // a / d for a CONSTANT divisor d with y = RN(1 / d): q0 = RN(a y), r = a - q0 d (exact in one fused operation), RN(q0 + r y) -- the
  // final step of the classical fused-multiply-add division, correctly rounded for every a whose quotient neither overflows nor is
  // subnormal (Markstein 1990).  Bit-identical to __ddiv_rn at a quarter of its cost; the divisors here are cs^2 and two multiples.
  static __device__ __forceinline__ double div_const(double a, double d, double y) {
    const double q0 = __dmul_rn(a, y);
    const double r  = __fma_rn(-q0, d, a);
    return __fma_rn(r, y, q0);
  }
};
template <>
struct Ar<float, true> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float div_const(float a, float d, float y) {
    const float q0 = __fmul_rn(a, y);
    const float r  = __fmaf_rn(-q0, d, a);
    return __fmaf_rn(r, y, q0);
  }
};
template <class Real>
struct Ar<Real, false> {
  static __device__ __forceinline__ Real add(Real a, Real b) { return a + b; }
  static __device__ __forceinline__ Real sub(Real a, Real b) { return a - b; }
  static __device__ __forceinline__ Real mul(Real a, Real b) { return a * b; }
  static __device__ __forceinline__ Real div(Real a, Real b) { return a / b; }
  static __device__ __forceinline__ Real div_const(Real a, Real d, Real) { return a / d; }
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
  // correctly rounded reciprocals of the three constant divisors, in the arithmetic type (Ar::div_const)
  static __device__ __forceinline__ Real r_cssq() { return Real(1) / static_cast<Real>(1.0 / 3.0); }
  static __device__ __forceinline__ Real r_c2() { return Real(1) / static_cast<Real>(2.0 * (1.0 / 3.0) * (1.0 / 3.0)); }
  static __device__ __forceinline__ Real r_c3() { return Real(1) / static_cast<Real>(2.0 * (1.0 / 3.0)); }

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
      const Real t = A::sub(A::add(A::add(Real(1), A::div_const(cuv, cssq(), r_cssq())), A::div_const(A::mul(cuv, cuv), c2(), r_c2())), A::div_const(vs, c3(), r_c3()));
      return A::mul(A::mul(w, rho), t);
    } else {
      return w * rho * (Real(1) + Real(3) * cuv + Real(4.5) * cuv * cuv - Real(1.5) * vs);
    }
  }
  // equilibrium_func.h:109-111
  static __device__ __forceinline__ Real symm_eq_one(Real w, Real rho, Real cuv, Real vs) {
    if constexpr(STRICT) {
      const Real t = A::sub(A::add(Real(1), A::div_const(A::mul(cuv, cuv), c2(), r_c2())), A::div_const(vs, c3(), r_c3()));
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
      const Real vterm = A::div_const(vs, c3(), r_c3());
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
          const Real a   = A::div_const(cuv, cssq(), r_cssq());
          const Real b   = A::div_const(A::mul(cuv, cuv), c2(), r_c2());
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
  //   TRT: f_i' = f_i - omega+ (f+_i - feq+_i) - omega- (f-_i - feq-_i), symmetric / antisymmetric parts over opposite pairs
  //   MRT: f' = f - M^-1 S M (f - feq) in the orthogonal moment basis of lattice.h (MrtBasis), evaluated as
  //        f' = f - s0 (f - feq) - sum_k (s_k - s0) / |M_k|^2 M_k^T M_k (f - feq)   with s0 = p.omega the most common rate
  //        (mrt_base_rate) and p.rates[k] = (s_k - s0) / |row k|^2: rows relaxing at s0 are skipped (a uniform branch);
  //        the conserved rows are never touched, every other row sums to zero and is orthogonal to the momentum rows, so mass
  //        and momentum are conserved to rounding for any set of rates
  // moment K of f - feq: sum over the non-zero entries of row K in ascending direction order (entries are compile-time constants)
  template <int K, int I>
  static __device__ __forceinline__ void mrt_moment(const Real (&fneq)[Q], Real& m, bool first) {
    if constexpr(I < Q) {
      constexpr int c = MrtBasis<L>::m(K, I);
      if constexpr(c == 0) {
        mrt_moment<K, I + 1>(fneq, m, first);
      } else {
        const Real t = c == 1 ? fneq[I] : (c == -1 ? -fneq[I] : A::mul(static_cast<Real>(c), fneq[I]));
        m = first ? t : A::add(m, t);
        mrt_moment<K, I + 1>(fneq, m, false);
      }
    }
  }
  template <int K, int I>
  static __device__ __forceinline__ void mrt_back(Real d, Real (&f)[Q]) {
    if constexpr(I < Q) {
      constexpr int c = MrtBasis<L>::m(K, I);
      if constexpr(c != 0) f[I] = A::sub(f[I], c == 1 ? d : (c == -1 ? -d : A::mul(static_cast<Real>(c), d)));
      mrt_back<K, I + 1>(d, f);
    }
  }
  template <int K>
  static __device__ __forceinline__ void mrt_row(const DevParams<Real>& p, const Real (&fneq)[Q], Real (&f)[Q]) {
    if constexpr(K < Q) {
      if(p.rates[K] != Real(0)) { // same for every thread: rows relaxing at the base rate are done already
        Real m = 0;
        mrt_moment<K, 0>(fneq, m, true);
        mrt_back<K, 0>(A::mul(p.rates[K], m), f);
      }
      mrt_row<K + 1>(p, fneq, f);
    }
  }

  template <int COLL>
  static __device__ __forceinline__ void collide(const DevParams<Real>& p, const Real (&fo)[Q], const Real (&fe)[Q], Real (&f)[Q]) {
    if constexpr(COLL == COLL_BGK) {
#pragma unroll
      for(int i = 0; i < Q; ++i) {
        if constexpr(STRICT) f[i] = A::add(A::mul(p.om1, fo[i]), A::mul(p.omega, fe[i]));
        else f[i] = fo[i] + p.omega * (fe[i] - fo[i]);
      }
    } else if constexpr(COLL == COLL_TRT) {
#pragma unroll
      for(int i = 0; i < Q; ++i) {
        const int  j   = L::opp(i);
        const Real fp  = A::mul(Real(0.5), A::add(fo[i], fo[j]));
        const Real fm  = A::mul(Real(0.5), A::sub(fo[i], fo[j]));
        const Real fep = A::mul(Real(0.5), A::add(fe[i], fe[j]));
        const Real fem = A::mul(Real(0.5), A::sub(fe[i], fe[j]));
        f[i]           = A::sub(A::sub(fo[i], A::mul(p.omega, A::sub(fp, fep))), A::mul(p.omega_minus, A::sub(fm, fem)));
      }
    } else {
      Real fneq[Q];
#pragma unroll
      for(int i = 0; i < Q; ++i) {
        fneq[i] = A::sub(fo[i], fe[i]);
        f[i]    = A::sub(fo[i], A::mul(p.omega, fneq[i]));
      }
      mrt_row<D + 1>(p, fneq, f);
    }
  }
};

// ------------------------------------------------------------------------------------------- gathers
// anti-bounce-back slot of pressure entry `entry`, direction j, given the bounced population f[c, opp j]
// (bnd_pressure.h:100: fold[c,opp] = -f[c,dist] + 2 * symmEq(dist, p, u_ext))
template <class L, class Real, bool STRICT>
__device__ __forceinline__ Real abb_value(const SlotTables<Real>& p, int32_t entry, int j, Real fopp) {
  using P = Phys<L, Real, STRICT>;
  using A = Ar<Real, STRICT>;
  const int          oj = L::opp(j);
  const AbbDev<Real> e  = p.abb[entry];
  Real               u[L::D];
#pragma unroll
  for(int d = 0; d < L::D; ++d) u[d] = p.uext[static_cast<size_t>(entry) * 3 + d];
  const Real vs  = P::vsq(u);
  const Real cuv = P::cu_rt(oj, u);
  const Real se  = P::symm_eq_one(static_cast<Real>(L::w(oj)), e.p, cuv, vs);
  return A::add(-fopp, A::mul(Real(2), se));
}

// value of a non-pull slot of device cell `cell`, direction J (generic path)
template <class L, class Real, bool STRICT>
__device__ __noinline__ Real special_slot(const SlotTables<Real> p, const Real* __restrict__ Abuf, int32_t code, int32_t cell, int j) {
  using A = Ar<Real, STRICT>;
  const int     kind = link_kind(code);
  const int32_t pl   = link_payload(code);
  const int     oj   = L::opp(j);
  switch(kind) {
    case LK_COPY: {
      const CopySrcDev cs = p.copytab[pl];
      return Abuf[pop_index<L>(cs.dir, cs.cell, p.stride, p.pr)];
    }
    case LK_BB: return Abuf[pop_index<L>(oj, cell, p.stride, p.pr)];
    case LK_BB_ADD: {
      // bnd_dirichlet.h:92,111-117: fold = f; then one += per addend
      Real                  v = Abuf[pop_index<L>(oj, cell, p.stride, p.pr)];
      const AddEntryT<Real> e = p.addtab[pl];
      for(int t = 0; t < e.n; ++t) v = A::add(v, e.v[t]);
      return v;
    }
    case LK_ABB: return abb_value<L, Real, STRICT>(p, pl, j, Abuf[pop_index<L>(oj, cell, p.stride, p.pr)]);
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
    for(int j = 0; j < Q; ++j) fold[j] = Abuf[pop_index<L>(j, cell, p.stride, p.pr)];
    return;
  }
  int32_t code[Q - 1];
#pragma unroll
  for(int j = 0; j < Q - 1; ++j) code[j] = __ldg(&p.codes[static_cast<size_t>(j) * p.gen_stride + g]);
#pragma unroll
  for(int j = 0; j < Q - 1; ++j) {
    if(code[j] >= 0) fold[j] = Abuf[pop_index<L>(j, code[j], p.stride, p.pr)];
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
  // template offsets are positions in direction J's own in-chunk layout (plan.hpp: build_template + lay_perm)
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
  Real v = Abuf[pop_index<L>(L::opp(J), cell, p.stride, p.pr)]; // bnd_dirichlet.h:92
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
    for(int j = 0; j < Q; ++j) fold[j] = Abuf[pop_index<L>(j, cell, p.stride, p.pr)];
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

// ------------------------------------------------------------------------------------------- main kernels
// moments, equilibrium, collision of one cell: m_fold -> m_f (solver.cpp:513-613)
template <class L, class Real, bool STRICT, int COLL>
__device__ __forceinline__ void collide_cell(const DevParams<Real>& p, const Real (&fold)[L::Q], Real (&f)[L::Q], Real& rho, Real (&u)[L::D]) {
  using P = Phys<L, Real, STRICT>;
  Real feq[L::Q];
  P::moments(fold, rho, u);
  P::equilibrium(rho, u, feq);
  P::template collide<COLL>(p, fold, feq, f);
}

template <class L, class Real>
__device__ __forceinline__ void store_vars(const DevParams<Real>& p, int32_t cell, Real rho, const Real (&u)[L::D]) {
  if(p.vars_out != nullptr) {
#pragma unroll
    for(int d = 0; d < L::D; ++d) p.vars_out[static_cast<size_t>(d) * p.stride + cell] = u[d];
    p.vars_out[static_cast<size_t>(L::D) * p.stride + cell] = rho;
  }
}

template <class L, class Real, bool STRICT, int COLL>
__device__ __forceinline__ void update_and_store(const DevParams<Real>& p, int32_t cell, const Real (&fold)[L::Q]) {
  constexpr int Q = L::Q, D = L::D;
  Real rho, u[D], f[Q];
  collide_cell<L, Real, STRICT, COLL>(p, fold, f, rho, u);
  if(p.B != nullptr) { // nullptr: moments only (lbm_b200_get_moments)
#pragma unroll
    for(int j = 0; j < Q; ++j) p.B[pop_index<L>(j, cell, p.stride, p.pr)] = f[j];
  }
  store_vars<L, Real>(p, cell, rho, u);
}

#ifndef LBM_THREADS
#define LBM_THREADS 256
#endif
#ifndef LBM_MINBLOCKS
#define LBM_MINBLOCKS 2
#endif
constexpr int kThreads = LBM_THREADS;

// ---- generic path: one thread per cell, per-slot link codes (cells next to obstacles, ragged ends, slow chunks)
template <class L, class Real, bool STRICT, int COLL>
__global__ void __launch_bounds__(kThreads, LBM_MINBLOCKS) k_step_generic(const __grid_constant__ DevParams<Real> p) {
  const int32_t gl = blockIdx.x * kThreads + threadIdx.x;
  if(gl >= p.n_gen) return;
  const int32_t cell = p.gen_begin + p.gen_off + gl;
  Real          fold[L::Q];
  gather_generic<L, Real, STRICT>(p, p.A, cell, fold);
  update_and_store<L, Real, STRICT, COLL>(p, cell, fold);
}

// ---- chunk path: persistent CTAs, the pulled populations of a tile staged in shared memory ----------------------------------
// Up to three CTAs per SM walk TILES (a whole SFC chunk, or -- fp64 -- its lower / upper half along the slowest lexicographic axis)
// handed out by a global ticket counter.  For every tile the (Q-1) moving populations of its cells -- already PULLED, i.e.
// shifted by c_j -- are copied global -> shared with cp.async, NSTAGE (2) tiles deep, so the loads of the next tile are in flight
// while this one is collided.  Thanks to the per-direction in-chunk layouts (lattice.h) every copy is a 16-byte piece of a
// whole 32-byte sector that lies in exactly one (neighbour) chunk: no partially used DRAM sector, no per-cell index, no
// template table -- the source of a piece follows from the direction's constants and 3^D neighbour-chunk bases.  Pieces whose
// source chunk is a wall are redirected to the bounce-back source (same position in the opposite direction's array, which
// shares the layout); addends / anti-bounce-back are applied when the cell is collided.  Threads then read their cell's Q-1
// values from shared memory (XOR-swizzled 16-byte units: conflict free), collide, write the result back IN PLACE, and the
// stage is copied out shared -> global with 128-bit loads / stores.  The rest population never moves: it goes through
// registers.  Independent CTAs on one SM overlap each other's phases (copy issue / collide / copy out).
#ifndef LBM_FAST_THREADS
#define LBM_FAST_THREADS 256
#endif
constexpr int kFastThreads = LBM_FAST_THREADS;

template <class L>
__device__ __forceinline__ constexpr int count_unaligned_dirs() {
  int n = 0;
  for(int j = 0; j < L::Q - 1; ++j) {
    const int lay = layout_of<L>(j);
    const int ax0 = lay == 0 ? 0 : (lay == 1 ? 1 : 2);
    n += L::c(j, ax0 < L::D ? ax0 : 0) != 0 ? 1 : 0;
  }
  return n;
}

template <class L, class Real>
struct FastCfg {
  static constexpr int D = L::D, Q = L::Q, QM = L::Q - 1, CH = L::CHUNK, NSEL = L::NSEL;
  static constexpr int LB  = L::CHUNK_LEVELS;            // bits per axis inside a chunk
  static constexpr int S   = 1 << LB;                    // cells per axis
  static constexpr int EPU = 16 / static_cast<int>(sizeof(Real)); // reals per 16-byte unit
  // tiles per chunk: halves along the slowest lexicographic axis, as long as a half still covers whole 32-byte sectors along
  // that axis (3D fp32: 8 floats = one sector, so the chunk stays whole)
#ifdef LBM_FAST_NSPLIT
  static constexpr int NSPLIT = LBM_FAST_NSPLIT;
#else
  static constexpr int NSPLIT = (D == 3 && sizeof(Real) == 4) ? 1 : 2;
#endif
  static constexpr int TS  = CH / NSPLIT;                // cells per tile
  static constexpr int TB  = NSPLIT == 2 ? LB - 1 : LB;  // bits of the slowest axis inside a tile
  static constexpr int UPD = TS / EPU;                   // 16-byte units per direction and tile
  // directions that move along the fastest axis of their own layout (D3Q27 corners; 2D: c_x != 0) are staged UNSHIFTED along that
  // axis -- whole aligned rows like every other direction -- plus one extra column per direction with the element that comes from
  // the neighbour chunk in that axis; the shift happens when the cell reads its slot
  static constexpr int NU   = count_unaligned_dirs<L>();
  static constexpr int ROWS = TS / S;                    // rows (along the fastest lexicographic axis) per tile
  static constexpr int XCOL = NU * ROWS;                 // reals of all extra columns
  static constexpr int STAGE_ELEMS = QM * TS + XCOL;
  static constexpr int STAGE_BYTES = STAGE_ELEMS * static_cast<int>(sizeof(Real));
  static_assert(STAGE_BYTES % 16 == 0, "stages must keep 16-byte alignment");
  // Two stages per CTA and as many CTAs per SM as the 228 KB of shared memory hold (up to three): measured on B200 (256^3 D3Q19
  // fp64), 3 CTAs x 2 stages x 36.9 KB reach 0.97 of the copy bandwidth where 2 CTAs x 3 stages reached 0.87 -- independent CTAs hide
  // each other's barrier-separated phases better than deeper prefetch does.  The register budget follows (<= 85 at three CTAs; the
  // compiler fits the D3Q19 kernels into 76-80 without spilling).
#ifdef LBM_FAST_STAGES
  static constexpr int NSTAGE = LBM_FAST_STAGES;
#else
  static constexpr int NSTAGE = 2;
#endif
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES;
#ifdef LBM_FAST_MINBLOCKS
  static constexpr int MINB = LBM_FAST_MINBLOCKS;
#else
  static constexpr int MINB = (229000 / (SMEM_BYTES + 2048)) >= 3 ? 3 : ((229000 / (SMEM_BYTES + 2048)) >= 1 ? (229000 / (SMEM_BYTES + 2048)) : 1);
#endif
  static constexpr int RN = NSTAGE + 1;                  // ring of neighbour-base rows
  static constexpr int RT = NSTAGE + 2;                  // ring of tickets
  static constexpr int PAST_END = NSTAGE + 2;            // tickets every CTA draws beyond the last tile
  static constexpr int CPT = (TS + kFastThreads - 1) / kFastThreads; // cells per thread and tile
  static constexpr int SELF = D == 2 ? 4 : 13;
  static_assert(NSTAGE >= 2, "the tile pipeline needs two stages");
};

// lexicographic in-chunk offset o (x fastest) -> position in a direction's array: lay_perm (lattice.h).  Inside a TILE the same
// order is kept with the slowest lexicographic axis (z; y in 2D) reduced to its TB tile bits:
//   tile-local lexicographic offset ot = x | y << LB | zt << 2 LB      (3D; 2D: x | yt << LB)
//   tile position of layout 0: ot;  layout 1: y | zt << LB | x << (LB + TB);  layout 2: zt | x << TB | y << (TB + LB)
template <class L, class Real>
__device__ __forceinline__ constexpr int tile_perm(int lay, int ot) {
  using C = FastCfg<L, Real>;
  if(L::D != 3 || lay == 0) return ot;
  const int M = C::S - 1, x = ot & M, y = (ot >> C::LB) & M, zt = ot >> (2 * C::LB);
  return lay == 1 ? (y | (zt << C::LB) | (x << (C::LB + C::TB))) : (zt | (x << C::TB) | (y << (C::TB + C::LB)));
}
// tile position (layout `lay`, tile half h) -> position in the chunk's array of that direction
template <class L, class Real>
__device__ __forceinline__ constexpr int tile_to_chunk_pos(int lay, int tp, int h) {
  using C = FastCfg<L, Real>;
  if(C::NSPLIT == 1) return tp;
  if(L::D != 3 || lay == 0) return tp + h * C::TS;
  const int M = C::S - 1, MT = (1 << C::TB) - 1;
  if(lay == 1) return (tp & M) | ((((tp >> C::LB) & MT) | (h << C::TB)) << C::LB) | ((tp >> (C::LB + C::TB)) << (2 * C::LB));
  return (tp & MT) | (h << C::TB) | (((tp >> C::TB) & M) << C::LB) | ((tp >> (C::TB + C::LB)) << (2 * C::LB));
}

// per-direction layout ids kept in shared memory for run-time indexing (the step-0 copy loop)
template <class L, int J>
__device__ __forceinline__ void fill_dir_words(uint32_t* s_dir, int tid) {
  if constexpr(J < L::Q - 1) {
    constexpr uint32_t w = static_cast<uint32_t>(layout_of<L>(J));
    if(tid == J) s_dir[J] = w;
    fill_dir_words<L, J + 1>(s_dir, tid);
  }
}
template <class L>
__device__ __forceinline__ constexpr bool dir_aligned(int j) {
  // the direction does not move along the fastest axis of its layout
  const int lay = layout_of<L>(j);
  const int ax0 = lay == 0 ? 0 : (lay == 1 ? 1 : 2);
  return L::c(j, ax0 < L::D ? ax0 : 0) == 0;
}
template <class L>
__device__ __forceinline__ constexpr int count_aligned() {
  int n = 0;
  for(int j = 0; j < L::Q - 1; ++j) n += dir_aligned<L>(j) ? 1 : 0;
  return n;
}
// position of an unaligned direction among the unaligned ones
template <class L>
__device__ __forceinline__ constexpr int unaligned_index(int j) {
  int n = 0;
  for(int i = 0; i < j; ++i) n += dir_aligned<L>(i) ? 0 : 1;
  return n;
}
// k-th aligned (unaligned) direction
template <class L>
__device__ __forceinline__ constexpr int nth_dir(int k, bool aligned) {
  for(int j = 0; j < L::Q - 1; ++j)
    if(dir_aligned<L>(j) == aligned) {
      if(k == 0) return j;
      --k;
    }
  return 0;
}

// position inside a stage of element `tp` (tile position in the direction's layout): XOR swizzle of the 16-byte unit index, so
// that the cells of a (half-)warp hit different banks whichever axis is the fastest one of the direction (verified by brute force
// for fp64: conflict free in all three layouts, whole chunks and half-chunk tiles)
template <class L, class Real>
__device__ __forceinline__ constexpr int stage_swizzle(int lay, int tp) {
  using C = FastCfg<L, Real>;
  if(L::D != 3) return tp;
  if(sizeof(Real) == 8) {
    if(C::NSPLIT == 2) {
      if(lay == 0) return tp ^ (((tp >> 6) & 1) << 2);
      if(lay == 1) return tp ^ (((tp >> 5) & 3) << 1);
      return tp ^ (((tp >> 5) & 1) << 1);
    }
    if(lay == 0) return tp ^ (((tp >> 6) & 1) << 2);
    if(lay == 1) return tp ^ (((tp >> 6) & 3) << 1);
    return tp ^ ((((tp >> 4) & 1) << 1) | (((tp >> 6) & 1) << 2));
  }
  if(lay == 0) return tp ^ (((tp >> 6) & 1) << 4);
  if(lay == 1) return tp ^ ((((tp >> 6) & 1) << 2) | (((tp >> 7) & 1) << 4));
  return tp ^ (((tp >> 6) & 1) << 2);
}

// virtual thread index -> tile-local lexicographic offset.  3D: lane bits are (x0, x1, y0, z0, x2), which together with the
// swizzle above makes the shared-memory accesses of all three layouts conflict free (fp64); 2D: plain row order.
template <class L>
__device__ __forceinline__ constexpr int thread_cell(int v) {
  if(L::D != 3) return v;
  const int x = (v & 3) | (((v >> 4) & 1) << 2);
  const int y = ((v >> 2) & 1) | (((v >> 5) & 3) << 1);
  const int z = ((v >> 3) & 1) | (((v >> 7) & 3) << 1);
  return x | (y << 3) | (z << 6);
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
#else
  std::memcpy(smem_dst, gsrc, 16);
#endif
}
template <int BYTES>
__device__ __forceinline__ void cp_async_small(void* smem_dst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
#else
  std::memcpy(smem_dst, gsrc, BYTES);
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
#endif
}

#if defined(__CUDACC__)
extern __shared__ __align__(128) unsigned char lbm_dyn_smem[];
#else
alignas(128) static unsigned char lbm_dyn_smem[232448]; // CPU harness: one block at a time
#endif

// ---- the copy engine: which sector of which (neighbour) chunk a 16-byte unit of a stage comes from.  The direction is a template
// parameter, so shifts, selector weights and layout fold to immediates; what is left per unit and tile is one shared-memory read
// of the neighbour base, one address and the cp.async itself.  WALLS = false is the variant for interior chunks (no neighbour is
// missing, the chunk's wall descriptor id is -1): it skips the bounce-back redirection.  tp = tile position of the unit's first
// element in direction J's layout, h = which half of the chunk the tile is.
template <class L, class Real, int J, bool ALIGNED, bool WALLS>
__device__ __forceinline__ void issue_unit(const DevParams<Real>& p, const Real* __restrict__ Abuf, Real* __restrict__ stg,
                                           const int32_t* __restrict__ nb, int32_t base, int tp, int h, Real* dst_override = nullptr) {
  using C = FastCfg<L, Real>;
  constexpr int lay  = layout_of<L>(J);
  constexpr int AX0  = lay == 0 ? 0 : (lay == 1 ? 1 : 2), AX1 = lay == 0 ? 1 : (lay == 1 ? 2 : 0), AX2 = lay == 0 ? 2 : (lay == 1 ? 0 : 1);
  constexpr int P3[3] = {1, 3, 9};
  constexpr int sa = ALIGNED ? 0 : L::c(J, AX0);
  constexpr int sb = AX1 < L::D ? L::c(J, AX1 < L::D ? AX1 : 0) : 0;
  constexpr int sc = AX2 < L::D ? L::c(J, AX2 < L::D ? AX2 : 0) : 0;
  constexpr int MASK = C::S - 1;
  const int pos0 = tile_to_chunk_pos<L, Real>(lay, tp, h);
  int sel = C::SELF, srcpos = pos0;
  if constexpr(sa != 0) {
    const int a2 = (pos0 & MASK) - sa;
    sel += (a2 >> C::LB) * P3[AX0];
    srcpos = (srcpos & ~MASK) | (a2 & MASK);
  }
  if constexpr(sb != 0) {
    const int b2 = ((pos0 >> C::LB) & MASK) - sb;
    sel += (b2 >> C::LB) * P3[AX1];
    srcpos = (srcpos & ~(MASK << C::LB)) | ((b2 & MASK) << C::LB);
  }
  if constexpr(sc != 0) {
    const int c2 = (pos0 >> (2 * C::LB)) - sc;
    sel += (c2 >> C::LB) * P3[AX2];
    srcpos = (srcpos & (C::S * C::S - 1)) | ((c2 & MASK) << (2 * C::LB));
  }
  const int32_t nbv = nb[sel];
  // element offsets fit 32 bits (plan.hpp refuses Q * npad >= 2^32)
  const uint32_t jbase = static_cast<uint32_t>(J) * static_cast<uint32_t>(p.stride);
  uint32_t off = jbase + static_cast<uint32_t>(nbv + srcpos);
  if constexpr(WALLS) {
    // wall: the bounce-back source, i.e. the same position in the opposite direction's array (bnd_dirichlet.h:92)
    if(nbv < 0) off = static_cast<uint32_t>(L::opp(J)) * static_cast<uint32_t>(p.stride) + static_cast<uint32_t>(base + pos0);
  }
  Real* dst = dst_override != nullptr ? dst_override : stg + J * C::TS + stage_swizzle<L, Real>(lay, tp);
  if constexpr(ALIGNED) cp_async_16(dst, Abuf + off);
  else cp_async_small<static_cast<int>(sizeof(Real))>(dst, Abuf + off);
}

// Work distribution: the kFastThreads threads form NG = kFastThreads / UNITS groups (UNITS = units per direction and tile);
// group g serves the directions LI = g, g + NG, ... of the list, every thread one unit per direction -- the same unit for all of
// them, so the decode of the unit index is shared.  With fewer threads than units a thread loops over its units instead.
// DIRS: 0 = the aligned directions, 1 = the rows of the unaligned ones (copied like aligned rows: the shift along the fastest axis
// is left to the reader), 2 = their extra columns (one real per row: the element that comes from the neighbour in that axis)
template <class L, class Real, int DIRS, bool WALLS, int LI, int STEP>
__device__ __forceinline__ void issue_dirs(const DevParams<Real>& p, const Real* __restrict__ Abuf, Real* __restrict__ stg,
                                           const int32_t* __restrict__ nb, int32_t base, int h, int unit_first, int unit_step) {
  using C = FastCfg<L, Real>;
  constexpr int N = DIRS == 0 ? count_aligned<L>() : C::NU;
  constexpr int UNITS = DIRS == 2 ? C::ROWS : C::UPD;
  if constexpr(LI < N) {
    constexpr int J = nth_dir<L>(LI, DIRS == 0);
    for(int u = unit_first; u < UNITS; u += unit_step) {
      if constexpr(DIRS == 2) {
        // row u of the tile: the cell at the end the direction comes from takes its value from the neighbour along the fastest axis
        constexpr int a_edge = L::c(J, 0) > 0 ? 0 : C::S - 1;
        issue_unit<L, Real, J, false, WALLS>(p, Abuf, stg, nb, base, u * C::S + a_edge, h, stg + C::QM * C::TS + LI * C::ROWS + u);
      } else {
        issue_unit<L, Real, J, true, WALLS>(p, Abuf, stg, nb, base, u * C::EPU, h);
      }
    }
    issue_dirs<L, Real, DIRS, WALLS, LI + STEP, STEP>(p, Abuf, stg, nb, base, h, unit_first, unit_step);
  }
}
template <class L, class Real, int DIRS, bool WALLS, int G>
__device__ __forceinline__ void issue_groups(const DevParams<Real>& p, const Real* __restrict__ Abuf, Real* __restrict__ stg,
                                             const int32_t* __restrict__ nb, int32_t base, int h, int tid) {
  using C = FastCfg<L, Real>;
  constexpr int UNITS = DIRS == 2 ? C::ROWS : C::UPD;
  constexpr int NG = kFastThreads >= UNITS ? kFastThreads / UNITS : 1;
  static_assert(kFastThreads >= UNITS ? kFastThreads % UNITS == 0 : UNITS % kFastThreads == 0, "thread count vs units per direction");
  if constexpr(G < NG) {
    if(NG == 1 || tid / UNITS == G)
      issue_dirs<L, Real, DIRS, WALLS, G, NG>(p, Abuf, stg, nb, base, h, NG == 1 ? tid : tid % UNITS, NG == 1 ? kFastThreads : UNITS);
    issue_groups<L, Real, DIRS, WALLS, G + 1>(p, Abuf, stg, nb, base, h, tid);
  }
}

// issue the copies of one tile into a stage: all moving populations of its cells, pulled
template <class L, class Real>
__device__ __forceinline__ void issue_tile_loads(const DevParams<Real>& p, const Real* __restrict__ Abuf, Real* __restrict__ stg,
                                                 const int32_t* __restrict__ nb, int32_t base, int h, const uint32_t* __restrict__ s_dir, int tid) {
  using C = FastCfg<L, Real>;
  if(p.first) {
    // step 0: m_fold is the initial condition itself -- every cell reads its own slots (no shift, no walls)
    for(int g = tid; g < C::QM * C::UPD; g += kFastThreads) {
      const int J = g / C::UPD, tp = (g % C::UPD) * C::EPU, lay = static_cast<int>(s_dir[J]);
      cp_async_16(stg + J * C::TS + stage_swizzle<L, Real>(lay, tp),
                  Abuf + static_cast<size_t>(J) * p.stride + base + tile_to_chunk_pos<L, Real>(lay, tp, h));
    }
    return;
  }
  if(nb[C::NSEL] < 0) { // interior chunk
    issue_groups<L, Real, 0, false, 0>(p, Abuf, stg, nb, base, h, tid);
    if constexpr(C::NU > 0) {
      issue_groups<L, Real, 1, false, 0>(p, Abuf, stg, nb, base, h, tid);
      issue_groups<L, Real, 2, false, 0>(p, Abuf, stg, nb, base, h, tid);
    }
  } else {
    issue_groups<L, Real, 0, true, 0>(p, Abuf, stg, nb, base, h, tid);
    if constexpr(C::NU > 0) {
      issue_groups<L, Real, 1, true, 0>(p, Abuf, stg, nb, base, h, tid);
      issue_groups<L, Real, 2, true, 0>(p, Abuf, stg, nb, base, h, tid);
    }
  }
}

// copy a collided stage out to buffer B: 128-bit shared loads and global stores, whole 32-byte sectors.  Every thread serves the
// same 16-byte unit in all of its directions, so where that unit sits -- in the stage (swizzled) and in the chunk -- is computed
// once per kernel for the three layouts (UnitPos) instead of once per direction and tile.
struct UnitPos { int soff[3]; int cpos[3]; };
template <class L, class Real>
__device__ __forceinline__ constexpr int half_stride(int lay) { // what tile half h adds to a chunk position
  using C = FastCfg<L, Real>;
  if(C::NSPLIT == 1) return 0;
  if(L::D != 3 || lay == 0) return C::TS;
  return lay == 1 ? (1 << (C::TB + C::LB)) : (1 << C::TB);
}
template <class L, class Real>
__device__ __forceinline__ UnitPos unit_positions(int tp) {
  UnitPos u;
#pragma unroll
  for(int lay = 0; lay < 3; ++lay) {
    u.soff[lay] = stage_swizzle<L, Real>(lay, tp);
    u.cpos[lay] = tile_to_chunk_pos<L, Real>(lay, tp, 0);
  }
  return u;
}
template <class L, class Real, int J, int STEP>
__device__ __forceinline__ void copy_out_dirs(const DevParams<Real>& p, const Real* __restrict__ stg, int32_t base, int h, int unit_first, int unit_step,
                                              const UnitPos& up) {
  using C = FastCfg<L, Real>;
  if constexpr(J < C::QM) {
    constexpr int lay = layout_of<L>(J);
    if constexpr(kFastThreads >= C::UPD) {
      const uint32_t off = static_cast<uint32_t>(J) * static_cast<uint32_t>(p.stride) + static_cast<uint32_t>(base + up.cpos[lay] + h * half_stride<L, Real>(lay));
      *reinterpret_cast<uint4*>(p.B + off) = *reinterpret_cast<const uint4*>(stg + J * C::TS + up.soff[lay]);
    } else {
      for(int u = unit_first; u < C::UPD; u += unit_step) {
        const int tp = u * C::EPU;
        const uint32_t off = static_cast<uint32_t>(J) * static_cast<uint32_t>(p.stride) + static_cast<uint32_t>(base + tile_to_chunk_pos<L, Real>(lay, tp, h));
        *reinterpret_cast<uint4*>(p.B + off) = *reinterpret_cast<const uint4*>(stg + J * C::TS + stage_swizzle<L, Real>(lay, tp));
      }
    }
    copy_out_dirs<L, Real, J + STEP, STEP>(p, stg, base, h, unit_first, unit_step, up);
  }
}
template <class L, class Real, int G>
__device__ __forceinline__ void copy_out_groups(const DevParams<Real>& p, const Real* __restrict__ stg, int32_t base, int h, int tid, const UnitPos& up) {
  using C = FastCfg<L, Real>;
  constexpr int NG = kFastThreads >= C::UPD ? kFastThreads / C::UPD : 1;
  if constexpr(G < NG) {
    if(NG == 1 || tid / C::UPD == G) copy_out_dirs<L, Real, G, NG>(p, stg, base, h, NG == 1 ? tid : tid % C::UPD, NG == 1 ? kFastThreads : C::UPD, up);
    copy_out_groups<L, Real, G + 1>(p, stg, base, h, tid, up);
  }
}

// read the staged values of one cell.  Aligned directions: the cell's own slot.  Unaligned directions (rows staged unshifted along
// the fastest axis): the slot one step against the direction -- or the extra column at the row's end, or, when the row's source chunk
// is a wall and the row therefore holds the bounce-back sources, the cell's own slot.
template <class L, class Real, int J>
__device__ __forceinline__ void read_stage(const Real* __restrict__ stg, const int32_t* __restrict__ nb, bool walls, bool first, int ot, int o,
                                           const int (&pos)[3], Real (&fold)[L::Q]) {
  using C = FastCfg<L, Real>;
  if constexpr(J < L::Q - 1) {
    constexpr int lay = layout_of<L>(J);
    if constexpr(dir_aligned<L>(J)) {
      fold[J] = stg[J * C::TS + pos[lay]];
    } else {
      constexpr int sa = L::c(J, 0), a_edge = sa > 0 ? 0 : C::S - 1, LI = unaligned_index<L>(J);
      const int a = ot & (C::S - 1);
      if(first) {
        fold[J] = stg[J * C::TS + pos[0]];
      } else if(a == a_edge) {
        fold[J] = stg[C::QM * C::TS + LI * C::ROWS + (ot >> C::LB)];
      } else {
        bool wallrow = false;
        if(walls) { // selector of the row's source chunk: the shifts along the other axes only
          int sel = C::SELF, w3 = 3;
#pragma unroll
          for(int d = 1; d < L::D; ++d) {
            const int x = (o >> (d * C::LB)) & (C::S - 1);
            if(L::c(J, d) > 0 && x == 0) sel -= w3;
            if(L::c(J, d) < 0 && x == C::S - 1) sel += w3;
            w3 *= 3;
          }
          wallrow = nb[sel] < 0;
        }
        fold[J] = stg[J * C::TS + stage_swizzle<L, Real>(0, wallrow ? ot : ot - sa)];
      }
    }
    read_stage<L, Real, J + 1>(stg, nb, walls, first, ot, o, pos, fold);
  }
}

// the staged (pulled or bounced) value of slot J of a wall chunk's cell -> m_fold: moving-wall addends (bnd_dirichlet.h:111-117) or
// the anti-bounce-back form of a pressure face (bnd_pressure.h:100); interior slots pass through
template <class L, class Real, bool STRICT, int J>
__device__ __forceinline__ void wall_fixups(const DevParams<Real>& p, const int32_t* __restrict__ nb, const AddEntryT<Real>* __restrict__ wall,
                                            int32_t cell, int o, const int (&edge)[3][2], Real (&fold)[L::Q]) {
  using A = Ar<Real, STRICT>;
  using C = FastCfg<L, Real>;
  if constexpr(J < L::Q - 1) {
    // selector of the pull source: one step against c_J; edge[d][0] = the cell sits at coordinate 0 of axis d, [1] = at S-1
    int sel = C::SELF, w3 = 1;
#pragma unroll
    for(int d = 0; d < L::D; ++d) {
      if(L::c(J, d) > 0) sel -= edge[d][0] * w3;
      if(L::c(J, d) < 0) sel += edge[d][1] * w3;
      w3 *= 3;
    }
    if(sel != C::SELF && nb[sel] < 0) {
      const AddEntryT<Real> e = wall[sel * (L::Q - 1) + J];
      if(e.n < 0) {
        const int32_t entry = p.chunk_abb[static_cast<size_t>(p.chunk_abb_base[cell / L::CHUNK]) * L::CHUNK + o];
        fold[J] = abb_value<L, Real, STRICT>(p.tabs, entry, J, fold[J]);
      } else {
        Real v = fold[J]; // statically indexed: the entry stays in registers
        if(e.n > 0) v = A::add(v, e.v[0]);
        if(e.n > 1) v = A::add(v, e.v[1]);
        if(e.n > 2) v = A::add(v, e.v[2]);
        fold[J] = v;
      }
    }
    wall_fixups<L, Real, STRICT, J + 1>(p, nb, wall, cell, o, edge, fold);
  }
}

template <class L, class Real, bool STRICT, int COLL>
__global__ void __launch_bounds__(kFastThreads, FastCfg<L, Real>::MINB) k_step_fast(const __grid_constant__ DevParams<Real> p) {
  using C = FastCfg<L, Real>;
  constexpr int Q = L::Q, QM = Q - 1, CH = L::CHUNK, TS = C::TS, NSEL = L::NSEL, NSTAGE = C::NSTAGE, NSPLIT = C::NSPLIT;
  __shared__ int32_t  s_nb[C::RN][NSEL + 1];
  __shared__ int32_t  s_tk[C::RT];
  __shared__ uint32_t s_dir[32];
  Real* const stages = reinterpret_cast<Real*>(lbm_dyn_smem);
  const Real* __restrict__ Abuf = p.A;
  const int tid = threadIdx.x;
  const int32_t n_tiles = p.n_fast_chunks * NSPLIT;
  const UnitPos upos = unit_positions<L, Real>((tid % (kFastThreads >= C::UPD ? C::UPD : kFastThreads)) * C::EPU);

  fill_dir_words<L, 0>(s_dir, tid);
  // Tiles are handed out dynamically, in curve order, from a global ticket counter that only ever grows: a launch over n
  // tiles with B CTAs advances it by exactly n + B * PAST_END (every CTA draws PAST_END tickets beyond the end), so the
  // host knows the first ticket of every launch and never resets anything.
  // Thread 0 always has one more ticket in flight (tk_pending): it is published an iteration after it was drawn, so that the
  // latency of the atomic never sits in front of a barrier.
  unsigned long long tk_pending = 0;
  if(tid == 0) {
    unsigned long long t[NSTAGE + 2];
#pragma unroll
    for(int k = 0; k <= NSTAGE + 1; ++k) t[k] = atomicAdd(p.ticket, 1ull);
#pragma unroll
    for(int k = 0; k <= NSTAGE; ++k) s_tk[k] = static_cast<int32_t>(t[k] - p.ticket_base);
    tk_pending = t[NSTAGE + 1];
  }
  __syncthreads();
  if(s_tk[0] >= n_tiles) return;
  // neighbour bases of the first NSTAGE tiles
  for(int k = 0; k < NSTAGE; ++k) {
    const int32_t tk = s_tk[k];
    if(tk < n_tiles && tid < NSEL + 1) s_nb[k][tid] = p.chunk_nb[static_cast<size_t>(p.chunk_off + tk / NSPLIT) * (NSEL + 1) + tid];
  }
  __syncthreads();
  for(int k = 0; k < NSTAGE - 1; ++k) {
    const int32_t tk = s_tk[k];
    if(tk < n_tiles)
      issue_tile_loads<L, Real>(p, Abuf, stages + static_cast<size_t>(k) * C::STAGE_ELEMS, s_nb[k], (p.chunk_off + tk / NSPLIT) * CH, tk % NSPLIT, s_dir, tid);
    cp_async_commit();
  }

  int outer_mine = 0; // outer tiles this CTA has written since it last told the communication stream
  for(int i = 0;; ++i) {
    const int32_t ticket = s_tk[i % C::RT];
    if(ticket >= n_tiles) break;
    const int      chunk = p.chunk_off + ticket / NSPLIT;
    const int      h     = ticket % NSPLIT;
    const int32_t  base  = chunk * CH;
    Real* const    stg   = stages + static_cast<size_t>(i % NSTAGE) * C::STAGE_ELEMS;
    const int32_t* nb    = s_nb[i % C::RN];
    // the rest population does not move: straight through registers
    Real frest[C::CPT];
#pragma unroll
    for(int k = 0; k < C::CPT; ++k) {
      const int v = tid + k * kFastThreads;
      if(v < TS) frest[k] = Abuf[static_cast<size_t>(QM) * p.stride + base + h * TS + thread_cell<L>(v)];
    }
    cp_async_wait<NSTAGE - 2>();
    __syncthreads(); // B1: this tile's stage is complete; every thread has left the previous iteration
    // peer-to-peer halo: tickets grow, so the first inner tile means every outer tile of this CTA has been copied out (by all of
    // its threads: B1) -- ONE fence and ONE atomic per CTA and step tell the communication stream how many those were
    if(p.outer_done != nullptr && outer_mine > 0 && ticket >= p.n_outer_tiles) {
      if(tid == 0) {
        __threadfence();
        atomicAdd(p.outer_done, static_cast<unsigned long long>(outer_mine));
      }
      outer_mine = 0;
    }
    // neighbour bases of tile i + NSTAGE: fetched now, published before B2
    const int32_t tk_nb  = s_tk[(i + NSTAGE) % C::RT];
    int32_t       nb_val = 0;
    if(tk_nb < n_tiles && tid < NSEL + 1) nb_val = __ldg(&p.chunk_nb[static_cast<size_t>(p.chunk_off + tk_nb / NSPLIT) * (NSEL + 1) + tid]);
    {
      // tile i + NSTAGE - 1 goes into the stage the previous iteration has just copied out
      const int32_t tk = s_tk[(i + NSTAGE - 1) % C::RT];
      if(tk < n_tiles)
        issue_tile_loads<L, Real>(p, Abuf, stages + static_cast<size_t>((i + NSTAGE - 1) % NSTAGE) * C::STAGE_ELEMS, s_nb[(i + NSTAGE - 1) % C::RN],
                                  (p.chunk_off + tk / NSPLIT) * CH, tk % NSPLIT, s_dir, tid);
      cp_async_commit();
    }
    const int32_t wid = nb[NSEL]; // wall descriptor of this chunk, -1: interior chunk
    const AddEntryT<Real>* wall = p.wall_desc + static_cast<size_t>(wid < 0 ? 0 : wid) * QM * NSEL;
    // With unaligned directions a cell reads slots that a neighbour writes its result to: all cells of the tile read first (the
    // values wait in registers), one more barrier, then they collide and write.  Without them every cell owns its slots.
    Real folds[C::NU > 0 ? C::CPT : 1][Q];
    if constexpr(C::NU > 0) {
#pragma unroll
      for(int k = 0; k < C::CPT; ++k) {
        const int v = tid + k * kFastThreads;
        if(v < TS) {
          const int ot = thread_cell<L>(v), o = ot + h * TS;
          int pos[3];
          pos[0] = stage_swizzle<L, Real>(0, ot);
          pos[1] = L::D == 3 ? stage_swizzle<L, Real>(1, tile_perm<L, Real>(1, ot)) : ot;
          pos[2] = L::D == 3 ? stage_swizzle<L, Real>(2, tile_perm<L, Real>(2, ot)) : ot;
          read_stage<L, Real, 0>(stg, nb, wid >= 0, p.first != 0, ot, o, pos, folds[k]);
        }
      }
      __syncthreads();
    }
#pragma unroll
    for(int k = 0; k < C::CPT; ++k) {
      const int v = tid + k * kFastThreads;
      if(v < TS) {
        const int ot = thread_cell<L>(v);   // tile-local lexicographic offset
        const int o  = ot + h * TS;         // in-chunk lexicographic offset
        int pos[3];
        pos[0] = stage_swizzle<L, Real>(0, ot);
        pos[1] = L::D == 3 ? stage_swizzle<L, Real>(1, tile_perm<L, Real>(1, ot)) : ot;
        pos[2] = L::D == 3 ? stage_swizzle<L, Real>(2, tile_perm<L, Real>(2, ot)) : ot;
        Real fold[Q], f[Q], rho, u[L::D];
        if constexpr(C::NU > 0) {
#pragma unroll
          for(int j = 0; j < QM; ++j) fold[j] = folds[k][j];
        } else {
#pragma unroll
          for(int j = 0; j < QM; ++j) fold[j] = stg[j * TS + pos[layout_of<L>(j)]];
        }
        fold[QM] = frest[k];
        if(wid >= 0 && !p.first) {
          int edge[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
          for(int d = 0; d < L::D; ++d) {
            const int x = (o >> (d * C::LB)) & (C::S - 1);
            edge[d][0]  = x == 0 ? 1 : 0;
            edge[d][1]  = x == C::S - 1 ? 1 : 0;
          }
          wall_fixups<L, Real, STRICT, 0>(p, nb, wall, base + o, o, edge, fold);
        }
        collide_cell<L, Real, STRICT, COLL>(p, fold, f, rho, u);
        if(p.B != nullptr) {
#pragma unroll
          for(int j = 0; j < QM; ++j) stg[j * TS + pos[layout_of<L>(j)]] = f[j];
          p.B[static_cast<size_t>(QM) * p.stride + base + o] = f[QM];
        }
        store_vars<L, Real>(p, base + o, rho, u);
      }
    }
    if(tk_nb < n_tiles && tid < NSEL + 1) s_nb[(i + NSTAGE) % C::RN][tid] = nb_val;
    if(tid == 0) { // the ticket of tile i + NSTAGE + 1 was drawn an iteration ago; draw the next one
      s_tk[(i + NSTAGE + 1) % C::RT] = static_cast<int32_t>(tk_pending - p.ticket_base);
      tk_pending = atomicAdd(p.ticket, 1ull);
    }
    __syncthreads(); // B2: the stage holds m_f of the whole tile
    if(p.B != nullptr) copy_out_groups<L, Real, 0>(p, stg, base, h, tid, upos);
    if(ticket < p.n_outer_tiles) ++outer_mine;
  }
  cp_async_wait<0>();
  if(p.outer_done != nullptr && outer_mine > 0) { // this CTA ended on an outer tile
    __syncthreads();
    if(tid == 0) {
      __threadfence();
      atomicAdd(p.outer_done, static_cast<unsigned long long>(outer_mine));
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
    const Real fv  = p.B[pop_index<L>(i, e.val, p.stride, p.pr)];
    p.B[pop_index<L>(i, e.target, p.stride, p.pr)] = A::sub(A::add(P::eq_one(static_cast<Real>(L::w(i)), e.p, cuv, vs), fv), feq[i]);
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
    const Real fv  = p.B[pop_index<L>(i, e.cell, p.stride, p.pr)];
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
// POP: the SoA array is a population array (per-direction in-chunk layouts); otherwise a plain per-cell array (m_vars, scratch)
template <class L, class Real, bool POP>
__global__ void k_unpack_aos(const double* __restrict__ aos, const int32_t* __restrict__ ref2dev, int64_t n, int width, Real* __restrict__ soa, int64_t stride,
                             PermRange pr) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(c >= n) return;
  const int32_t dv = ref2dev[c];
  for(int j = 0; j < width; ++j) soa[POP ? pop_index<L>(j, dv, stride, pr) : static_cast<size_t>(j) * stride + dv] = static_cast<Real>(aos[c * width + j]);
}
// population upload, indexed by DESTINATION: thread t fills position t of every direction's array (coalesced stores); the cell
// that lives there follows from the direction's in-chunk layout, its row of the host array from dev2ref (reads stay inside the
// 512-row window of the chunk, which L1 / L2 absorb)
template <class L, class Real>
__global__ void k_unpack_aos_pop(const double* __restrict__ aos, const int32_t* __restrict__ dev2ref, int64_t npad, Real* __restrict__ soa, int64_t stride,
                                 PermRange pr) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(t >= npad) return;
  const int32_t ti = static_cast<int32_t>(t);
  const bool inblock = L::D == 3 && (ti < pr.perm_end || (ti >= pr.gb_begin && ti < pr.gb_end));
#pragma unroll
  for(int j = 0; j < L::Q; ++j) {
    const int     lay  = layout_of<L>(j);
    const int32_t cell = (inblock && lay != 0) ? ((ti & ~511) | lay_perm_inv(lay, ti & 511)) : ti;
    const int32_t ref  = dev2ref[cell];
    if(ref >= 0) soa[static_cast<size_t>(j) * stride + t] = static_cast<Real>(aos[static_cast<size_t>(ref) * L::Q + j]);
  }
}
template <class L, class Real, bool POP>
__global__ void k_pack_aos(const Real* __restrict__ soa, const int32_t* __restrict__ ref2dev, int64_t n, int width, double* __restrict__ aos, int64_t stride,
                           PermRange pr) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(c >= n) return;
  const int32_t dv = ref2dev[c];
  for(int j = 0; j < width; ++j) aos[c * width + j] = static_cast<double>(soa[POP ? pop_index<L>(j, dv, stride, pr) : static_cast<size_t>(j) * stride + dv]);
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

// ---- peer-to-peer halo (solver_fused.cuh: p2p_exchange): small kernels of the communication stream.  They are sized to fit beside
// the persistent chunk CTAs (one warp, a handful of registers), which is what lets the exchange run while the inner tiles are updated.
#if defined(__CUDACC__)
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
#else
static inline unsigned long long ld_volatile_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
#endif
constexpr long long kSpinLimit = 1ll << 31; // ~ tens of seconds: a peer that never signals ends in an error, not in a hung GPU

// wait until a local counter has reached `target` (the outer tiles of this step are written)
static __global__ void k_wait_counter(const unsigned long long* ctr, unsigned long long target, int* err) {
  if(threadIdx.x != 0) return;
  long long spins = 0;
  while(ld_volatile_u64(ctr) < target) {
    __nanosleep(200);
    if(++spins > kSpinLimit) { *err = 1; break; }
  }
  __threadfence();
}
// tell every peer that this rank's populations of exchange `seq` have landed in its mailbox
static __global__ void k_p2p_signal(unsigned long long* const* flags, int n, unsigned long long seq) {
  const int k = threadIdx.x;
  if(k >= n) return;
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long*>(flags[k]) = seq;
  __threadfence_system();
}
// wait until every peer has signalled exchange `seq` in this rank's mailbox
static __global__ void k_p2p_wait(const unsigned long long* flags, int n, unsigned long long seq, int* err) {
  const int k = threadIdx.x;
  if(k >= n) return;
  long long spins = 0;
  while(ld_volatile_u64(flags + k) < seq) {
    __nanosleep(200);
    if(++spins > kSpinLimit) { *err = 2; break; }
  }
  __threadfence_system();
}

// initialCondition(): rho = 1, u = preset, f = feq  (solver.cpp:267-295)
template <class L, class Real, bool STRICT>
__global__ void k_init(Real* __restrict__ f, const Real* __restrict__ vars0, int64_t stride, int32_t ncells, PermRange pr) {
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
  for(int j = 0; j < Q; ++j) f[pop_index<L>(j, cell, stride, pr)] = feq[j];
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
