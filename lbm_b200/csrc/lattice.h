// lattice.h -- lattice descriptors shared by host planning code and device kernels.
//
// Direction order, opposite table and weights are those of the reference's LBMethod<> specialisations
// (/root/reference/src/lbm/constants.h:296-319 D2Q9, :322-365 D3Q19, :368-422 D3Q27); the rest population is
// last (index Q-1) and D3Q19 is the 18-direction prefix of D3Q27.  Tables live inside constexpr functions so
// that fully unrolled device loops fold them to immediates.
#pragma once
#include <cstdint>

#include "mrt_tables.h"

#if defined(__CUDACC__)
#define LBM_HD __host__ __device__ __forceinline__
#else
#define LBM_HD inline
#endif

namespace lbm {

template <int D, int Q>
struct Lattice;

template <>
struct Lattice<2, 9> {
  static constexpr int D = 2, Q = 9;
  static constexpr int NSEL = 9;            // neighbour chunks incl. self (3^D)
  static constexpr int CHUNK_LEVELS = 5;    // 2D chunk = 32 x 32 cells
  static constexpr int CHUNK = 1 << (2 * CHUNK_LEVELS);
  LBM_HD static constexpr int c(int i, int d) {
    constexpr int t[9][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {1, 1}, {1, -1}, {-1, -1}, {-1, 1}, {0, 0}};
    return t[i][d];
  }
  LBM_HD static constexpr int opp(int i) {
    constexpr int t[9] = {1, 0, 3, 2, 6, 7, 4, 5, 8};
    return t[i];
  }
  LBM_HD static constexpr double w(int i) { return i < 4 ? 1.0 / 9.0 : (i < 8 ? 1.0 / 36.0 : 4.0 / 9.0); }
};

LBM_HD constexpr int d3_c(int i, int d) {
  constexpr int t[27][3] = {{-1, 0, 0},  {1, 0, 0},   {0, -1, 0},   {0, 1, 0},   {0, 0, -1}, {0, 0, 1},  {-1, -1, 0},
                            {-1, 1, 0},  {1, -1, 0},  {1, 1, 0},    {-1, 0, -1}, {-1, 0, 1}, {1, 0, -1}, {1, 0, 1},
                            {0, -1, -1}, {0, -1, 1},  {0, 1, -1},   {0, 1, 1},   {-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1},
                            {-1, 1, 1},  {1, -1, -1}, {1, -1, 1},   {1, 1, -1},  {1, 1, 1},  {0, 0, 0}};
  return t[i][d];
}

template <>
struct Lattice<3, 19> {
  static constexpr int D = 3, Q = 19;
  static constexpr int NSEL = 27;
  static constexpr int CHUNK_LEVELS = 3;    // 3D chunk = 8 x 8 x 8 cells
  static constexpr int CHUNK = 1 << (3 * CHUNK_LEVELS);
  LBM_HD static constexpr int c(int i, int d) { return i == 18 ? 0 : d3_c(i, d); }
  LBM_HD static constexpr int opp(int i) {
    constexpr int t[19] = {1, 0, 3, 2, 5, 4, 9, 8, 7, 6, 13, 12, 11, 10, 17, 16, 15, 14, 18};
    return t[i];
  }
  LBM_HD static constexpr double w(int i) { return i < 6 ? 1.0 / 18.0 : (i < 18 ? 1.0 / 36.0 : 1.0 / 3.0); }
};

template <>
struct Lattice<3, 27> {
  static constexpr int D = 3, Q = 27;
  static constexpr int NSEL = 27;
  static constexpr int CHUNK_LEVELS = 3;
  static constexpr int CHUNK = 1 << (3 * CHUNK_LEVELS);
  LBM_HD static constexpr int c(int i, int d) { return d3_c(i, d); }
  LBM_HD static constexpr int opp(int i) {
    constexpr int t[27] = {1,  0,  3,  2,  5,  4,  9,  8,  7,  6,  13, 12, 11, 10,
                           17, 16, 15, 14, 25, 24, 23, 22, 21, 20, 19, 18, 26};
    return t[i];
  }
  LBM_HD static constexpr double w(int i) {
    return i < 6 ? 2.0 / 27.0 : (i < 18 ? 1.0 / 54.0 : (i < 26 ? 1.0 / 216.0 : 8.0 / 27.0));
  }
};


// ---- per-direction in-chunk layouts (3D) ---------------------------------------------------------------------------------
// Inside an 8^3 chunk the population array of direction j is stored with an axis the direction does NOT move along as the
// fastest one: a pull shifted by c_j then only ever moves whole 64-byte rows (8 reals along the fastest axis), so every
// 32-byte DRAM sector a chunk reads is used completely -- with x fastest for all directions the ten D3Q19 directions with
// c_x != 0 fetched one extra, 7/8 wasted sector per row from the x-neighbour chunk (+26 % sectors, profiles/r01_ncu_*).
//   layout 0: pos = x + 8 y + 64 z   (directions without x component; D3Q27 corners, which have no free axis; all of 2D)
//   layout 1: pos = y + 8 z + 64 x   (+-x, xz diagonals)
//   layout 2: pos = z + 8 x + 64 y   (+-y, xy diagonals)
// Opposite directions share a layout (c_opp = -c), so a bounce-back source row is the same row of the opposite array.
LBM_HD constexpr int layout_of_c(int D, int cx, int cy, int cz) {
  if(D != 3) return 0;
  const bool x = cx != 0, y = cy != 0, z = cz != 0;
  if(x && !y && !z) return 1;
  if(!x && y && !z) return 2;
  if(x && y && !z) return 2;
  if(x && !y && z) return 1;
  return 0;
}
template <class L>
LBM_HD constexpr int layout_of(int j) {
  return layout_of_c(L::D, L::c(j, 0), L::D > 1 ? L::c(j, 1) : 0, L::D > 2 ? L::c(j, 2) : 0);
}
// lexicographic in-chunk offset (x fastest) -> position in the array of a direction with layout `lay`; 3D chunks only (9 bits)
LBM_HD constexpr int lay_perm(int lay, int o) {
  return lay == 0 ? o : (lay == 1 ? ((o >> 3) | ((o & 7) << 6)) : ((o >> 6) | ((o & 63) << 3)));
}
LBM_HD constexpr int lay_perm_inv(int lay, int pos) {
  return lay == 0 ? pos : (lay == 1 ? (((pos & 63) << 3) | (pos >> 6)) : (((pos & 7) << 6) | (pos >> 3)));
}
// Which device cells live in chunk-shaped blocks (fast chunks, slow chunks, ghost blocks): only those are permuted.
struct PermRange { int32_t perm_end, gb_begin, gb_end; };
LBM_HD constexpr int32_t pop_slot(int lay, int32_t cell, PermRange r) {
  if(lay == 0) return cell;
  if(!(cell < r.perm_end || (cell >= r.gb_begin && cell < r.gb_end))) return cell;
  return (cell & ~511) | lay_perm(lay, cell & 511);
}

// ---- MRT moment bases (generated: tools/gen_mrt_tables.py -> mrt_tables.h) --------------------------------------------------
// Orthogonal integer basis M, rows: conserved moments first (density, momentum), then the non-conserved ones; kind 1 = shear
// (traceless second-order moments: their rate sets the viscosity), 2 = bulk, 3 = ghost (higher order).
template <class L>
struct MrtBasis;
#define LBM_MRT_BASIS(LATT, NAME, QQ)                                                                              \
  template <>                                                                                                      \
  struct MrtBasis<LATT> {                                                                                          \
    LBM_HD static constexpr int m(int k, int i) { constexpr int t[QQ][QQ] = LBM_MRT_##NAME##_M; return t[k][i]; }  \
    LBM_HD static constexpr int norm(int k) { constexpr int t[QQ] = LBM_MRT_##NAME##_NORM; return t[k]; }          \
    LBM_HD static constexpr int kind(int k) { constexpr int t[QQ] = LBM_MRT_##NAME##_KIND; return t[k]; }          \
  };
#define LBM_COMMA ,
LBM_MRT_BASIS(Lattice<2 LBM_COMMA 9>, D2Q9, 9)
LBM_MRT_BASIS(Lattice<3 LBM_COMMA 19>, D3Q19, 19)
LBM_MRT_BASIS(Lattice<3 LBM_COMMA 27>, D3Q27, 27)
#undef LBM_COMMA
#undef LBM_MRT_BASIS

// MRT base rate: the rate shared by the most non-conserved moments (ties: the one met first in basis order).  The operator is
// evaluated as  f' = f - s0 (f - feq) - sum_k (s_k - s0) / |M_k|^2 M_k^T M_k (f - feq):  rows whose rate equals s0 cost nothing, so a
// parametrisation with one common ghost rate needs the projections of the shear and bulk rows only.
inline double mrt_base_rate(const double* rates, int q, int ndim) {
  double best = rates[ndim + 1];
  int    best_n = 0;
  for(int k = ndim + 1; k < q; ++k) {
    int n = 0;
    for(int j = ndim + 1; j < q; ++j) n += rates[j] == rates[k] ? 1 : 0;
    if(n > best_n) { best_n = n; best = rates[k]; }
  }
  return best;
}

// Runtime view of the same tables for host-side planning code.
struct LatticeRT {
  int    D = 0, Q = 0, NSEL = 0, CHUNK = 0, CHUNK_LEVELS = 0;
  int    c[27][3] = {};
  int    opp[27]  = {};
  double w[27]    = {};
  int    lay[27]  = {}; // in-chunk layout of every direction's population array (layout_of)
};

template <class L>
inline LatticeRT make_rt() {
  LatticeRT r;
  r.D = L::D;
  r.Q = L::Q;
  r.NSEL = L::NSEL;
  r.CHUNK = L::CHUNK;
  r.CHUNK_LEVELS = L::CHUNK_LEVELS;
  for(int i = 0; i < L::Q; ++i) {
    for(int d = 0; d < L::D; ++d) r.c[i][d] = L::c(i, d);
    r.opp[i] = L::opp(i);
    r.w[i]   = L::w(i);
    r.lay[i] = layout_of<L>(i);
  }
  return r;
}

inline bool lattice_rt(int ndim, int ndist, LatticeRT* out) {
  // D1Q3 / D2Q5: Poisson equation types only (src/lbm/constants.h:248-292); no chunk geometry, they run in poisson.cuh
  if(ndim == 1 && ndist == 3) {
    LatticeRT r;
    r.D = 1; r.Q = 3;
    r.c[0][0] = -1; r.c[1][0] = 1;
    r.opp[0] = 1; r.opp[1] = 0; r.opp[2] = 2;
    r.w[0] = 1.0 / 6.0; r.w[1] = 1.0 / 6.0; r.w[2] = 2.0 / 3.0;
    *out = r;
    return true;
  }
  if(ndim == 2 && ndist == 5) {
    LatticeRT r;
    r.D = 2; r.Q = 5;
    const int c5[5][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {0, 0}};
    const int o5[5]    = {1, 0, 3, 2, 4};
    for(int i = 0; i < 5; ++i) {
      r.c[i][0] = c5[i][0];
      r.c[i][1] = c5[i][1];
      r.opp[i]  = o5[i];
      r.w[i]    = i < 4 ? 1.0 / 6.0 : 1.0 / 3.0;
    }
    *out = r;
    return true;
  }
  if(ndim == 2 && ndist == 9) { *out = make_rt<Lattice<2, 9>>(); return true; }
  if(ndim == 3 && ndist == 19) { *out = make_rt<Lattice<3, 19>>(); return true; }
  if(ndim == 3 && ndist == 27) { *out = make_rt<Lattice<3, 27>>(); return true; }
  return false;
}

// The reference's space-filling curve: per level the quadrant bits (x + 2y + 4z) go through a fixed look-up
// table and become one base-2^D digit, most significant level first, no rotation or reflection
// (/root/reference/include/common/math/hilbert.h:16-48, LUT at :31).
LBM_HD constexpr int sfc_lut(int quadrant) {
  constexpr int t[16] = {0, 3, 1, 2, 5, 4, 6, 7, 10, 9, 11, 8, 15, 14, 12, 13};
  return t[quadrant];
}
LBM_HD constexpr int sfc_lut_inv(int digit) {
  constexpr int t[16] = {0, 2, 3, 1, 5, 4, 6, 7, 11, 9, 8, 10, 14, 15, 13, 12};
  return t[digit];
}

// Link codes of the generic path: code >= 0 -> pull from device cell `code`, same direction.
// code < 0 -> bits 28..30 = kind, bits 0..27 = payload.
enum LinkKind : int {
  LK_COPY   = 0,  // payload -> copy table {cell, dir}: fold[c,j] = f[cell,dir]
  LK_BB     = 1,  // bounce back: fold[c,j] = f[c,opp j]
  LK_BB_ADD = 2,  // bounce back + up to 3 sequential addends (moving wall), payload -> addend table
  LK_ABB    = 3,  // anti bounce back (pressure), payload -> pressure entry
  LK_VALUE  = 4   // fold[c,j] = value table entry (stale slots: constant; periodic-with-pressure: per step)
};
LBM_HD constexpr int32_t link_code(int kind, int32_t payload) {
  return static_cast<int32_t>(0x80000000u | (static_cast<uint32_t>(kind) << 28) | static_cast<uint32_t>(payload));
}
LBM_HD constexpr int link_kind(int32_t code) { return (static_cast<uint32_t>(code) >> 28) & 7; }
LBM_HD constexpr int32_t link_payload(int32_t code) { return code & 0x0FFFFFFF; }

} // namespace lbm
