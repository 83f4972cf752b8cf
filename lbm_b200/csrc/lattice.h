// lattice.h -- lattice descriptors shared by host planning code and device kernels.
//
// Direction order, opposite table and weights are those of the reference's LBMethod<> specialisations
// (/root/reference/src/lbm/constants.h:296-319 D2Q9, :322-365 D3Q19, :368-422 D3Q27); the rest population is
// last (index Q-1) and D3Q19 is the 18-direction prefix of D3Q27.  Tables live inside constexpr functions so
// that fully unrolled device loops fold them to immediates.
#pragma once
#include <cstdint>

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

// Runtime view of the same tables for host-side planning code.
struct LatticeRT {
  int    D = 0, Q = 0, NSEL = 0, CHUNK = 0, CHUNK_LEVELS = 0;
  int    c[27][3] = {};
  int    opp[27]  = {};
  double w[27]    = {};
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
