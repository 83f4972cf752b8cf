// grid_box.hpp -- synthetic uniform box in the reference's table format (host code, OpenMP).
//
// Restates for a box what the reference's pipeline would produce:
//   order      ascending key of hilbert::index (/root/reference/include/common/math/hilbert.h:16-48)
//   axis nghbr same-level neighbour, -1 outside the box; periodic sides are linked at grid level
//              (/root/reference/src/cartesiangrid.h:608-706)
//   diagonals  composition of axis steps, x then y then z (/root/reference/src/cartesiangrid.h:451-493; the 3D
//              case is "Not implemented" there -- the slot order used is LBMethod<D3Q27>::m_dirs,
//              /root/reference/src/lbm/constants.h:368-399).  For a box every intermediate cell of a composed step
//              exists exactly when the final cell does, so the composition reduces to a per-axis range test.
// Cells are enumerated by walking the keys in order, so no sort is needed; a prefix count maps key -> cell id.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "lattice.h"

namespace lbm {

inline void key_to_xyz(int ndim, int level, int64_t key, int64_t* xyz) {
  xyz[0] = xyz[1] = xyz[2] = 0;
  const int mask = (1 << ndim) - 1;
  for(int l = 0; l < level; ++l) {
    const int q = sfc_lut_inv(static_cast<int>((key >> (ndim * (level - 1 - l))) & mask));
    for(int d = 0; d < ndim; ++d) xyz[d] = (xyz[d] << 1) | ((q >> d) & 1);
  }
}
inline int64_t xyz_to_key(int ndim, int level, const int64_t* xyz) {
  int64_t key = 0;
  for(int l = 0; l < level; ++l) {
    int q = 0;
    for(int d = 0; d < ndim; ++d) q |= static_cast<int>((xyz[d] >> (level - 1 - l)) & 1) << d;
    key = (key << ndim) | sfc_lut(q);
  }
  return key;
}

// hilbert::index as the reference evaluates it: on unit-cube coordinates, by repeated halving in floating point
// (/root/reference/include/common/math/hilbert.h:16-48). Supports up to 4 dimensions like the reference.
inline int64_t sfc_index_unit(int ndim, const double* x, int level) {
  double pos[4] = {0, 0, 0, 0};
  for(int d = 0; d < ndim; ++d) pos[d] = x[d];
  int64_t index = 0;
  for(int l = 0; l < level; ++l) {
    int q = 0;
    for(int d = 0; d < ndim; ++d)
      if(pos[d] >= 0.5) q |= 1 << d;
    // index += 2^(ndim * (level - 1 - l)) * LUT[q] (hilbert.h:40-41).  Written as a sum, not as a bit field: in 1D the LUT maps the upper
    // half to 3, which does not fit one bit (keys 0, 3, 6, 9, ... -- still unique and ascending in x)
    index = index * (int64_t(1) << ndim) + sfc_lut(q);
    for(int d = 0; d < ndim; ++d) pos[d] = 2 * pos[d] - ((q >> d) & 1);
  }
  return index;
}

struct BoxIndex {
  int ndim = 0, level = 0;
  int64_t shape[3] = {1, 1, 1};
  int64_t n = 0;
  std::vector<int32_t> rank;    // key -> cell id or -1
  std::vector<int32_t> key_of;  // cell id -> key (32 bits suffice up to 1024^3 keys... stored as uint32 bit pattern)
};

inline bool box_index(int ndim, const int64_t* shape, BoxIndex& B, std::string* err) {
  if(ndim != 2 && ndim != 3) { *err = "box: ndim must be 2 or 3"; return false; }
  int64_t maxs = 0;
  for(int d = 0; d < ndim; ++d) {
    if(shape[d] <= 0) { *err = "box: bad shape"; return false; }
    B.shape[d] = shape[d];
    if(shape[d] > maxs) maxs = shape[d];
  }
  int level = 1;
  while((int64_t(1) << level) < maxs) ++level;
  if(ndim * level > 31) { *err = "box: too many keys for the 32-bit index"; return false; }
  B.ndim = ndim;
  B.level = level;
  const int64_t nkeys = int64_t(1) << (ndim * level);
  B.rank.assign(static_cast<size_t>(nkeys), -1);
  const int64_t BL = 1 << 16, nblocks = (nkeys + BL - 1) / BL;
  std::vector<int64_t> blocksum(static_cast<size_t>(nblocks) + 1, 0);
#pragma omp parallel for schedule(static)
  for(int64_t b = 0; b < nblocks; ++b) {
    int64_t cnt = 0;
    for(int64_t k = b * BL; k < std::min(nkeys, (b + 1) * BL); ++k) {
      int64_t xyz[3];
      key_to_xyz(ndim, level, k, xyz);
      bool in = true;
      for(int d = 0; d < ndim; ++d) in = in && xyz[d] < shape[d];
      B.rank[k] = in ? 1 : 0;
      cnt += in ? 1 : 0;
    }
    blocksum[b + 1] = cnt;
  }
  for(int64_t b = 0; b < nblocks; ++b) blocksum[b + 1] += blocksum[b];
  B.n = blocksum[nblocks];
  B.key_of.assign(static_cast<size_t>(B.n), 0);
#pragma omp parallel for schedule(static)
  for(int64_t b = 0; b < nblocks; ++b) {
    int64_t run = blocksum[b];
    for(int64_t k = b * BL; k < std::min(nkeys, (b + 1) * BL); ++k) {
      if(B.rank[k]) {
        B.rank[k] = static_cast<int32_t>(run);
        B.key_of[run] = static_cast<int32_t>(k);
        ++run;
      } else B.rank[k] = -1;
    }
  }
  return true;
}

// rows of the box's neighbour table for an arbitrary list of (global) cell ids -- what a rank of a partitioned run needs
inline bool box_rows(int ndim, const int64_t* shape, const int32_t* periodic, const int64_t* cells, int64_t ncells, int64_t* nghbr,
                     int stride, double* center, std::string* err) {
  BoxIndex B;
  if(!box_index(ndim, shape, B, err)) return false;
  const int nn = ndim == 2 ? 8 : 26;
  if(stride < nn) { *err = "box: stride too small"; return false; }
  int64_t maxs = 0;
  for(int d = 0; d < ndim; ++d) maxs = std::max(maxs, shape[d]);
  const double h = 1.0 / static_cast<double>(maxs);
  int dirs[26][3];
  for(int i = 0; i < nn; ++i)
    for(int d = 0; d < 3; ++d) dirs[i][d] = ndim == 2 ? (d < 2 ? Lattice<2, 9>::c(i, d) : 0) : Lattice<3, 27>::c(i, d);
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for(int64_t r = 0; r < ncells; ++r) {
    const int64_t c = cells[r];
    if(c < 0 || c >= B.n) { bad |= 1; continue; }
    int64_t xyz[3];
    key_to_xyz(ndim, B.level, static_cast<uint32_t>(B.key_of[c]), xyz);
    for(int i = 0; i < nn; ++i) {
      int64_t n[3] = {0, 0, 0};
      bool    ok   = true;
      for(int d = 0; d < ndim; ++d) {
        int64_t v = xyz[d] + dirs[i][d];
        if(periodic[d]) v = (v + shape[d]) % shape[d];
        else if(v < 0 || v >= shape[d]) ok = false;
        n[d] = v;
      }
      nghbr[r * stride + i] = ok ? B.rank[xyz_to_key(ndim, B.level, n)] : -1;
    }
    for(int i = nn; i < stride; ++i) nghbr[r * stride + i] = -1;
    if(center != nullptr)
      for(int d = 0; d < ndim; ++d) center[r * ndim + d] = (static_cast<double>(xyz[d]) + 0.5) * h;
  }
  if(bad) { *err = "box: cell id out of range"; return false; }
  return true;
}

inline bool box_topology(int ndim, const int64_t* shape, const int32_t* periodic, int64_t* nghbr, int stride, double* center,
                         int64_t* coords, std::string* err) {
  if(ndim != 2 && ndim != 3) { *err = "box: ndim must be 2 or 3"; return false; }
  const int nn = ndim == 2 ? 8 : 26;
  if(stride < nn) { *err = "box: stride too small"; return false; }
  int64_t maxs = 0;
  for(int d = 0; d < ndim; ++d) {
    if(shape[d] <= 0) { *err = "box: bad shape"; return false; }
    if(shape[d] > maxs) maxs = shape[d];
  }
  int level = 1;
  while((int64_t(1) << level) < maxs) ++level;
  const int64_t nkeys = int64_t(1) << (ndim * level);
  // rank[key] = number of kept cells with a smaller key
  std::vector<int32_t> rank(static_cast<size_t>(nkeys) + 1);
  const int64_t        BL = 1 << 16;
  const int64_t        nblocks = (nkeys + BL - 1) / BL;
  std::vector<int64_t> blocksum(static_cast<size_t>(nblocks) + 1, 0);
#pragma omp parallel for schedule(static)
  for(int64_t b = 0; b < nblocks; ++b) {
    int64_t cnt = 0;
    for(int64_t k = b * BL; k < std::min(nkeys, (b + 1) * BL); ++k) {
      int64_t xyz[3];
      key_to_xyz(ndim, level, k, xyz);
      bool in = true;
      for(int d = 0; d < ndim; ++d) in = in && xyz[d] < shape[d];
      rank[k] = in ? 1 : 0;
      cnt += in ? 1 : 0;
    }
    blocksum[b + 1] = cnt;
  }
  for(int64_t b = 0; b < nblocks; ++b) blocksum[b + 1] += blocksum[b];
#pragma omp parallel for schedule(static)
  for(int64_t b = 0; b < nblocks; ++b) {
    int64_t run = blocksum[b];
    for(int64_t k = b * BL; k < std::min(nkeys, (b + 1) * BL); ++k) {
      const int32_t in = rank[k];
      rank[k]          = in ? static_cast<int32_t>(run) : -1;
      run += in;
    }
  }
  const double h = 1.0 / static_cast<double>(maxs);
  // direction table: 2D in the reference's D2Q9 slot order, 3D in D3Q27 order
  int dirs[26][3];
  for(int i = 0; i < nn; ++i)
    for(int d = 0; d < 3; ++d) dirs[i][d] = ndim == 2 ? (d < 2 ? Lattice<2, 9>::c(i, d) : 0) : Lattice<3, 27>::c(i, d);
#pragma omp parallel for schedule(static)
  for(int64_t k = 0; k < nkeys; ++k) {
    const int64_t c = rank[k];
    if(c < 0) continue;
    int64_t xyz[3];
    key_to_xyz(ndim, level, k, xyz);
    for(int i = 0; i < nn; ++i) {
      int64_t n[3] = {0, 0, 0};
      bool    ok   = true;
      for(int d = 0; d < ndim; ++d) {
        int64_t v = xyz[d] + dirs[i][d];
        if(periodic[d]) v = (v + shape[d]) % shape[d];
        else if(v < 0 || v >= shape[d]) ok = false;
        n[d] = v;
      }
      nghbr[c * stride + i] = ok ? rank[xyz_to_key(ndim, level, n)] : -1;
    }
    for(int i = nn; i < stride; ++i) nghbr[c * stride + i] = -1;
    if(center != nullptr)
      for(int d = 0; d < ndim; ++d) center[c * ndim + d] = (static_cast<double>(xyz[d]) + 0.5) * h;
    if(coords != nullptr)
      for(int d = 0; d < ndim; ++d) coords[c * ndim + d] = xyz[d];
  }
  return true;
}

} // namespace lbm
