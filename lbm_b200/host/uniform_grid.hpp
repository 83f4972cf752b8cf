// uniform_grid.hpp -- rows of a single-level grid on demand (host code, OpenMP): the partitioned runs' view of the grid pipeline.
//
// grid.hpp builds the whole cell tree like the reference does (/root/reference/src/gridgenerator/cartesiangrid_generation.h);
// that is what a single GPU needs, but a rank of a domain-decomposed run only needs the table rows of its own SFC range, and the
// reference has no decomposition to mirror.  For partitionLevel == uniformLevel == maxRfnmtLvl the result of
// the pipeline has a closed form, evaluated here cell by cell:
//   kept cells    a cell of the 2^L bounding cube survives iff it is cut by an object or its centre is inside the flow region
//                 (markOutsideCells / floodCells :554-603 reduce to this when the cut cells separate inside from outside, which
//                 is checked against grid.hpp in tests/test_uniform_grid.py); centres are accumulated level by level exactly
//                 like refineCell :428-431, so they are bit-identical to the generated ones
//   order         ascending key of hilbert::index (include/common/math/hilbert.h:16-48): id = key - #removed cells below it
//   axis nghbrs   the same-level neighbour if it is kept (findChildLevelNghbrs :453-506 + deletion :541-546)
//   diagonals     composition of axis steps x, y, z; -1 as soon as one intermediate cell is missing (src/cartesiangrid.h:451-493)
//   surfaces      identifyBndrySurfaces (src/cartesiangrid.h:553-603) over the cells that miss an axis neighbour
// Only the sorted list of removed keys is stored (the solid: O(volume of the obstacle)), never a table over the whole cube.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../csrc/grid_box.hpp"
#include "geometry.hpp"
#include "grid.hpp"
#include "json.hpp"

namespace lbmhost {

class UniformGrid {
 public:
  int     ndim = 3, level = 0, nn_axis = 6, nn_diag = 26;
  int64_t side = 1, nkeys = 1, n = 0;
  double  bbmin[3] = {0, 0, 0}, bbmax[3] = {0, 0, 0}, cog[3] = {0, 0, 0}, length_on_level[64] = {0};
  std::shared_ptr<GeometryManager> geom;
  std::vector<int64_t> removed;     // sorted keys of the cube cells that are not part of the grid
  std::vector<Surface> surfaces;    // as SolverGrid::surfaces, built by build_surfaces()
  Json                 boundary;

  void configure(const Json& cfg) {
    GridGen g; // same parsing, same bounding box / root length arithmetic
    g.configure(cfg);
    if(g.part_level != g.level || g.uni_level != g.level) throw std::runtime_error("UniformGrid: single-level grids only");
    if(g.align) throw std::runtime_error("UniformGrid: alignNodesWithSurface is not supported");
    ndim = g.ndim;
    level = g.level;
    nn_axis = 2 * ndim;
    nn_diag = ndim == 2 ? 8 : 26;
    if(ndim * level > 62) throw std::runtime_error("UniformGrid: too many levels");
    side = int64_t(1) << level;
    nkeys = int64_t(1) << (ndim * level);
    geom = g.geom;
    for(int d = 0; d < 3; ++d) { bbmin[d] = g.bbmin[d]; bbmax[d] = g.bbmax[d]; cog[d] = g.cog[d]; }
    for(int l = 0; l < 64; ++l) length_on_level[l] = g.length_on_level[l];
    if(cfg.has("solver") && cfg.at("solver").has("boundary")) {
      boundary = cfg.at("solver").at("boundary");
      for(const auto& gk : boundary.obj)
        for(const auto& sk : gk.second.obj)
          if(sk.second.opt_str("type", "notset") == "periodic" && !sk.second.opt_bool("generateBndry", true))
            throw std::runtime_error("UniformGrid: grid-level periodic connections are not supported (use the box provider)");
    }
    scan();
  }

  double cell_length() const { return length_on_level[level]; }

  // centre of the cube cell with integer coordinates xyz: refineCell's accumulation, level by level
  void center_of(const int64_t* xyz, double* c) const {
    for(int d = 0; d < ndim; ++d) {
      double x = cog[d];
      for(int l = 1; l <= level; ++l) x = x + 0.5 * length_on_level[l] * (((xyz[d] >> (level - l)) & 1) ? 1 : -1);
      c[d] = x;
    }
  }
  bool kept_xyz(const int64_t* xyz) const {
    for(int d = 0; d < ndim; ++d)
      if(xyz[d] < 0 || xyz[d] >= side) return false;
    double c[3];
    center_of(xyz, c);
    return geom->cut_with_cell(c, cell_length()) || geom->point_inside(c);
  }
  bool is_removed(int64_t key) const { return std::binary_search(removed.begin(), removed.end(), key); }
  // cell id of a kept key
  int64_t id_of_key(int64_t key) const { return key - (std::lower_bound(removed.begin(), removed.end(), key) - removed.begin()); }
  int64_t key_of_id(int64_t id) const {
    // the number of removed keys below the answer is the first i with removed[i] - i > id
    int64_t lo = 0, hi = static_cast<int64_t>(removed.size());
    while(lo < hi) {
      const int64_t mid = (lo + hi) / 2;
      if(removed[mid] - mid > id) hi = mid;
      else lo = mid + 1;
    }
    return id + lo;
  }
  // id of the cell at xyz, -1 if it is outside the cube or removed
  int64_t id_at(const int64_t* xyz) const {
    for(int d = 0; d < ndim; ++d)
      if(xyz[d] < 0 || xyz[d] >= side) return -1;
    const int64_t key = lbm::xyz_to_key(ndim, level, xyz);
    if(is_removed(key)) return -1;
    return id_of_key(key);
  }
  // N(c, slot) for c at xyz: axis step, or the composition of axis steps in x, y, z order
  int64_t neighbor_at(const int64_t* xyz, int slot) const {
    int64_t cur[3] = {xyz[0], xyz[1], ndim == 3 ? xyz[2] : 0};
    int64_t id = -1;
    for(int d = 0; d < ndim; ++d) {
      const int cd = ndim == 2 ? lbm::Lattice<2, 9>::c(slot, d) : lbm::Lattice<3, 27>::c(slot, d);
      if(cd == 0) continue;
      cur[d] += cd;
      id = id_at(cur);
      if(id < 0) return -1;
    }
    return id;
  }

  // rows of the push table (and centres) for arbitrary cell ids
  bool rows(const int64_t* ids, int64_t count, int64_t* nghbr, int stride, double* center, std::string* err) const {
    if(stride < nn_diag) { *err = "UniformGrid: stride too small"; return false; }
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for(int64_t r = 0; r < count; ++r) {
      if(ids[r] < 0 || ids[r] >= n) { bad |= 1; continue; }
      int64_t xyz[3];
      lbm::key_to_xyz(ndim, level, key_of_id(ids[r]), xyz);
      for(int s = 0; s < nn_diag; ++s) nghbr[r * stride + s] = neighbor_at(xyz, s);
      for(int s = nn_diag; s < stride; ++s) nghbr[r * stride + s] = -1;
      if(center != nullptr) center_of(xyz, center + r * ndim);
    }
    if(bad) { *err = "UniformGrid: cell id out of range"; return false; }
    return true;
  }
  // pull sources: src[r*stride + s] = the cell whose push in direction s lands in ids[r] (-1 if none).  The table is not
  // symmetric where a composed diagonal step passes through a removed cell, so the source's own composition is evaluated.
  bool sources(const int64_t* ids, int64_t count, int64_t* src, int stride, std::string* err) const {
    if(stride < nn_diag) { *err = "UniformGrid: stride too small"; return false; }
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for(int64_t r = 0; r < count; ++r) {
      if(ids[r] < 0 || ids[r] >= n) { bad |= 1; continue; }
      int64_t xyz[3];
      lbm::key_to_xyz(ndim, level, key_of_id(ids[r]), xyz);
      for(int s = 0; s < nn_diag; ++s) {
        int64_t from[3] = {xyz[0], xyz[1], ndim == 3 ? xyz[2] : 0};
        for(int d = 0; d < ndim; ++d) from[d] -= ndim == 2 ? lbm::Lattice<2, 9>::c(s, d) : lbm::Lattice<3, 27>::c(s, d);
        const int64_t sid = id_at(from);
        src[r * stride + s] = (sid >= 0 && neighbor_at(from, s) == ids[r]) ? sid : -1;
      }
      for(int s = nn_diag; s < stride; ++s) src[r * stride + s] = -1;
    }
    if(bad) { *err = "UniformGrid: cell id out of range"; return false; }
    return true;
  }

  // determineBoundaryCells + identifyBndrySurfaces (src/cartesiangrid.h:510-603) restricted to the cells that can qualify: those
  // with a missing axis neighbour (cube faces, cells next to a removed cell), visited in ascending id like the reference's loops
  void build_surfaces() {
    surfaces.clear();
    std::vector<int64_t> cand;
    // cube faces
    for(int d = 0; d < ndim; ++d) {
      const int64_t plane = ndim == 2 ? side : side * side;
      for(int sgn = 0; sgn < 2; ++sgn) {
#pragma omp parallel
        {
          std::vector<int64_t> mine;
#pragma omp for schedule(static) nowait
          for(int64_t p = 0; p < plane; ++p) {
            int64_t xyz[3] = {0, 0, 0};
            int64_t rest = p;
            for(int e = 0; e < ndim; ++e) {
              if(e == d) { xyz[e] = sgn ? side - 1 : 0; continue; }
              xyz[e] = rest % side;
              rest /= side;
            }
            const int64_t id = id_at(xyz);
            if(id >= 0) mine.push_back(id);
          }
#pragma omp critical
          cand.insert(cand.end(), mine.begin(), mine.end());
        }
      }
    }
    // kept axis neighbours of removed cells
#pragma omp parallel
    {
      std::vector<int64_t> mine;
#pragma omp for schedule(static) nowait
      for(int64_t i = 0; i < static_cast<int64_t>(removed.size()); ++i) {
        int64_t xyz[3];
        lbm::key_to_xyz(ndim, level, removed[i], xyz);
        for(int dir = 0; dir < nn_axis; ++dir) {
          int64_t nb[3] = {xyz[0], xyz[1], xyz[2]};
          nb[dir / 2] += dir % 2 ? 1 : -1;
          const int64_t id = id_at(nb);
          if(id >= 0) mine.push_back(id);
        }
      }
#pragma omp critical
      cand.insert(cand.end(), mine.begin(), mine.end());
    }
    std::sort(cand.begin(), cand.end());
    cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
    const int64_t nc = static_cast<int64_t>(cand.size());
    std::vector<double>  ctr(static_cast<size_t>(nc) * ndim);
    std::vector<int64_t> nb(static_cast<size_t>(nc) * nn_axis);
    std::vector<char>    bnd(static_cast<size_t>(nc), 0);
    const double len = cell_length();
#pragma omp parallel for schedule(static)
    for(int64_t k = 0; k < nc; ++k) {
      int64_t xyz[3];
      lbm::key_to_xyz(ndim, level, key_of_id(cand[k]), xyz);
      center_of(xyz, &ctr[k * ndim]);
      int have = 0;
      for(int dir = 0; dir < nn_axis; ++dir) {
        nb[k * nn_axis + dir] = neighbor_at(xyz, dir);
        have += nb[k * nn_axis + dir] != -1;
      }
      bnd[k] = geom->cut_with_cell(&ctr[k * ndim], len) && have != nn_axis;
    }
    std::set<std::pair<int64_t, int>> assigned;
    for(const auto& gk : boundary.obj) {
      const size_t nkeys_g = gk.second.size();
      for(const auto& sk : gk.second.obj) {
        Surface s;
        s.name = nkeys_g > 1 ? gk.first + "_" + sk.first : gk.first;
        const int d0 = sk.first == "all" ? 0 : SolverGrid::dir_id(sk.first);
        const int d1 = sk.first == "all" ? nn_axis : SolverGrid::dir_id(sk.first) + 1;
        if(d1 > nn_axis) throw std::runtime_error("ERROR: Invalid direction " + sk.first);
        for(int dir = d0; dir < d1; ++dir) {
          for(int64_t k = 0; k < nc; ++k) {
            if(!bnd[k] || nb[k * nn_axis + dir] != -1) continue;
            if(!geom->cut_with_cell(gk.first, &ctr[k * ndim], len)) continue;
            if(!assigned.insert({cand[k], dir}).second) continue;
            s.cells.push_back(cand[k]);
            std::array<double, 3> nrm = {0, 0, 0};
            nrm[dir / 2] = dir % 2 ? 1.0 : -1.0;
            s.normal[cand[k]] = nrm;
          }
        }
        surfaces.push_back(std::move(s));
      }
    }
  }

 private:
  // one pass over the cube in key order: the keys whose cell is not kept
  void scan() {
    const int64_t BL = int64_t(1) << 15, nblocks = (nkeys + BL - 1) / BL;
    std::vector<std::vector<int64_t>> per_block(static_cast<size_t>(nblocks));
#pragma omp parallel for schedule(dynamic, 16)
    for(int64_t b = 0; b < nblocks; ++b) {
      for(int64_t k = b * BL; k < std::min(nkeys, (b + 1) * BL); ++k) {
        int64_t xyz[3];
        lbm::key_to_xyz(ndim, level, k, xyz);
        if(!kept_xyz(xyz)) per_block[b].push_back(k);
      }
    }
    removed.clear();
    for(auto& v : per_block) removed.insert(removed.end(), v.begin(), v.end());
    n = nkeys - static_cast<int64_t>(removed.size());
  }
};

} // namespace lbmhost
