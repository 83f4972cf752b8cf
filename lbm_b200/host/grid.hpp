// grid.hpp -- host topology builder: JSON geometry -> SFC-ordered cell list, neighbour table, boundary surfaces.
//
// Rebuilds, from scratch, the tables the reference's grid pipeline hands to the LBM solver,
// bit-exactly, including its quirks:
//   GridGen::generate   root cell + level-by-level refinement, cut flags inherited from boundary parents, neighbour linking,
//                       flood-fill inside/outside marking and deletion with swap-from-the-end
//                       (/root/reference/src/gridgenerator/cartesiangrid_generation.h:71-147, 419-603), then ordering by the key
//                       of hilbert::index of the normalised centre (:664-718, include/common/math/hilbert.h:16-48)
//   SolverGrid::load    leaf/bndry properties (src/cartesiangrid.h:502-546), named boundary surfaces with first-come
//                       (cell,dir) assignment and last-normal-wins (:553-603, src/common/surface.h:52-55), grid-level periodic
//                       links (:608-706), diagonal neighbours by composition of axis steps (:451-493)
// Multi-level grids: partitionLevel < uniformLevel keeps the cells of every level from the partition
// level up in the list (partition cells in curve order, then one block per level in parent order with the deletion swaps of
// cartesiangrid_generation.h:508-552), maxRfnmtLvl > uniformLevel adds the children of the cut cells of the highest level
// (gridGenerator.cpp:206-216, markBndryCells / refineMarkedCells :186-230).  The reference steps every level as its own lattice
// ("todo: skip non-leaf cells", solver.cpp:525), so the tables are all the solver needs.  alignNodesWithSurface is supported
// on single-level grids only (every reference configuration that uses it is single-level).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <map>
#include <set>
#include <stack>
#include <string>
#include <vector>

#include "../csrc/grid_box.hpp"
#include "geometry.hpp"
#include "json.hpp"

namespace lbmhost {

struct GenCell {
  double  center[3] = {0, 0, 0};
  int64_t nghbr[6]  = {-1, -1, -1, -1, -1, -1};
  int64_t child[8]  = {-1, -1, -1, -1, -1, -1, -1, -1};
  int64_t parent    = -1;
  bool    bndry = false, inside = false, marked = false;
  int     level = 0;
};

// cartesian::childDir / nghbrInside / nghbrParentChildId (include/common/math/cartesian.h:72-153) for up to 3 dimensions:
// child id bit d = sign of the offset along dimension d
inline int child_dir(int child, int d) { return ((child >> d) & 1) ? 1 : -1; }
// neighbour of child `c` in direction `dir` inside the same parent (-1 if it lies in the neighbouring parent)
inline int nghbr_inside(int c, int dir) {
  const int d = dir / 2, positive = dir % 2;
  const int bit = (c >> d) & 1;
  if(positive == bit) return -1; // already on that side of the parent
  return c ^ (1 << d);
}
// child of the neighbouring parent that touches child `c` across direction `dir`
inline int nghbr_parent_child(int c, int dir) {
  const int d = dir / 2, positive = dir % 2;
  const int bit = (c >> d) & 1;
  if(positive != bit) return -1;
  return c ^ (1 << d);
}

class GridGen {
 public:
  int    ndim = 2, level = 0;         // level = maxRfnmtLvl (CartesianGrid::maxLvl, gridGenerator.cpp:185)
  int    part_level = 0, uni_level = 0;
  double bbmin[3] = {0, 0, 0}, bbmax[3] = {0, 0, 0}, cog[3] = {0, 0, 0};
  double length_on_level[64] = {0};
  std::shared_ptr<GeometryManager> geom;
  std::vector<GenCell> cells; // final level, SFC order (after alignment: with the holes of deleted cells filled from the end)
  bool   align = false;
  int    align_dir = 1;

  void configure(const Json& cfg) {
    ndim = static_cast<int>(cfg.at("dim").as_int());
    if(ndim < 1 || ndim > 3) throw std::runtime_error("Only dimensions 1, 2 and 3 are supported by this host.");
    const long long part = cfg.at("partitionLevel").as_int(), uni = cfg.at("uniformLevel").as_int();
    const long long maxr = cfg.opt_int("maxRfnmtLvl", uni);
    if(part > uni) throw std::runtime_error("Invalid definition of grid level partitionLevel >= uniformLevel");
    if(maxr < uni) throw std::runtime_error("Invalid definition of grid level uniformLevel >= maxRfnmtLvl");
    align = cfg.opt_bool("alignNodesWithSurface", false);
    if(align && (part != uni || maxr != uni))
      throw std::runtime_error("alignNodesWithSurface on a multi-level grid is not supported");
    align_dir = static_cast<int>(cfg.opt_int("alignDir", 1));
    if(align && (align_dir < 0 || align_dir >= ndim)) throw std::runtime_error("Invalid alignDir");
    level = static_cast<int>(maxr);
    part_level = static_cast<int>(part);
    uni_level  = static_cast<int>(uni);
    geom  = std::make_shared<GeometryManager>();
    if(!cfg.has("geometry")) throw std::runtime_error("The required configuration value is missing: geometry");
    geom->setup(cfg.at("geometry"), ndim);
    // bounding box: gridGenerator.cpp:263-275
    if(geom->size() == 0 || cfg.has("boundingBox")) {
      if(!cfg.has("boundingBox")) throw std::runtime_error("no geometry and no boundingBox");
      const auto bb = cfg.at("boundingBox").as_doubles();
      for(int d = 0; d < ndim; ++d) { bbmin[d] = bb[2 * d]; bbmax[d] = bb[2 * d + 1]; }
    } else {
      geom->bbox(bbmin, bbmax);
    }
    // cartesiangrid_base.h:125-143
    double ext[3];
    int    decisive = 0;
    for(int d = 0; d < ndim; ++d) {
      ext[d]   = std::abs(bbmax[d] - bbmin[d]);
      decisive = ext[d] > ext[decisive] ? d : decisive;
      cog[d]   = bbmin[d] + 0.5 * (bbmax[d] - bbmin[d]);
    }
    length_on_level[0] = (1.0 + 1.0 / std::pow(2.0, 100.0)) * ext[decisive];
    for(int l = 1; l < 64; ++l) length_on_level[l] = 0.5 * length_on_level[l - 1];
  }

  void generate() {
    const int NC = 1 << ndim, NN = 2 * ndim;
    std::vector<GenCell> cur(1);
    for(int d = 0; d < ndim; ++d) cur[0].center[d] = cog[d];
    cur[0].bndry = true; // cartesiangrid_generation.h:108
    for(int l = 0; l < part_level; ++l) {
      const double len = length_on_level[l + 1];
      std::vector<GenCell> next(cur.size() * NC);
      // refineCell :419-451
      for(size_t p = 0; p < cur.size(); ++p) {
        for(int c = 0; c < NC; ++c) {
          GenCell& ch = next[p * NC + c];
          for(int d = 0; d < ndim; ++d) ch.center[d] = cur[p].center[d] + 0.5 * len * child_dir(c, d);
          ch.parent = static_cast<int64_t>(p);
          if(cur[p].bndry) ch.bndry = geom->cut_with_cell(ch.center, len);
          cur[p].child[c] = static_cast<int64_t>(p * NC + c);
        }
      }
      // findChildLevelNghbrs :453-506
      for(size_t p = 0; p < cur.size(); ++p) {
        for(int c = 0; c < NC; ++c) {
          const int64_t id = cur[p].child[c];
          if(id < 0) continue;
          for(int dir = 0; dir < NN; ++dir) {
            if(next[id].nghbr[dir] != -1) continue;
            const int in = nghbr_inside(c, dir);
            if(in >= 0) {
              next[id].nghbr[dir] = cur[p].child[in];
            } else {
              const int     pc = nghbr_parent_child(c, dir);
              const int64_t pn = cur[p].nghbr[dir];
              if(pn != -1 && pc >= 0 && cur[pn].child[pc] != -1) next[id].nghbr[dir] = cur[pn].child[pc];
            }
          }
        }
      }
      // markOutsideCells :554-580 with floodCells :582-603 (LIFO stack, directions ascending)
      for(size_t i = 0; i < next.size(); ++i) {
        if(next[i].marked) continue;
        next[i].marked = true;
        next[i].inside = next[i].bndry || geom->point_inside(next[i].center);
        if(next[i].bndry) continue;
        const bool inside = next[i].inside;
        std::stack<int64_t> st;
        st.push(static_cast<int64_t>(i));
        while(!st.empty()) {
          const int64_t cc = st.top();
          st.pop();
          for(int dir = 0; dir < NN; ++dir) {
            const int64_t nb = next[cc].nghbr[dir];
            if(nb == -1 || next[nb].marked) continue;
            next[nb].marked = true;
            if(!next[nb].bndry) {
              next[nb].inside = inside;
              st.push(nb);
            } else {
              next[nb].inside = true;
            }
          }
        }
      }
      // deleteOutsideCells :508-552: from the end to the begin, the hole is filled with the current last cell
      int64_t end = static_cast<int64_t>(next.size());
      for(int64_t i = end - 1; i >= 0; --i) {
        if(next[i].inside) continue;
        for(int dir = 0; dir < NN; ++dir) {
          const int64_t nb = next[i].nghbr[dir];
          if(nb != -1) next[nb].nghbr[dir ^ 1] = -1;
        }
        if(next[i].parent >= 0) {
          for(int c = 0; c < NC; ++c)
            if(cur[next[i].parent].child[c] == i) cur[next[i].parent].child[c] = -1;
        }
        if(i != end - 1) {
          // copyCell(end-1 -> i) :616-646
          next[i] = next[end - 1];
          for(int dir = 0; dir < NN; ++dir) {
            const int64_t nb = next[i].nghbr[dir];
            if(nb != -1) next[nb].nghbr[dir ^ 1] = i;
          }
          if(next[i].parent >= 0) {
            for(int c = 0; c < NC; ++c)
              if(cur[next[i].parent].child[c] == end - 1) { cur[next[i].parent].child[c] = i; break; }
          }
        }
        --end;
      }
      next.resize(static_cast<size_t>(end));
      for(GenCell& g : next) g.marked = false;
      cur.swap(next);
    }
    // reorderHilbertCurve :664-718: ascending key of the normalised centre
    const double L0 = length_on_level[0];
    std::vector<int64_t> key(cur.size());
    for(size_t i = 0; i < cur.size(); ++i) {
      double x[3];
      for(int d = 0; d < ndim; ++d) x[d] = ((cur[i].center[d] - cog[d]) + 0.5 * L0) / L0;
      key[i] = lbm::sfc_index_unit(ndim, x, part_level);
    }
    std::vector<int64_t> order(cur.size());
    for(size_t i = 0; i < order.size(); ++i) order[i] = static_cast<int64_t>(i);
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key[a] < key[b]; });
    for(size_t i = 1; i < order.size(); ++i)
      if(key[order[i]] == key[order[i - 1]]) throw std::runtime_error("Duplicated Hilbert Ids found!");
    std::vector<int64_t> newpos(cur.size());
    for(size_t i = 0; i < order.size(); ++i) newpos[order[i]] = static_cast<int64_t>(i);
    cells.resize(cur.size());
    for(size_t i = 0; i < order.size(); ++i) {
      cells[i] = cur[order[i]];
      cells[i].parent = -1; // cartesiangrid_generation.h:143
      for(int dir = 0; dir < NN; ++dir)
        if(cells[i].nghbr[dir] != -1) cells[i].nghbr[dir] = newpos[cells[i].nghbr[dir]];
      for(int c = 0; c < 8; ++c) cells[i].child[c] = -1;
      cells[i].level = part_level;
    }
    // uniformRefineGrid :151-171: every level up to uniformLevel is appended behind the list, the parents stay
    int64_t lb = 0, le = static_cast<int64_t>(cells.size());
    for(int l = part_level; l < uni_level; ++l) refine_block(l, &lb, &le, false);
    // boundary refinement, gridGenerator.cpp:206-210: children of the cut cells of the highest level only
    for(int l = uni_level; l < level; ++l) refine_block(l, &lb, &le, true);
    if(align) transform_to_extent();
  }

 private:
  // refineGrid<true, UNIFORM> (:339-376) of the level block [*lb, *le): refineCell (:419-451) for all / for the cut cells only
  // (refineGridMarkedOnly :407-415), findChildLevelNghbrs (:453-506), deleteOutsideCells (:508-552).  On return [*lb, *le) is the
  // new level's block.
  void refine_block(int l, int64_t* lb, int64_t* le, bool marked_only) {
    const int NC = 1 << ndim, NN = 2 * ndim;
    const double  len = length_on_level[l + 1];
    const int64_t pb = *lb, pe = *le, cb = static_cast<int64_t>(cells.size());
    int64_t k = 0;
    for(int64_t p = pb; p < pe; ++p) {
      if(marked_only && !cells[p].bndry) continue; // markBndryCells :212-225
      cells.resize(cells.size() + NC);
      for(int c = 0; c < NC; ++c) {
        const int64_t id = cb + k * NC + c;
        GenCell&      ch = cells[id];
        ch = GenCell();
        for(int d = 0; d < ndim; ++d) ch.center[d] = cells[p].center[d] + 0.5 * len * child_dir(c, d);
        ch.level  = l + 1;
        ch.parent = p;
        if(cells[p].bndry) ch.bndry = geom->cut_with_cell(ch.center, len);
        cells[p].child[c] = id;
      }
      ++k;
    }
    for(int64_t p = pb; p < pe; ++p) {
      for(int c = 0; c < NC; ++c) {
        const int64_t id = cells[p].child[c];
        if(id < 0) continue;
        for(int dir = 0; dir < NN; ++dir) {
          if(cells[id].nghbr[dir] != -1) continue;
          const int in = nghbr_inside(c, dir);
          if(in >= 0) {
            cells[id].nghbr[dir] = cells[p].child[in];
          } else {
            const int     pc = nghbr_parent_child(c, dir);
            const int64_t pn = cells[p].nghbr[dir];
            if(pn != -1 && pc >= 0 && cells[pn].child[pc] != -1) cells[id].nghbr[dir] = cells[pn].child[pc];
          }
        }
      }
    }
    int64_t end = static_cast<int64_t>(cells.size());
    // markOutsideCells / floodCells :554-603
    for(int64_t i = cb; i < end; ++i) cells[i].marked = false;
    for(int64_t i = cb; i < end; ++i) {
      if(cells[i].marked) continue;
      cells[i].marked = true;
      cells[i].inside = cells[i].bndry || geom->point_inside(cells[i].center);
      if(cells[i].bndry) continue;
      const bool inside = cells[i].inside;
      std::stack<int64_t> st;
      st.push(i);
      while(!st.empty()) {
        const int64_t cc = st.top();
        st.pop();
        for(int dir = 0; dir < NN; ++dir) {
          const int64_t nb = cells[cc].nghbr[dir];
          if(nb == -1 || cells[nb].marked) continue;
          cells[nb].marked = true;
          if(!cells[nb].bndry) {
            cells[nb].inside = inside;
            st.push(nb);
          } else {
            cells[nb].inside = true;
          }
        }
      }
    }
    // deleteCell :527-552 from the end of the block to its begin, the hole is filled with the block's last cell (copyCell :616-646)
    for(int64_t i = end - 1; i >= cb; --i) {
      if(cells[i].inside) continue;
      const int64_t par = cells[i].parent;
      for(int c = 0; c < NC; ++c)
        if(cells[par].child[c] == i) { cells[par].child[c] = -1; break; }
      for(int dir = 0; dir < NN; ++dir) {
        const int64_t nb = cells[i].nghbr[dir];
        if(nb != -1) cells[nb].nghbr[dir ^ 1] = -1;
      }
      if(i != end - 1) {
        cells[i] = cells[end - 1];
        for(int dir = 0; dir < NN; ++dir) {
          const int64_t nb = cells[i].nghbr[dir];
          if(nb != -1) cells[nb].nghbr[dir ^ 1] = i;
        }
        const int64_t mp = cells[i].parent;
        for(int c = 0; c < NC; ++c)
          if(cells[mp].child[c] == end - 1) { cells[mp].child[c] = i; break; }
      }
      --end;
    }
    cells.resize(static_cast<size_t>(end));
    *lb = cb;
    *le = end;
  }

  // transformMaxRfnmtLvlToExtent, cartesiangrid_generation.h:256-304: stretch the cell centres so that the outermost
  // centres lie ON the bounding box in alignDir (all directions for a square domain), scale the cell length, then delete what
  // is now outside (deleteOutsideCells<CHECKALL = true>, :508-525,556-561)
  void transform_to_extent() {
    const int NN = 2 * ndim;
    double emin[3] = {0, 0, 0}, emax[3] = {0, 0, 0};
    for(int d = 0; d < ndim; ++d) {
      emin[d] = std::numeric_limits<double>::max();
      emax[d] = std::numeric_limits<double>::min(); // sic: the smallest positive double, as in the reference
    }
    for(const GenCell& c : cells)
      for(int d = 0; d < ndim; ++d) {
        if(emin[d] > c.center[d]) emin[d] = c.center[d];
        if(emax[d] < c.center[d]) emax[d] = c.center[d];
      }
    bool identical = true;
    for(int d = 0; d < ndim && identical; ++d)
      for(int e = d + 1; e < ndim; ++e)
        if(!(std::abs(emin[d] - emin[e]) < kEps) || !(std::abs(emax[d] - emax[e]) < kEps)) { identical = false; break; }
    const double t = bbmax[align_dir] / (emax[align_dir] - emin[align_dir]);
    for(GenCell& c : cells)
      for(int d = 0; d < ndim; ++d) {
        if(d == align_dir || identical) c.center[d] = c.center[d] * t - (t * emax[d] - bbmax[d]);
        else c.center[d] = c.center[d] * t;
      }
    length_on_level[level] *= t;
    const double len = length_on_level[level];
    for(GenCell& c : cells) {
      c.bndry  = geom->cut_with_cell(c.center, len);
      c.inside = c.bndry || geom->point_inside(c.center);
    }
    int64_t end = static_cast<int64_t>(cells.size());
    for(int64_t i = end - 1; i >= 0; --i) {
      if(cells[i].inside) continue;
      for(int dir = 0; dir < NN; ++dir) {
        const int64_t nb = cells[i].nghbr[dir];
        if(nb != -1) cells[nb].nghbr[dir ^ 1] = -1;
      }
      if(i != end - 1) {
        cells[i] = cells[end - 1];
        for(int dir = 0; dir < NN; ++dir) {
          const int64_t nb = cells[i].nghbr[dir];
          if(nb != -1) cells[nb].nghbr[dir ^ 1] = i;
        }
      }
      --end;
    }
    cells.resize(static_cast<size_t>(end));
  }

 public:
};

struct Surface {
  std::string                            name;
  std::vector<int64_t>                   cells;   // in insertion order, duplicates possible (key "all")
  std::map<int64_t, std::array<double, 3>> normal; // last normal stored for a cell wins (surface.h:52-55)
};

class SolverGrid {
 public:
  int     ndim = 2, nn_axis = 4, nn_diag = 8, max_level = 0;
  int64_t n = 0;
  double  cell_length = 0, bbmin[3] = {0, 0, 0}, bbmax[3] = {0, 0, 0}; // cell_length = lengthOnLvl(maxLvl)
  double  length_on_level[64] = {0};
  std::vector<int> level;        // per cell
  bool    multi_level = false;
  int     part_level = 0;        // partitionLvl; max_level doubles as currentHighestLvl (the generator always reaches maxRfnmtLvl)
  std::vector<int64_t>  nghbr;  // n * nn_diag
  std::vector<double>   center; // n * ndim
  std::vector<uint16_t> props;  // CellProperties bits (gridcell_properties.h:7-26): bndry = bit 4, leaf = bit 14
  std::vector<Surface>  surfaces; // in creation order (geometry name, then key, lexicographic)
  int64_t n_leaf = 0, n_bnd = 0;

  const Surface* find(const std::string& name) const {
    for(const Surface& s : surfaces)
      if(s.name == name) return &s;
    return nullptr;
  }
  int64_t& nb(int64_t c, int dir) { return nghbr[static_cast<size_t>(c) * nn_diag + dir]; }
  int64_t nb(int64_t c, int dir) const { return nghbr[static_cast<size_t>(c) * nn_diag + dir]; }

  static int dir_id(const std::string& s) { // include/common/constants.h:104-125
    static const char* names[6] = {"-x", "+x", "-y", "+y", "-z", "+z"};
    for(int i = 0; i < 6; ++i)
      if(s == names[i]) return i;
    throw std::runtime_error("ERROR: Invalid direction " + s);
  }

  // loadGridInplace, src/cartesiangrid.h:264-320
  void load(const GridGen& g, const Json& solver_cfg) {
    ndim = g.ndim;
    nn_axis = 2 * ndim;
    nn_diag = ndim == 1 ? 2 : (ndim == 2 ? 8 : 26); // cartesian::maxNoNghbrsDiag<NDIM>()
    max_level = g.level;
    n = static_cast<int64_t>(g.cells.size());
    cell_length = g.length_on_level[g.level];
    for(int l = 0; l < 64; ++l) length_on_level[l] = g.length_on_level[l];
    multi_level = g.part_level != g.level;
    part_level  = g.part_level;
    level.resize(static_cast<size_t>(n));
    for(int64_t c = 0; c < n; ++c) level[c] = multi_level ? g.cells[c].level : g.level;
    for(int d = 0; d < ndim; ++d) { bbmin[d] = g.bbmin[d]; bbmax[d] = g.bbmax[d]; }
    nghbr.assign(static_cast<size_t>(n) * nn_diag, -1);
    center.resize(static_cast<size_t>(n) * ndim);
    props.assign(static_cast<size_t>(n), 0);
    for(int64_t c = 0; c < n; ++c) {
      for(int dir = 0; dir < nn_axis; ++dir) nb(c, dir) = g.cells[c].nghbr[dir];
      for(int d = 0; d < ndim; ++d) center[c * ndim + d] = g.cells[c].center[d];
    }
    // setProperties :502-508: leaf = no children
    n_leaf = 0;
    for(int64_t c = 0; c < n; ++c) {
      bool leaf = true;
      for(int k = 0; k < 8; ++k) leaf = leaf && g.cells[c].child[k] < 0;
      if(leaf) { props[c] |= 1u << 14; ++n_leaf; }
    }
    // determineBoundaryCells :510-546: only parent-less cells and children of boundary cells are tested (parents precede their
    // children in the list)
    for(int64_t c = 0; c < n; ++c) {
      const int64_t par = g.cells[c].parent;
      if(par != -1 && !(props[par] & (1u << 4))) continue;
      bool b = g.geom->cut_with_cell(&center[c * ndim], length_on_level[level[c]]);
      if(b) {
        int have = 0;
        for(int dir = 0; dir < nn_axis; ++dir) have += nb(c, dir) != -1;
        if(have == nn_axis) b = false;
      }
      if(b) {
        props[c] |= 1u << 4;
        if(props[c] & (1u << 14)) ++n_bnd;
      }
    }
    // identifyBndrySurfaces :553-603
    const Json& boundary = solver_cfg.at("boundary");
    std::set<std::pair<int64_t, int>> assigned;
    for(const auto& gk : boundary.obj) {
      const size_t nkeys = gk.second.size();
      for(const auto& sk : gk.second.obj) {
        Surface s;
        s.name = nkeys > 1 ? gk.first + "_" + sk.first : gk.first;
        const int d0 = sk.first == "all" ? 0 : dir_id(sk.first);
        const int d1 = sk.first == "all" ? nn_axis : dir_id(sk.first) + 1;
        for(int dir = d0; dir < d1; ++dir) {
          for(int64_t c = 0; c < n; ++c) {
            if(!(props[c] & (1u << 4)) || nb(c, dir) != -1) continue;
            if(!g.geom->cut_with_cell(gk.first, &center[c * ndim], length_on_level[level[c]])) continue;
            if(!assigned.insert({c, dir}).second) continue;
            s.cells.push_back(c);
            std::array<double, 3> nrm = {0, 0, 0};
            nrm[dir / 2] = dir % 2 ? 1.0 : -1.0;
            s.normal[c] = nrm;
          }
        }
        surfaces.push_back(std::move(s));
      }
    }
    // setupPeriodicConnections :608-641 (grid-level periodicity: type periodic with generateBndry false)
    {
      // The reference collects the connections in a std::unordered_map and walks it while erasing the partner of every entry it
      // visits (:611-639).  libstdc++ links a new node at the head of its list, so the walk visits the entries in REVERSE
      // insertion order: for "cube_+x" <-> "cube_-x" the pairing runs as (A = cube_-x, B = cube_+x).  The order only matters on
      // multi-level grids, where one cell finds several partners and the last assignment wins (pinned by the
      // couette_ml_p4u5m7 fixture).
      std::vector<std::pair<std::string, std::string>> conn;
      for(const auto& gk : boundary.obj)
        for(const auto& sk : gk.second.obj) {
          const Json& c = sk.second;
          if(c.opt_str("type", "notset") == "periodic" && !c.opt_bool("generateBndry", true))
            conn.emplace_back(gk.first + "_" + sk.first, c.at("connection").as_string());
        }
      auto has = [&](const std::string& k) {
        for(const auto& kv : conn)
          if(kv.first == k) return true;
        return false;
      };
      std::set<std::string> done;
      for(auto it = conn.rbegin(); it != conn.rend(); ++it) {
        if(done.count(it->first)) continue;
        if(!has(it->second)) throw std::runtime_error("Invalid periodic setup!");
        done.insert(it->second);
        const Surface *a = find(it->first), *b = find(it->second);
        if(a == nullptr || b == nullptr) throw std::runtime_error("Invalid periodic setup!");
        add_periodic(*a, *b);
      }
    }
    // addDiagonalNghbrs :451-493 (2D); 3D: the same composition rule in LBMethod<D3Q27>::m_dirs slot order (extension)
    std::vector<int64_t> axis(static_cast<size_t>(n) * nn_axis);
    for(int64_t c = 0; c < n; ++c)
      for(int dir = 0; dir < nn_axis; ++dir) axis[c * nn_axis + dir] = nb(c, dir);
    auto step = [&](int64_t c, int dir) -> int64_t { return c < 0 ? -1 : axis[c * nn_axis + dir]; };
    for(int64_t c = 0; c < n; ++c) {
      for(int slot = nn_axis; slot < nn_diag; ++slot) {
        int64_t cur = c;
        for(int d = 0; d < ndim; ++d) {
          const int cd = ndim == 2 ? lbm::Lattice<2, 9>::c(slot, d) : lbm::Lattice<3, 27>::c(slot, d);
          if(cd != 0) cur = step(cur, 2 * d + (cd > 0 ? 1 : 0));
        }
        nb(c, slot) = cur;
      }
    }
  }

 private:
  void add_periodic(const Surface& A, const Surface& B) { // :645-706
    for(int64_t a : A.cells) {
      for(int64_t b : B.cells) {
        int mismatch = -1;
        for(int d = 0; d < ndim; ++d) {
          if(std::abs(center[a * ndim + d] - center[b * ndim + d]) > kEps) {
            if(mismatch >= 0) { mismatch = -1; break; }
            mismatch = d;
          }
        }
        if(mismatch < 0) continue;
        const int dir = 2 * mismatch;
        if(center[a * ndim + mismatch] > center[b * ndim + mismatch]) {
          nb(b, dir)     = a;
          nb(a, dir + 1) = b;
        } else {
          nb(a, dir)     = b;
          nb(b, dir + 1) = a;
        }
      }
    }
  }
};

} // namespace lbmhost
