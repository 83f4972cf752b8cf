// partition.hpp -- SFC-range partition of the cell list for one rank of a multi-GPU run (host code, no CUDA).
//
// The reference has no domain decomposition (`partitionLevel`, halo / window cell properties and the
// load-balancing weights are declared but every implementation is a stub, /root/reference/src/cartesiangrid.h:709-710,
// src/loadbalancing_weights.h:6-31).  This is the B200-native design of SURVEY.md section 8e: equal-count contiguous ranges of the
// curve (uniform weights, the reference's only WeightMethod), ghost copies of the remote cells an owned cell pushes to or pulls
// from, per-peer lists of exactly the (cell, direction) populations that cross the cut, and -- where a pressure surface lies within
// two cells of a cut -- the velocity halo of LBMBnd_Pressure's inward neighbours (src/lbm/bnd/bnd_pressure.h:68-84).
// Both sides of every exchange derive their lists from the tables alone, in (global cell id, direction) order, so no set-up
// communication is needed.  Table rows come from a caller-supplied callback (a full table, the synthetic box, or the on-demand
// provider of lbm_b200/host/uniform_grid.hpp), so no rank ever needs the whole table.
// lbm_b200/partition.py is the numpy twin used as the cross-check (tests/test_partition_native.py: identical arrays).
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace lbm {

// rows[r*stride + j] = N(ids[r], j); sources[r*stride + j] = the cell whose push in direction j lands in ids[r]; either may be null
using RowsFn = int (*)(void* user, const int64_t* ids, int64_t n, int64_t* rows, int64_t* sources);

struct PressureSurface {
  const int64_t* cells;   // global ids, application order
  const double*  normals; // [n*ndim]
  int64_t        n;
};

struct Partition {
  int     rank = 0, world = 1, stride = 0, qm = 0;
  int64_t lo = 0, hi = 0;
  std::vector<int64_t> ghosts;                 // global ids, ascending
  std::vector<int64_t> nghbr;                  // [(n_owned + n_ghost) * stride] local ids
  std::vector<int32_t> peers;
  std::vector<int64_t> send_count, recv_count, send_cell, recv_cell;
  std::vector<int32_t> send_dir, recv_dir;
  std::vector<int64_t> vsend_count, vrecv_count, vsend_cell, vrecv_cell;
  std::string error;

  int64_t n_owned() const { return hi - lo; }
  int64_t n_ghost() const { return static_cast<int64_t>(ghosts.size()); }
  int64_t to_local(int64_t g) const {
    if(g < 0) return -1;
    if(g >= lo && g < hi) return g - lo;
    auto it = std::lower_bound(ghosts.begin(), ghosts.end(), g);
    if(it == ghosts.end() || *it != g) return -1;
    return n_owned() + (it - ghosts.begin());
  }
};

inline int64_t partition_bound(int64_t n, int world, int r) { return static_cast<int64_t>((static_cast<__int128>(r) * n) / world); }
inline int partition_owner(int64_t n, int world, int64_t g) {
  // largest r with bound(r) <= g
  int r = static_cast<int>((static_cast<__int128>(g) * world) / n);
  if(r >= world) r = world - 1;
  while(r + 1 < world && partition_bound(n, world, r + 1) <= g) ++r;
  while(r > 0 && partition_bound(n, world, r) > g) --r;
  return r;
}

// LBMBnd_Pressure::apply (bnd_pressure.h:58-66): the inside direction is the axis direction against the first non-zero component
inline int inward_direction(const double* nrm, int ndim) {
  for(int d = 0; d < ndim; ++d) {
    if(nrm[d] < 0) return 2 * d + 1;
    if(nrm[d] > 0) return 2 * d;
  }
  return -1;
}

inline bool build_partition(int64_t N, int ndist, int stride, int ndim, int rank, int world, RowsFn fn, void* user,
                            const std::vector<PressureSurface>& pressure, Partition& P) {
  P = Partition();
  P.rank = rank;
  P.world = world;
  P.stride = stride;
  const int qm = ndist - 1;
  P.qm = qm;
  if(N <= 0 || world < 1 || rank < 0 || rank >= world || stride < qm || fn == nullptr) { P.error = "bad partition arguments"; return false; }
  const int64_t lo = partition_bound(N, world, rank), hi = partition_bound(N, world, rank + 1);
  P.lo = lo;
  P.hi = hi;
  const int64_t no = hi - lo;
  if(no <= 0) { P.error = "a rank would own no cells"; return false; }
  auto owner_of = [&](int64_t g) { return partition_owner(N, world, g); };
  auto mine = [&](int64_t g) { return g >= lo && g < hi; };

  std::vector<int64_t> own(static_cast<size_t>(no));
  for(int64_t k = 0; k < no; ++k) own[k] = lo + k;
  std::vector<int64_t> rows_own(static_cast<size_t>(no) * stride), src_own(static_cast<size_t>(no) * stride);
  if(fn(user, own.data(), no, rows_own.data(), src_own.data()) != 0) { P.error = "row provider failed"; return false; }

  // ---- pressure stencil of every entry of every pressure surface (all ranks see the same global lists)
  std::vector<int64_t> pc, pn1, pn2;
  std::vector<int>     pins;
  for(const PressureSurface& s : pressure) {
    if(s.n <= 0) continue;
    std::vector<int> ins(static_cast<size_t>(s.n));
    for(int64_t k = 0; k < s.n; ++k) {
      ins[k] = inward_direction(s.normals + k * ndim, ndim);
      if(ins[k] < 0) { P.error = "pressure boundary: zero normal"; return false; }
      if(s.cells[k] < 0 || s.cells[k] >= N) { P.error = "pressure boundary: cell id out of range"; return false; }
    }
    std::vector<int64_t> r1(static_cast<size_t>(s.n) * stride), n1(static_cast<size_t>(s.n));
    if(fn(user, s.cells, s.n, r1.data(), nullptr) != 0) { P.error = "row provider failed"; return false; }
    for(int64_t k = 0; k < s.n; ++k) {
      n1[k] = r1[k * stride + ins[k]];
      if(n1[k] < 0) { P.error = "pressure boundary: cell without two inward neighbours"; return false; }
    }
    if(fn(user, n1.data(), s.n, r1.data(), nullptr) != 0) { P.error = "row provider failed"; return false; }
    for(int64_t k = 0; k < s.n; ++k) {
      const int64_t n2 = r1[k * stride + ins[k]];
      if(n2 < 0) { P.error = "pressure boundary: cell without two inward neighbours"; return false; }
      pc.push_back(s.cells[k]);
      pn1.push_back(n1[k]);
      pn2.push_back(n2);
      pins.push_back(ins[k]);
    }
  }
  {
    // the reference applies the entries one after the other and reads m_vars of n1 / n2: a neighbour that an earlier entry has
    // already rewritten makes the result order-dependent (plan.hpp rejects the same thing inside one rank)
    std::unordered_map<int64_t, int64_t> first;
    for(size_t k = 0; k < pc.size(); ++k) first.emplace(pc[k], static_cast<int64_t>(k));
    for(size_t k = 0; k < pc.size(); ++k)
      for(int64_t nb : {pn1[k], pn2[k]}) {
        auto it = first.find(nb);
        if(it != first.end() && it->second < static_cast<int64_t>(k)) {
          P.error = "pressure boundary: inward neighbour is itself a pressure boundary cell (order-dependent in the reference)";
          return false;
        }
      }
  }
  // velocity items (entry, which of n1 / n2) in entry order
  const size_t ni = pc.size() * 2;
  std::vector<int64_t> item_cell(ni), item_nb(ni);
  std::vector<int>     item_cown(ni), item_nown(ni);
  for(size_t k = 0; k < pc.size(); ++k) {
    item_cell[2 * k] = item_cell[2 * k + 1] = pc[k];
    item_nb[2 * k]     = pn1[k];
    item_nb[2 * k + 1] = pn2[k];
  }
  for(size_t i = 0; i < ni; ++i) {
    item_cown[i] = owner_of(item_cell[i]);
    item_nown[i] = owner_of(item_nb[i]);
  }

  // ---- ghosts: remote cells an owned cell pushes to or pulls from, and remote inward neighbours of my pressure cells
  {
    std::vector<int64_t>& g = P.ghosts;
    for(int64_t k = 0; k < no; ++k)
      for(int j = 0; j < qm; ++j) {
        const int64_t a = rows_own[k * stride + j], b = src_own[k * stride + j];
        if(a >= 0 && !mine(a)) g.push_back(a);
        if(b >= 0 && !mine(b)) g.push_back(b);
      }
    for(size_t i = 0; i < ni; ++i)
      if(item_cown[i] == rank && item_nown[i] != rank) g.push_back(item_nb[i]);
    std::sort(g.begin(), g.end());
    g.erase(std::unique(g.begin(), g.end()), g.end());
  }
  const int64_t ng = P.n_ghost(), nl = no + ng;
  P.nghbr.assign(static_cast<size_t>(nl) * stride, -1);
  for(int64_t k = 0; k < no; ++k)
    for(int j = 0; j < qm; ++j) P.nghbr[k * stride + j] = P.to_local(rows_own[k * stride + j]);
  std::vector<int64_t> rows_g(static_cast<size_t>(ng) * stride);
  if(ng > 0) {
    if(fn(user, P.ghosts.data(), ng, rows_g.data(), nullptr) != 0) { P.error = "row provider failed"; return false; }
    for(int64_t k = 0; k < ng; ++k)
      for(int j = 0; j < qm; ++j) {
        const int64_t t = rows_g[k * stride + j];
        P.nghbr[(no + k) * stride + j] = mine(t) ? t - lo : -1; // a ghost row only keeps its links into owned cells
      }
  }
  // send: my cell pushes (direction j) into a cell another rank owns; receive: a ghost pushes into a cell I own
  struct Item { int64_t cell; int32_t dir; int owner; };
  std::vector<Item> snd, rcv;
  for(int64_t k = 0; k < no; ++k)
    for(int j = 0; j < qm; ++j) {
      const int64_t t = rows_own[k * stride + j];
      if(t >= 0 && !mine(t)) snd.push_back({k, j, owner_of(t)});
    }
  for(int64_t k = 0; k < ng; ++k)
    for(int j = 0; j < qm; ++j)
      if(P.nghbr[(no + k) * stride + j] >= 0) rcv.push_back({no + k, j, owner_of(P.ghosts[k])});
  // the pressure boundary condition finds n1 = N(c, inside), n2 = N(n1, inside) in the table: give it the links of the ghost cells it
  // walks over (after the receive list was taken: these links carry no populations)
  for(size_t k = 0; k < pc.size(); ++k) {
    if(item_cown[2 * k] != rank) continue;
    const int64_t l1 = P.to_local(pn1[k]), l2 = P.to_local(pn2[k]);
    if(l1 < 0 || l2 < 0) { P.error = "internal: pressure neighbour missing from the ghost set"; return false; }
    P.nghbr[l1 * stride + pins[k]] = l2;
  }
  // peers
  {
    std::vector<int32_t>& p = P.peers;
    for(const Item& s : snd) p.push_back(s.owner);
    for(const Item& r : rcv) p.push_back(r.owner);
    for(size_t i = 0; i < ni; ++i) {
      if(item_nown[i] == rank && item_cown[i] != rank) p.push_back(item_cown[i]);
      if(item_cown[i] == rank && item_nown[i] != rank) p.push_back(item_nown[i]);
    }
    std::sort(p.begin(), p.end());
    p.erase(std::unique(p.begin(), p.end()), p.end());
  }
  for(int32_t q : P.peers) {
    int64_t cs = 0, cr = 0, vs = 0, vr = 0;
    for(const Item& s : snd)
      if(s.owner == q) { P.send_cell.push_back(s.cell); P.send_dir.push_back(s.dir); ++cs; }
    for(const Item& r : rcv)
      if(r.owner == q) { P.recv_cell.push_back(r.cell); P.recv_dir.push_back(r.dir); ++cr; }
    for(size_t i = 0; i < ni; ++i) {
      if(item_nown[i] == rank && item_cown[i] == q && q != rank) { P.vsend_cell.push_back(item_nb[i] - lo); ++vs; }
    }
    for(size_t i = 0; i < ni; ++i) {
      if(item_cown[i] == rank && item_nown[i] == q && q != rank) { P.vrecv_cell.push_back(P.to_local(item_nb[i])); ++vr; }
    }
    P.send_count.push_back(cs);
    P.recv_count.push_back(cr);
    P.vsend_count.push_back(vs);
    P.vrecv_count.push_back(vr);
  }
  return true;
}

} // namespace lbm
