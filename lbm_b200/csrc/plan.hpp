// plan.hpp -- host-side device plan: turns the reference's tables into the B200 data layout.
//
// Input  (reference semantics): the push table CartesianGrid::neighbor(cell,dir)
//         (/root/reference/src/cartesiangrid.h:111-124), boundary conditions in application order
//         (src/lbm/bnd/bnd.h:71-142), optional forcing (src/lbm/solver.cpp:626-696).
// Output (device semantics): for every population slot fold[c,j] exactly one *link* that says where its value
//         comes from after the reference's passes 6-8 (preApply -> push -> apply) have all run:
//           PULL(src)    the push source (inverse of the push table; the table is NOT assumed symmetric)
//           BB / BB_ADD  bounce back (+ moving-wall addends, added one by one like the reference does)
//           ABB          anti bounce back pressure
//           COPY         periodic boundary condition copy
//           VALUE        a stored number: slots nothing ever writes keep their initial value (the reference never
//                        clears m_fold), periodic-with-pressure slots get a value recomputed every step
//         "last writer wins" is resolved here, once, in the reference's order.
// Layout: cells are regrouped into SFC chunks (aligned cubes of the reference's curve: 8^3 cells in 3D, 32^2 in
//         2D).  Chunks in which every slot is a plain pull that follows the curve's template ("fast" chunks) need
//         no per-cell index at all: one shared template + 3^D neighbour-chunk bases per chunk.  Everything else
//         goes through 32-bit link codes.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

#include "lattice.h"

namespace lbm {

static constexpr double kEps = std::numeric_limits<double>::epsilon(); // GDoubleEps, include/common/sfcmm_types.h:50

enum BcKind { BC_WALL_BB = 1, BC_WALL_BB_TANGENTIAL = 2, BC_DIRICHLET_BB = 3, BC_PRESSURE = 4, BC_PERIODIC = 5,
              BC_WALL_EQ = 6, BC_WALL_NEEM = 7, BC_WALL_NEBB = 8, // 6..8: wet-node walls, handled by sequential.cuh
              BC_POISSON_DIRICHLET = 9, BC_POISSON_NEUMANN = 10 }; // Poisson equation types, handled by poisson.cuh

struct BcInput {
  int                  kind = 0;
  std::vector<int64_t> cells;
  std::vector<double>  normals;   // n * D
  double               value[3] = {0, 0, 0};
  double               tangential = 0;
  double               pressure   = std::numeric_limits<double>::quiet_NaN();
  std::vector<int64_t> connected; // periodic
  bool                 has_velocity = false; // wet-node walls: the "velocity" key is present (value[] holds it)
  std::vector<double>  values;               // Poisson NEEM conditions: m_value per entry
  double               grad = 0;             // Neumann: m_gradValue
};

struct PlanInput {
  LatticeRT            L;
  int64_t              n = 0;
  std::vector<int32_t> nghbr;           // n * (Q-1) push table, compacted from the caller's int64 table
  std::vector<double>  center;          // n*D or empty
  double               bbmin[3] = {0, 0, 0}, bbmax[3] = {0, 0, 0}, cell_length = 0;
  std::vector<BcInput> bcs;
  bool                 forcing = false;
  std::vector<int64_t> inlet, outlet;
  double               gradient = 0;
  // Poisson equation types (solver.cpp EQ == LBEquationType::Poisson): m_dt and poisson_D
  bool                 poisson = false;
  double               poisson_dt = 0, poisson_rate = 0;
  std::vector<int32_t> nghbr_wide;      // D2Q5 only: all 8 columns of the caller's 2D table -- the NEEM conditions extrapolate
                                        // along a DIAGONAL at corners (bnd_dirichlet.h:294-312), which the lattice itself lacks
  // multi-GPU: the last n_ghost cells of the list are copies of cells owned by other ranks (never updated here, their
  // populations arrive by halo exchange); halo lists are (local cell, direction) pairs per peer, already in wire order
  int64_t              n_ghost = 0;
  std::vector<int32_t> peers;
  std::vector<int64_t> send_count, recv_count;
  std::vector<int64_t> send_cell, recv_cell;
  std::vector<int32_t> send_dir, recv_dir;
  // velocity halo of the pressure boundary condition: per peer, owned cells whose velocity is sent / ghost cells that receive one
  std::vector<int64_t> vsend_count, vrecv_count, vsend_cell, vrecv_cell;
};

struct CopySrc { int32_t cell, dir; };
struct AddEntry { double v[3]; int32_t n; int32_t pad; };
// one anti-bounce-back (pressure) boundary entry: u_ext = 1.5 u(n1) - 0.5 u(n2)   (bnd_pressure.h:78-84)
struct AbbEntry { int32_t cell, n1, n2, pad; double p; };
// forcing: f[target,:] = eq(w, p, cu(val), vsq(val)) + f[val,:] - feq[val,:]      (solver.cpp:651-693)
struct ForceEntry { int32_t target, val; double p; };
// periodic with pressure: value[vbase+i] = eq(i,p,u(c)) + f[c,i] - feq[c,i]        (bnd_periodic.h:101-108)
struct PerPEntry { int32_t cell, vbase; double p; };
// m_vars fix-ups done by boundary conditions after the moments pass, resolved to the last writer
struct VarFix { int32_t cell, var, abb, comp; double value; }; // abb >= 0: vars = uext[abb][comp]; else constant

struct Plan {
  LatticeRT L;
  int64_t   n = 0;         // reference cells
  int64_t   npad = 0;      // device cells (stride of the SoA arrays)
  int       CH = 0;
  int64_t   n_fast_chunks = 0, n_slow_chunks = 0, n_loose = 0;
  int64_t   n_ghost_blocks = 0;               // remote chunks mirrored as CH-cell blocks in the ghost range
  int64_t   n_fast_outer = 0, n_gen_outer = 0; // leading fast chunks / generic cells that feed the halo exchange
  int64_t   gen_begin = 0; // first device cell of the generic range
  int64_t   n_gen = 0;     // cells in the generic range
  int64_t   gen_stride = 0;
  std::vector<int32_t>  ref2dev, dev2ref;
  std::vector<uint16_t> tmpl;       // (Q-1) * CH : sel << 10 | off
  std::vector<int32_t>  chunk_nb;   // n_fast_chunks * (NSEL + 1): device bases (-1: wall selector), then the wall descriptor id
  std::vector<AddEntry> wall_desc;  // per wall descriptor: NSEL * (Q-1) addend entries [selector][direction] (bounce-back slots of
                                    // wall chunks); n = -1 marks an anti-bounce-back (pressure) slot, whose entry id comes from chunk_abb
  std::vector<int32_t>  chunk_abb_base; // per fast chunk: row of chunk_abb, -1 if the chunk has no pressure cell
  std::vector<int32_t>  chunk_abb;      // [rows][CH] anti-bounce-back entry of the cell at that (device) offset, -1 elsewhere
  std::vector<int32_t>  codes;      // (Q-1) * gen_stride
  std::vector<CopySrc>  copytab;
  std::vector<AddEntry> addtab;
  std::vector<AbbEntry> abb;
  std::vector<double>   values;     // static part first, then dynamic part
  std::vector<int64_t>  stale_ref;  // per static value k>=1: device cell * Q + dir of the stale slot it backs
  int64_t               n_values_static = 0;
  std::vector<ForceEntry> force;
  std::vector<PerPEntry>  perp;
  std::vector<VarFix>     varfix;
  std::vector<int32_t>    u0_cells; // cells with a preset initial velocity (Dirichlet BB, bnd_dirichlet.h:44-50)
  std::vector<double>     u0_vals;  // D per entry
  int64_t slots_bc = 0, slots_stale = 0;
  int64_t n_owned = 0, n_ghost = 0, ghost_begin = 0; // owned device cells are [0, ghost_begin); the ghosts follow
  // per-direction in-chunk layouts (lattice.h): device cells of chunk-shaped blocks -- fast and slow chunks [0, perm_end), ghost
  // blocks [gb_begin, gb_end) -- are permuted inside their aligned CH-cell block, loose cells are not
  int64_t perm_end = 0, gb_begin = 0, gb_end = 0;
  PermRange perm_range() const { return PermRange{static_cast<int32_t>(perm_end), static_cast<int32_t>(gb_begin), static_cast<int32_t>(gb_end)}; }
  int64_t pop_index(int j, int64_t dev_cell) const { return static_cast<int64_t>(j) * npad + pop_slot(L.lay[j], static_cast<int32_t>(dev_cell), perm_range()); }
  std::vector<int64_t> send_index, recv_index;       // flat device indices dir * npad + cell, wire order
  std::vector<int32_t> vsend_cells;                  // device cells whose velocity travels with the halo, wire order
  int64_t              n_vrecv = 0;                  // velocity items received (3 reals each)
  std::string error;
};

inline bool in_direction(const LatticeRT& L, const double* normal, int dist) {
  double dot = 0; // constants.h:83-86
  for(int d = 0; d < L.D; ++d) dot += normal[d] * L.c[dist][d];
  return dot >= kEps;
}

// offset along the curve inside one chunk <-> local coordinates (see sfc_lut in lattice.h)
inline void chunk_offset_to_xyz(const LatticeRT& L, int off, int* xyz) {
  xyz[0] = xyz[1] = xyz[2] = 0;
  const int bits = L.D;
  for(int l = 0; l < L.CHUNK_LEVELS; ++l) {
    const int digit = (off >> (bits * (L.CHUNK_LEVELS - 1 - l))) & ((1 << bits) - 1);
    const int q     = sfc_lut_inv(digit);
    for(int d = 0; d < L.D; ++d) xyz[d] = (xyz[d] << 1) | ((q >> d) & 1);
  }
}
inline int chunk_xyz_to_offset(const LatticeRT& L, const int* xyz) {
  int off = 0;
  for(int l = 0; l < L.CHUNK_LEVELS; ++l) {
    int q = 0;
    for(int d = 0; d < L.D; ++d) q |= ((xyz[d] >> (L.CHUNK_LEVELS - 1 - l)) & 1) << d;
    off = (off << L.D) | sfc_lut(q);
  }
  return off;
}

// Inside a chunk the DEVICE keeps cells in lexicographic order (x fastest), not in curve order: a warp then reads whole
// aligned rows for every direction without an x component and a row shifted by one cell otherwise, instead of the scattered
// sectors a Morton-like order produces (measured: the curve order inside chunks cost 16 % of the bandwidth).  The chunk as a
// whole still is a run of the reference's curve; only the position of a cell inside its chunk changes.
inline int chunk_xyz_to_lex(const LatticeRT& L, const int* xyz) {
  const int S = 1 << L.CHUNK_LEVELS;
  int off = 0;
  for(int d = L.D - 1; d >= 0; --d) off = off * S + xyz[d];
  return off;
}
inline void chunk_lex_to_xyz(const LatticeRT& L, int off, int* xyz) {
  const int S = 1 << L.CHUNK_LEVELS;
  xyz[0] = xyz[1] = xyz[2] = 0;
  for(int d = 0; d < L.D; ++d) { xyz[d] = off % S; off /= S; }
}

// template entry for slot (offset o, direction j): which neighbour chunk the pull source lies in and where.
// lex = false: offsets along the reference's curve (used to recognise chunks in the reference's list);
// lex = true : offsets in the device's in-chunk order (what the kernel uses)
inline void build_template(const LatticeRT& L, std::vector<uint16_t>& tmpl, bool lex = false) {
  const int S = 1 << L.CHUNK_LEVELS;
  tmpl.assign(static_cast<size_t>(L.Q - 1) * L.CHUNK, 0);
  for(int o = 0; o < L.CHUNK; ++o) {
    int xyz[3];
    if(lex) chunk_lex_to_xyz(L, o, xyz);
    else chunk_offset_to_xyz(L, o, xyz);
    for(int j = 0; j < L.Q - 1; ++j) {
      int src[3] = {0, 0, 0}, sel = 0, mul = 1;
      for(int d = 0; d < L.D; ++d) {
        int v = xyz[d] - L.c[j][d]; // pull: the source sits one step against the direction
        int s = 1;
        if(v < 0) { v += S; s = 0; }
        else if(v >= S) { v -= S; s = 2; }
        src[d] = v;
        sel += s * mul;
        mul *= 3;
      }
      tmpl[static_cast<size_t>(j) * L.CHUNK + o] =
          static_cast<uint16_t>((sel << 10) | (lex ? chunk_xyz_to_lex(L, src) : chunk_xyz_to_offset(L, src)));
    }
  }
}

// LBMBndCell_periodic::init (bnd_periodic.h:31-98) for entry k of a periodic boundary condition: the outward directions whose
// tangentially shifted target stays inside the bounding box, and for each the first cell of the connected surface whose
// centre matches the shifted centre in ANY coordinate.  Returns false (with *err) where the reference would abort.
inline bool periodic_links(const PlanInput& in, const BcInput& bc, int64_t k, int* setd, int64_t* links, int* nset, std::string* err) {
  const LatticeRT& L = in.L;
  const int D = L.D, Q = L.Q;
  const double  maxMatch = 10 * kEps;
  const int64_t c   = bc.cells[k];
  const double* nrm = &bc.normals[k * D];
  const double* ctr = &in.center[c * D];
  int ns = 0;
  for(int dist = 0; dist < Q; ++dist) {
    if(!in_direction(L, nrm, dist)) continue;
    bool inside = true;
    for(int d = 0; d < D; ++d) {
      const double x = std::abs(nrm[d]) > 0 ? in.bbmin[d] : ctr[d] + L.c[dist][d] * in.cell_length;
      if(x < in.bbmin[d] || x > in.bbmax[d]) inside = false;
    }
    if(inside) setd[ns++] = dist;
  }
  for(int id = 0; id < ns; ++id) {
    double ca[3];
    for(int d = 0; d < D; ++d) ca[d] = std::abs(nrm[d]) > 0 ? ctr[d] : ctr[d] + L.c[setd[id]][d] * in.cell_length;
    int64_t link = -1;
    for(size_t q = 0; q < bc.connected.size() && link < 0; ++q) {
      const double* cb = &in.center[bc.connected[q] * D];
      for(int d = 0; d < D; ++d)
        if(std::abs(ca[d] - cb[d]) <= maxMatch) { link = bc.connected[q]; break; }
    }
    if(link < 0) { *err = "periodic boundary: no cell to link"; return false; }
    links[id] = link;
  }
  if(ns == 0) { *err = "periodic boundary cell without outward direction"; return false; }
  for(int a = 0; a < ns; ++a)
    for(int b = a + 1; b < ns; ++b)
      if(links[a] == links[b]) { *err = "Invalid periodic bnd (cell has been linked twice)"; return false; } // bnd_periodic.h:158-166
  *nset = ns;
  return true;
}

struct SlotDesc {
  int     kind;      // LinkKind, or -1 for "dynamic value of periodic-with-pressure"
  int64_t a = 0;     // COPY: src cell (ref id); ABB: entry id; VALUE(dyn): value index
  int     b = 0;     // COPY: dir
  double  add[3] = {0, 0, 0};
  int     nadd = 0;
};

inline int self_sel(const LatticeRT& L) { return L.D == 2 ? 4 : 13; }

inline bool build_plan(const PlanInput& in, Plan& P) {
  const LatticeRT& L = in.L;
  const int     Q = L.Q, D = L.D, QM = L.Q - 1, CH = L.CHUNK;
  const int64_t N = in.n;
  P = Plan();
  P.L = L;
  P.n = N;
  P.CH = CH;
  if(N <= 0 || in.nghbr.empty()) { P.error = "no topology set"; return false; }
  if(N >= (int64_t(1) << 28) * 8) { P.error = "too many cells for 32-bit device indices"; return false; }
  auto NB = [&](int64_t c, int j) -> int64_t { return in.nghbr[static_cast<size_t>(c) * QM + j]; };

  // ---- 1. pull table = inverse of the push table (serial order: the highest source wins, like the
  //         reference's loop would leave it if two cells pushed to one slot)
  std::vector<int32_t> pull(static_cast<size_t>(N) * QM, -1);
#pragma omp parallel for schedule(static)
  for(int j = 0; j < QM; ++j) {
    for(int64_t s = 0; s < N; ++s) {
      const int64_t t = NB(s, j);
      if(t >= 0 && t < N) pull[static_cast<size_t>(t) * QM + j] = static_cast<int32_t>(s);
    }
  }

  for(const BcInput& bc : in.bcs)
    if(bc.kind >= BC_WALL_EQ) { P.error = "internal: wet-node walls and Poisson conditions are handled by the sequential paths"; return false; }
  if(in.poisson) { P.error = "internal: the Poisson equation has no fused device plan"; return false; }
  // forcing and the periodic boundary condition rebuild m_fold of OTHER cells (x-neighbours, linked cells), which a partition cut
  // may hand to another rank: not partitioned (the value cells would be ghosts without links of their own)
  if(in.n_ghost > 0) {
    if(in.forcing) { P.error = "forcing is not partitioned"; return false; }
    for(const BcInput& bc : in.bcs)
      if(bc.kind == BC_PERIODIC) { P.error = "the periodic boundary condition is not partitioned (use grid-level periodic links)"; return false; }
  }
  // ---- 2. boundary conditions, in the reference's order: preApply writes, then the push, then apply writes
  std::unordered_map<int64_t, SlotDesc> over;  // slot key c*Q+j -> final descriptor
  auto key = [&](int64_t c, int j) { return c * Q + j; };
  std::vector<char> is_bc_cell(static_cast<size_t>(N), 0);

  // 2a. preApply (periodic). A slot that also has a push source is overwritten by the push afterwards.
  for(const BcInput& bc : in.bcs) {
    if(bc.kind != BC_PERIODIC) continue;
    if(in.center.empty()) { P.error = "periodic boundary condition needs set_geometry"; return false; }
    const int64_t nb = static_cast<int64_t>(bc.cells.size());
    for(int64_t k = 0; k < nb; ++k) {
      const int64_t c = bc.cells[k];
      int     setd[27], ns = 0;
      int64_t links[27];
      if(!periodic_links(in, bc, k, setd, links, &ns, &P.error)) return false;
      if(!std::isnan(bc.pressure)) {
        // bnd_periodic.h:101-108: all Q populations of the first linked cell get a recomputed value
        PerPEntry e;
        e.cell  = static_cast<int32_t>(c); // ref id for now
        e.vbase = static_cast<int32_t>(P.perp.size()) * Q; // dynamic value index, rebased below
        e.p     = bc.pressure;
        P.perp.push_back(e);
        for(int i = 0; i < Q; ++i) {
          SlotDesc sd;
          sd.kind = -1;
          sd.a    = e.vbase + i;
          over[key(links[0], i)] = sd;
        }
        P.varfix.push_back(VarFix{static_cast<int32_t>(links[0]), D, -1, 0, bc.pressure});
      } else {
        for(int id = 0; id < ns; ++id) {
          SlotDesc sd;
          sd.kind = LK_COPY;
          sd.a    = c;
          sd.b    = setd[id];
          over[key(links[id], setd[id])] = sd;
        }
        P.varfix.push_back(VarFix{static_cast<int32_t>(links[0]), D, -1, 0, 1.0});
      }
    }
  }
  // preApply of the pressure BC only sets rho (bnd_pressure.h:43-50); apply sets it again to the same value.
  // 2b. the push overrides preApply writes
  for(auto it = over.begin(); it != over.end();) {
    const int64_t c = it->first / Q;
    const int     j = static_cast<int>(it->first % Q);
    if(j < QM && pull[static_cast<size_t>(c) * QM + j] >= 0) it = over.erase(it);
    else ++it;
  }
  // the rest population is "pushed" onto itself (solver.cpp:737): fold[c,Q-1] = f[c,Q-1] always wins over preApply
  for(auto it = over.begin(); it != over.end();) {
    if(static_cast<int>(it->first % Q) == QM) it = over.erase(it);
    else ++it;
  }

  // 2c. apply, in order
  std::vector<char> abb_cell(static_cast<size_t>(N), 0);
  for(const BcInput& bc : in.bcs) {
    const int64_t nb = static_cast<int64_t>(bc.cells.size());
    if(bc.kind == BC_WALL_BB_TANGENTIAL && D != 2) {
      P.error = "tangential wall velocity is implemented for 2D only (reference: bnd_wall.h:52-54)";
      return false;
    }
    for(int64_t k = 0; k < nb; ++k) {
      const int64_t c = bc.cells[k];
      if(c < 0 || c >= N) { P.error = "boundary cell id out of range"; return false; }
      const double* nrm = &bc.normals[k * D];
      if(bc.kind == BC_PERIODIC) continue;
      int abb_id = -1;
      if(bc.kind == BC_PRESSURE) {
        // bnd_pressure.h:58-84
        int ins = -1;
        for(int d = 0; d < D && ins < 0; ++d) {
          if(nrm[d] < 0) ins = 2 * d + 1;
          else if(nrm[d] > 0) ins = 2 * d;
        }
        if(ins < 0) { P.error = "pressure boundary: zero normal"; return false; }
        const int64_t n1 = NB(c, ins);
        const int64_t n2 = n1 >= 0 ? NB(n1, ins) : -1;
        if(n1 < 0 || n2 < 0) { P.error = "pressure boundary: cell without two inward neighbours"; return false; }
        // the reference applies entries one after the other and reads m_vars of n1/n2: if one of them had its
        // velocity rewritten by an earlier pressure entry of the same pass the result depends on that order
        if(abb_cell[n1] || abb_cell[n2]) {
          P.error = "pressure boundary: inward neighbour is itself a pressure boundary cell (order-dependent in the reference)";
          return false;
        }
        AbbEntry e;
        e.cell = static_cast<int32_t>(c);
        e.n1   = static_cast<int32_t>(n1);
        e.n2   = static_cast<int32_t>(n2);
        e.pad  = 0;
        e.p    = bc.pressure;
        abb_id = static_cast<int>(P.abb.size());
        P.abb.push_back(e);
        P.varfix.push_back(VarFix{static_cast<int32_t>(c), D, -1, 0, bc.pressure});
        for(int d = 0; d < D; ++d) P.varfix.push_back(VarFix{static_cast<int32_t>(c), d, abb_id, d, 0.0});
      }
      for(int i = 0; i < QM; ++i) {
        if(NB(c, i) != -1 || !in_direction(L, nrm, i)) continue; // bnd_dirichlet.h:86-88
        const int op = L.opp[i];
        SlotDesc  sd;
        switch(bc.kind) {
          case BC_WALL_BB: sd.kind = LK_BB; break;
          case BC_WALL_BB_TANGENTIAL: {
            // bnd_wall.h:31-72: value[inside dir] = u_t * (t . c_inside), t = (n_y, n_x); bnd_dirichlet.h:111-112
            const double t[2] = {nrm[1], nrm[0]};
            const double tdot = t[0] * L.c[op][0] + t[1] * L.c[op][1];
            const double ndot = nrm[0] * L.c[op][0] + nrm[1] * L.c[op][1];
            const double nn   = std::sqrt(double(L.c[op][0] * L.c[op][0] + L.c[op][1] * L.c[op][1]));
            const bool parallel = std::abs(std::acos(ndot / nn) - 3.14159265358979323846) < 10 * kEps;
            const double val    = parallel ? 0.0 : bc.tangential * tdot;
            sd.kind   = LK_BB_ADD;
            sd.add[0] = 1.0 * 2.0 / (1.0 / 3.0) * L.w[op] * val;
            sd.nadd   = 1;
            break;
          }
          case BC_DIRICHLET_BB: {
            // bnd_dirichlet.h:113-117: one addition per dimension, in order
            sd.kind = LK_BB_ADD;
            sd.nadd = D;
            for(int d = 0; d < D; ++d) sd.add[d] = 1.0 * 2.0 / (1.0 / 3.0) * L.w[op] * L.c[op][d] * bc.value[d];
            break;
          }
          case BC_PRESSURE: sd.kind = LK_ABB; sd.a = abb_id; break;
          default: P.error = "unknown boundary kind"; return false;
        }
        over[key(c, op)] = sd;
      }
      if(bc.kind == BC_PRESSURE) abb_cell[c] = 1;
      if(bc.kind == BC_DIRICHLET_BB) {
        P.u0_cells.push_back(static_cast<int32_t>(c));
        for(int d = 0; d < D; ++d) P.u0_vals.push_back(bc.value[d]);
      }
    }
  }
  // a pressure entry that reads the velocity of a cell which a LATER entry rewrites is fine (it reads the raw
  // moments, as the reference does); an EARLIER one was rejected above.

  // ---- 3. forcing pairs (solver.cpp:651-693), resolved once instead of an O(N_in*N_out) search per step
  if(in.forcing) {
    if(in.center.empty()) { P.error = "forcing needs set_geometry"; return false; }
    std::unordered_map<int32_t, size_t> written;
    auto add_force = [&](int64_t target, int64_t val, double p) -> bool {
      if(val < 0) { P.error = "forcing: boundary cell without x-neighbour"; return false; }
      auto it = written.find(static_cast<int32_t>(target));
      ForceEntry e{static_cast<int32_t>(target), static_cast<int32_t>(val), p};
      if(it != written.end()) P.force[it->second] = e; // a later match overwrites an earlier one
      else { written[static_cast<int32_t>(target)] = P.force.size(); P.force.push_back(e); }
      return true;
    };
    for(int64_t a : in.inlet) {
      const int64_t val = NB(a, 1);
      if(val < 0) { P.error = "forcing: inlet cell without +x neighbour"; return false; }
      for(int64_t b : in.outlet)
        if(std::abs(in.center[val * D + 1] - in.center[b * D + 1]) < kEps)
          if(!add_force(b, val, 1.0)) return false;
    }
    for(int64_t b : in.outlet) {
      const int64_t val = NB(b, 0);
      if(val < 0) { P.error = "forcing: outlet cell without -x neighbour"; return false; }
      for(int64_t a : in.inlet)
        if(std::abs(in.center[a * D + 1] - in.center[val * D + 1]) < kEps)
          if(!add_force(a, val, 1.0 + in.gradient)) return false;
    }
    for(const ForceEntry& e : P.force)
      if(written.count(e.val)) { P.error = "forcing: value cell is itself a forced cell (order-dependent in the reference)"; return false; }
  }

  // ---- 4. classify cells: plain = every slot is a pull from an owned cell
  const int64_t NO = N - in.n_ghost; // owned cells come first, ghosts last
  if(NO <= 0) { P.error = "no owned cells"; return false; }
  P.n_owned = NO;
  P.n_ghost = in.n_ghost;
  std::vector<char> plain(static_cast<size_t>(N), 1);
#pragma omp parallel for schedule(static)
  for(int64_t c = 0; c < N; ++c) {
    if(c >= NO) { plain[c] = 0; continue; }
    for(int j = 0; j < QM; ++j) {
      const int32_t ps = pull[static_cast<size_t>(c) * QM + j];
      if(ps < 0 || ps >= NO) { plain[c] = 0; break; }
    }
  }
  for(const auto& kv : over) plain[kv.first / Q] = 0;

  // ---- 5. SFC chunks: runs of CH consecutive cells that are internally ordered like the curve template
  std::vector<uint16_t> tmpl_sfc;
  build_template(L, tmpl_sfc, false); // reference (curve) order: chunk recognition
  build_template(L, P.tmpl, true);    // device order: what the kernels read -- offsets in each direction's own in-chunk layout
  for(int j = 0; j < QM; ++j)
    for(int o = 0; o < CH; ++o) {
      uint16_t& t = P.tmpl[static_cast<size_t>(j) * CH + o];
      t = static_cast<uint16_t>((t & ~1023) | lay_perm(L.lay[j], t & 1023));
    }
  std::vector<int32_t> sfc2lex(static_cast<size_t>(CH));
  for(int o = 0; o < CH; ++o) {
    int xyz[3];
    chunk_offset_to_xyz(L, o, xyz);
    sfc2lex[o] = chunk_xyz_to_lex(L, xyz);
  }
  const int SELF = self_sel(L);
  std::vector<int64_t> cand_base; // reference index of the first cell of every candidate chunk
  std::vector<int32_t> chunk_of(static_cast<size_t>(N), -1);
  {
    int64_t b = 0;
    while(b + CH <= NO) {
      bool ok = true;
      for(int o = 0; o < CH && ok; ++o) {
        for(int j = 0; j < QM; ++j) {
          const uint16_t t = tmpl_sfc[static_cast<size_t>(j) * CH + o];
          if((t >> 10) != SELF) continue;
          if(pull[static_cast<size_t>(b + o) * QM + j] != b + (t & 1023)) { ok = false; break; }
        }
      }
      if(ok) {
        for(int o = 0; o < CH; ++o) chunk_of[b + o] = static_cast<int32_t>(cand_base.size());
        cand_base.push_back(b);
        b += CH;
      } else {
        b += 1;
      }
    }
  }
  const int64_t nc = static_cast<int64_t>(cand_base.size());
  // fast chunk: every slot is either a pull that lands in a candidate chunk at the template offset, or -- for all slots
  // that would pull from one and the same (missing) neighbour chunk -- a bounce-back written by a wall boundary condition,
  // with the same addends for every cell of the chunk in a given direction (no-slip wall, moving lid).  Wall chunks of a
  // box therefore need no per-cell index either; chunks on edges where two different walls meet stay generic.
  std::vector<char> odd(static_cast<size_t>(N), 0); // a BC overrides a slot that also has a push source (asymmetric tables)
  for(const auto& kv : over) {
    const int64_t c = kv.first / Q;
    const int     j = static_cast<int>(kv.first % Q);
    if(kv.second.kind != LK_BB && kv.second.kind != LK_BB_ADD && kv.second.kind != LK_ABB) odd[c] = 1;
    else if(j < QM && pull[static_cast<size_t>(c) * QM + j] >= 0) odd[c] = 1;
  }
  // ---- ghost blocks (partitioned runs): a chunk at a partition cut pulls from ghost cells.  If all ghosts that one (chunk,
  // selector) pair needs sit at consistent template offsets they are the visible part of ONE remote chunk: that chunk gets
  // a CH-cell block of its own in the ghost range, its ghosts are stored at their template offsets, and the cut chunk stays
  // on the index-free path (the block is mostly padding; only the cells that cross the cut are ever filled by the halo).
  std::vector<int32_t> ghost_group(static_cast<size_t>(in.n_ghost), -1); // ghost (index c - NO) -> block
  std::vector<int32_t> ghost_off(static_cast<size_t>(in.n_ghost), -1);   // its curve offset inside the block
  std::vector<int32_t> key_group;                                        // (chunk k, selector) -> block or -1, size nc * NSEL
  int32_t n_ghost_blocks = 0;
  if(in.n_ghost > 0 && nc > 0) {
    const size_t nkeys = static_cast<size_t>(nc) * L.NSEL;
    std::vector<int32_t> parent(nkeys);
    for(size_t i = 0; i < nkeys; ++i) parent[i] = static_cast<int32_t>(i);
    std::function<int32_t(int32_t)> find = [&](int32_t x) {
      while(parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
      return x;
    };
    std::vector<char>    key_used(nkeys, 0), key_bad(nkeys, 0);
    std::vector<int32_t> ghost_key(static_cast<size_t>(in.n_ghost), -1);
    for(int64_t k = 0; k < nc; ++k) {
      const int64_t b = cand_base[k];
      for(int o = 0; o < CH; ++o) {
        for(int j = 0; j < QM; ++j) {
          const int64_t src = pull[static_cast<size_t>(b + o) * QM + j];
          if(src < NO) continue;
          const uint16_t t   = tmpl_sfc[static_cast<size_t>(j) * CH + o];
          const int32_t  ky  = static_cast<int32_t>(k * L.NSEL + (t >> 10));
          const int32_t  off = t & 1023;
          const size_t   g   = static_cast<size_t>(src - NO);
          key_used[ky] = 1;
          if((t >> 10) == SELF) key_bad[ky] = 1;
          if(ghost_off[g] == -1) { ghost_off[g] = off; ghost_key[g] = ky; }
          else {
            if(ghost_off[g] != off) key_bad[ky] = key_bad[ghost_key[g]] = 1;
            const int32_t ra = find(ky), rb = find(ghost_key[g]);
            if(ra != rb) parent[ra] = rb;
          }
        }
      }
    }
    // a block is valid if no key in it is bad and no two ghosts claim the same offset
    std::vector<char> root_bad(nkeys, 0);
    for(size_t i = 0; i < nkeys; ++i)
      if(key_used[i] && key_bad[i]) root_bad[find(static_cast<int32_t>(i))] = 1;
    {
      std::unordered_map<int64_t, int32_t> taken; // (root, offset) -> ghost
      for(size_t g = 0; g < ghost_off.size(); ++g) {
        if(ghost_key[g] < 0) continue;
        const int32_t r  = find(ghost_key[g]);
        const int64_t kk = static_cast<int64_t>(r) * 1024 + ghost_off[g];
        auto it = taken.find(kk);
        if(it == taken.end()) taken[kk] = static_cast<int32_t>(g);
        else root_bad[r] = 1;
      }
    }
    std::vector<int32_t> root_block(nkeys, -1);
    key_group.assign(nkeys, -1);
    for(size_t i = 0; i < nkeys; ++i) {
      if(!key_used[i]) continue;
      const int32_t r = find(static_cast<int32_t>(i));
      if(root_bad[r]) continue;
      if(root_block[r] < 0) root_block[r] = n_ghost_blocks++;
      key_group[i] = root_block[r];
    }
    for(size_t g = 0; g < ghost_off.size(); ++g) {
      if(ghost_key[g] < 0) continue;
      const int32_t r = find(ghost_key[g]);
      ghost_group[g] = root_bad[r] ? -1 : root_block[r];
    }
  }
  const int64_t GHOSTBLOCK = -1000; // nbref marker: GHOSTBLOCK - block id

  const int64_t WALL = -2;
  std::vector<char>    fast(static_cast<size_t>(nc), 0);
  std::vector<int64_t> nbref(static_cast<size_t>(nc) * L.NSEL, -1);
  std::vector<int32_t> wall_of(static_cast<size_t>(nc), -1);      // wall descriptor per candidate chunk
  std::vector<std::vector<AddEntry>> chunk_wall(static_cast<size_t>(nc));
  // pressure (anti-bounce-back) chunks: a chunk on a pressure in-/outlet face stays index-free as well -- every slot that would
  // pull from the missing neighbour chunk is an anti-bounce-back slot of the cell's pressure entry (curve offset -> entry id)
  std::vector<std::vector<int32_t>> chunk_abb_ids(static_cast<size_t>(nc));
#pragma omp parallel for schedule(dynamic, 64)
  for(int64_t k = 0; k < nc; ++k) {
    const int64_t b = cand_base[k];
    bool ok = true, has_wall = false;
    int64_t* nbk = &nbref[static_cast<size_t>(k) * L.NSEL];
    nbk[SELF] = b;
    // one descriptor entry per (missing neighbour chunk, direction): a chunk on an EDGE of the domain bounces some cells off one wall
    // and some off the other in the same direction, but which wall it is follows from the selector of the missing source
    const int             WD = L.NSEL * QM;
    std::vector<AddEntry> wd(static_cast<size_t>(WD));
    std::vector<char>     wd_set(static_cast<size_t>(WD), 0);
    std::vector<int32_t>  abb_of;
    for(int o = 0; o < CH && ok; ++o) {
      const int64_t c = b + o;
      if(odd[c]) { ok = false; break; }
      for(int j = 0; j < QM; ++j) {
        const uint16_t t   = tmpl_sfc[static_cast<size_t>(j) * CH + o];
        const int      sel = t >> 10;
        const int64_t  src = pull[static_cast<size_t>(c) * QM + j];
        if(src < 0) {
          if(plain[c] || sel == SELF) { ok = false; break; }
          auto it = over.find(key(c, j)); // read-only lookups: safe from several threads
          if(it == over.end()) { ok = false; break; } // stale slot
          AddEntry e{};
          if(it->second.kind == LK_ABB) {
            e.n = -1;
            if(abb_of.empty()) abb_of.assign(static_cast<size_t>(CH), -1);
            const int32_t id = static_cast<int32_t>(it->second.a);
            if(abb_of[o] != -1 && abb_of[o] != id) { ok = false; break; } // a corner cell of two pressure surfaces: generic
            abb_of[o] = id;
          } else {
            e.n = it->second.kind == LK_BB_ADD ? it->second.nadd : 0;
          }
          for(int d = 0; d < 3; ++d) e.v[d] = d < e.n ? it->second.add[d] : 0.0;
          const int wj = sel * QM + j;
          if(!wd_set[wj]) { wd[wj] = e; wd_set[wj] = 1; }
          else if(wd[wj].n != e.n || std::memcmp(wd[wj].v, e.v, sizeof(e.v)) != 0) { ok = false; break; }
          if(nbk[sel] == -1) nbk[sel] = WALL;
          else if(nbk[sel] != WALL) { ok = false; break; }
          has_wall = true;
          continue;
        }
        if(src >= NO) { // pull from a ghost cell: fine if it sits in a ghost block at the template offset
          const int32_t grp = key_group.empty() ? -1 : key_group[static_cast<size_t>(k) * L.NSEL + sel];
          if(grp < 0) { ok = false; break; }
          if(nbk[sel] == -1) nbk[sel] = GHOSTBLOCK - grp;
          else if(nbk[sel] != GHOSTBLOCK - grp) { ok = false; break; }
          continue;
        }
        const int64_t nb0 = src - (t & 1023);
        if(nbk[sel] == -1) {
          if(nb0 < 0 || nb0 + CH > NO || chunk_of[nb0] < 0 || cand_base[chunk_of[nb0]] != nb0) { ok = false; break; }
          nbk[sel] = nb0;
        } else if(nbk[sel] != nb0) { ok = false; break; }
      }
    }
    fast[k] = ok ? 1 : 0;
    if(ok && has_wall) {
      for(int j = 0; j < WD; ++j) if(!wd_set[j]) { wd[j] = AddEntry{}; }
      chunk_wall[k] = wd;
      chunk_abb_ids[k] = abb_of;
    }
  }
  // deduplicate wall descriptors (a box has a handful)
  for(int64_t k = 0; k < nc; ++k) {
    if(!fast[k] || chunk_wall[k].empty()) continue;
    int32_t id = -1;
    const size_t WD = static_cast<size_t>(L.NSEL) * QM;
    for(size_t w = 0; w < P.wall_desc.size() / WD && id < 0; ++w) {
      bool same = true;
      for(size_t j = 0; j < WD && same; ++j) {
        const AddEntry& a = P.wall_desc[w * WD + j];
        const AddEntry& c2 = chunk_wall[k][j];
        same = a.n == c2.n && std::memcmp(a.v, c2.v, sizeof(a.v)) == 0;
      }
      if(same) id = static_cast<int32_t>(w);
    }
    if(id < 0) {
      id = static_cast<int32_t>(P.wall_desc.size() / WD);
      P.wall_desc.insert(P.wall_desc.end(), chunk_wall[k].begin(), chunk_wall[k].end());
    }
    wall_of[k] = id;
  }

  // ---- 6. device layout: fast chunks, slow chunks (kept contiguous so fast chunks can address them by
  //         template), loose cells; all in reference (SFC) order within their group.  In a partitioned run the "outer"
  //         units -- chunks / loose cells holding a cell whose populations a peer needs -- come first in their group, so
  //         that they can be updated by a first launch and exchanged while a second launch updates the rest.
  std::vector<char> is_send(static_cast<size_t>(N), 0);
  for(int64_t c : in.send_cell)
    if(c >= 0 && c < NO) is_send[c] = 1;
  std::vector<char> outer_chunk(static_cast<size_t>(nc), 0);
  for(int64_t k = 0; k < nc; ++k)
    for(int o = 0; o < CH; ++o)
      if(is_send[cand_base[k] + o]) { outer_chunk[k] = 1; break; }
  P.ref2dev.assign(static_cast<size_t>(N), -1);
  std::vector<int64_t> cand_dev(static_cast<size_t>(nc), -1);
  std::vector<int64_t> fast_order;
  int64_t pos = 0;
  for(int pass = 0; pass < 2; ++pass)
    for(int64_t k = 0; k < nc; ++k)
      if(fast[k] && (outer_chunk[k] != 0) == (pass == 0)) {
        cand_dev[k] = pos;
        pos += CH;
        fast_order.push_back(k);
        if(pass == 0) ++P.n_fast_outer;
      }
  P.n_fast_chunks = static_cast<int64_t>(fast_order.size());
  P.gen_begin = pos;
  // slow chunks first (all of them count as outer: they are few, and keeping every chunk-shaped block in one leading range makes
  // "is this cell stored in the per-direction layouts?" a single comparison), then the loose cells, outer ones first
  for(int64_t k = 0; k < nc; ++k)
    if(!fast[k]) { cand_dev[k] = pos; pos += CH; ++P.n_slow_chunks; }
  P.perm_end = pos;
  for(int pass = 0; pass < 2; ++pass) {
    for(int64_t c = 0; c < NO; ++c)
      if(chunk_of[c] < 0 && (is_send[c] != 0) == (pass == 0)) { P.ref2dev[c] = static_cast<int32_t>(pos++); ++P.n_loose; }
    if(pass == 0) P.n_gen_outer = in.send_cell.empty() ? 0 : pos - P.gen_begin;
  }
  for(int64_t k = 0; k < nc; ++k)
    for(int o = 0; o < CH; ++o) P.ref2dev[cand_base[k] + o] = static_cast<int32_t>(cand_dev[k] + sfc2lex[o]);
  P.n_gen  = pos - P.gen_begin;
  P.ghost_begin = pos;
  if(n_ghost_blocks > 0) pos = (pos + CH - 1) / CH * CH; // ghost blocks are chunk-shaped: aligned like chunks
  const int64_t ghost_block_base = pos;
  P.gb_begin = pos;
  pos += static_cast<int64_t>(n_ghost_blocks) * CH;
  P.gb_end = pos;
  P.n_ghost_blocks = n_ghost_blocks;
  for(int64_t c = NO; c < N; ++c) {
    const size_t g = static_cast<size_t>(c - NO);
    if(ghost_group[g] >= 0) P.ref2dev[c] = static_cast<int32_t>(ghost_block_base + static_cast<int64_t>(ghost_group[g]) * CH + sfc2lex[ghost_off[g]]);
    else P.ref2dev[c] = static_cast<int32_t>(pos++);
  }
  P.npad   = (pos + CH - 1) / CH * CH; // whole blocks: the in-chunk permutation never leaves the arrays
  if(static_cast<int64_t>(Q) * P.npad >= (int64_t(1) << 32)) { P.error = "too many populations for 32-bit element offsets (Q * cells >= 2^32)"; return false; }
  P.gen_stride = (P.n_gen + 63) / 64 * 64;
  P.dev2ref.assign(static_cast<size_t>(P.npad), -1);
  for(int64_t c = 0; c < N; ++c) P.dev2ref[P.ref2dev[c]] = static_cast<int32_t>(c);

  const int NBW = L.NSEL + 1;
  P.chunk_nb.assign(static_cast<size_t>(P.n_fast_chunks) * NBW, 0);
  for(size_t f = 0; f < fast_order.size(); ++f) {
    const int64_t k = fast_order[f];
    for(int s = 0; s < L.NSEL; ++s) {
      const int64_t r = nbref[static_cast<size_t>(k) * L.NSEL + s];
      // a selector no slot uses (e.g. cube corners for D3Q19) stays at the chunk itself
      int32_t v = static_cast<int32_t>(cand_dev[k]);
      if(r == WALL) v = -1;
      else if(r <= GHOSTBLOCK) v = static_cast<int32_t>(ghost_block_base + (GHOSTBLOCK - r) * CH);
      else if(r >= 0) v = static_cast<int32_t>(cand_dev[chunk_of[r]]);
      P.chunk_nb[f * NBW + s] = v;
    }
    P.chunk_nb[f * NBW + L.NSEL] = wall_of[k];
  }
  P.chunk_abb_base.assign(static_cast<size_t>(P.n_fast_chunks), -1);
  for(size_t f = 0; f < fast_order.size(); ++f) {
    const std::vector<int32_t>& ids = chunk_abb_ids[fast_order[f]];
    if(ids.empty()) continue;
    P.chunk_abb_base[f] = static_cast<int32_t>(P.chunk_abb.size() / CH);
    P.chunk_abb.resize(P.chunk_abb.size() + CH, -1);
    int32_t* row = &P.chunk_abb[P.chunk_abb.size() - CH];
    for(int o = 0; o < CH; ++o) row[sfc2lex[o]] = ids[o];
  }

  // ---- 7. link codes of the generic range
  // dynamic values (periodic with pressure) come after the static ones in the value table
  P.codes.assign(static_cast<size_t>(QM) * P.gen_stride, link_code(LK_VALUE, 0));
  // value 0 is a dummy for padding cells
  P.values.assign(1, 0.0);
  std::vector<std::pair<int64_t, int64_t>> stale_slots; // (ref cell, dir) -> static value index, filled at init
  std::vector<std::pair<size_t, int64_t>>  dyn_codes;   // code position -> dynamic value index
  for(int64_t c = 0; c < NO; ++c) {
    const int64_t dv = P.ref2dev[c];
    if(dv < P.gen_begin) continue;
    const int64_t g = dv - P.gen_begin;
    for(int j = 0; j < QM; ++j) {
      const size_t  at = static_cast<size_t>(j) * P.gen_stride + g;
      auto          it = over.find(key(c, j));
      const int32_t ps = pull[static_cast<size_t>(c) * QM + j];
      if(it == over.end()) {
        if(ps >= 0) P.codes[at] = P.ref2dev[ps];
        else {
          // nothing ever writes this slot: it keeps the value initialCondition() gave it
          P.codes[at] = link_code(LK_VALUE, static_cast<int32_t>(P.values.size()));
          stale_slots.emplace_back(c, j);
          P.values.push_back(0.0);
          ++P.slots_stale;
        }
        continue;
      }
      const SlotDesc& sd = it->second;
      ++P.slots_bc;
      switch(sd.kind) {
        case LK_COPY:
          P.codes[at] = link_code(LK_COPY, static_cast<int32_t>(P.copytab.size()));
          P.copytab.push_back(CopySrc{P.ref2dev[sd.a], sd.b});
          break;
        case LK_BB: P.codes[at] = link_code(LK_BB, 0); break;
        case LK_BB_ADD: {
          AddEntry e;
          e.n = sd.nadd;
          e.pad = 0;
          for(int d = 0; d < 3; ++d) e.v[d] = sd.add[d];
          P.codes[at] = link_code(LK_BB_ADD, static_cast<int32_t>(P.addtab.size()));
          P.addtab.push_back(e);
          break;
        }
        case LK_ABB: P.codes[at] = link_code(LK_ABB, static_cast<int32_t>(sd.a)); break;
        case -1: dyn_codes.emplace_back(at, sd.a); break;
        default: P.error = "internal: bad slot kind"; return false;
      }
    }
  }
  P.n_values_static = static_cast<int64_t>(P.values.size());
  for(auto& dc : dyn_codes) P.codes[dc.first] = link_code(LK_VALUE, static_cast<int32_t>(P.n_values_static + dc.second));
  P.values.resize(P.values.size() + P.perp.size() * Q, 0.0);
  if(P.values.size() >= (size_t(1) << 28) || P.copytab.size() >= (size_t(1) << 28) || P.addtab.size() >= (size_t(1) << 28)) {
    P.error = "too many boundary slots for 28-bit payloads";
    return false;
  }
  // stale slot bookkeeping: remember (cell, dir) so init can store the initial equilibrium there
  P.stale_ref.reserve(stale_slots.size());
  for(auto& s : stale_slots) P.stale_ref.push_back(static_cast<int64_t>(P.ref2dev[s.first]) * Q + s.second);

  // ---- 8. translate the remaining reference ids to device ids
  {
    // inward neighbours of a pressure cell that another rank owns: their velocity arrives with the halo; the entry then
    // refers to the slot of the receive buffer, encoded as -(slot + 1)
    int64_t nvs = 0, nvr = 0;
    for(int64_t v : in.vsend_count) nvs += v;
    for(int64_t v : in.vrecv_count) nvr += v;
    if(nvs != static_cast<int64_t>(in.vsend_cell.size()) || nvr != static_cast<int64_t>(in.vrecv_cell.size())
       || (!in.vsend_count.empty() && in.vsend_count.size() != in.peers.size()) || (!in.vrecv_count.empty() && in.vrecv_count.size() != in.peers.size())) {
      P.error = "velocity halo lists are inconsistent";
      return false;
    }
    std::unordered_map<int64_t, int32_t> vslot;
    for(size_t k = 0; k < in.vrecv_cell.size(); ++k) {
      const int64_t c = in.vrecv_cell[k];
      if(c < NO || c >= N) { P.error = "velocity halo receive entry is not a ghost cell"; return false; }
      vslot.emplace(c, static_cast<int32_t>(k)); // a ghost may be listed more than once: every copy carries the same value
    }
    P.n_vrecv = nvr;
    P.vsend_cells.resize(in.vsend_cell.size());
    for(size_t k = 0; k < in.vsend_cell.size(); ++k) {
      const int64_t c = in.vsend_cell[k];
      if(c < 0 || c >= NO) { P.error = "velocity halo send entry is not an owned cell"; return false; }
      P.vsend_cells[k] = P.ref2dev[c];
    }
    auto nb_dev = [&](int32_t ref, int32_t* out) -> bool {
      if(ref < NO) { *out = P.ref2dev[ref]; return true; }
      auto it = vslot.find(ref);
      if(it == vslot.end()) return false;
      *out = -(it->second + 1);
      return true;
    };
    for(AbbEntry& e : P.abb) {
      e.cell = P.ref2dev[e.cell];
      if(!nb_dev(e.n1, &e.n1) || !nb_dev(e.n2, &e.n2)) {
        P.error = "pressure boundary: an inward neighbour belongs to another rank and lbm_b200_set_vars_halo does not list it";
        return false;
      }
    }
  }
  for(ForceEntry& e : P.force) { e.target = P.ref2dev[e.target]; e.val = P.ref2dev[e.val]; }
  for(PerPEntry& e : P.perp) { e.cell = P.ref2dev[e.cell]; e.vbase += static_cast<int32_t>(P.n_values_static); }
  for(VarFix& v : P.varfix) v.cell = P.ref2dev[v.cell];
  for(int32_t& c : P.u0_cells) c = P.ref2dev[c];
  // halo lists -> flat device indices into the SoA population arrays
  {
    int64_t ns = 0, nr = 0;
    for(int64_t v : in.send_count) ns += v;
    for(int64_t v : in.recv_count) nr += v;
    if(ns != static_cast<int64_t>(in.send_cell.size()) || nr != static_cast<int64_t>(in.recv_cell.size())
       || in.send_cell.size() != in.send_dir.size() || in.recv_cell.size() != in.recv_dir.size()) {
      P.error = "halo lists are inconsistent";
      return false;
    }
    P.send_index.resize(in.send_cell.size());
    P.recv_index.resize(in.recv_cell.size());
    for(size_t k = 0; k < in.send_cell.size(); ++k) {
      const int64_t c = in.send_cell[k];
      if(c < 0 || c >= NO || in.send_dir[k] < 0 || in.send_dir[k] >= Q) { P.error = "halo send entry out of range"; return false; }
      P.send_index[k] = P.pop_index(in.send_dir[k], P.ref2dev[c]);
    }
    for(size_t k = 0; k < in.recv_cell.size(); ++k) {
      const int64_t c = in.recv_cell[k];
      if(c < NO || c >= N || in.recv_dir[k] < 0 || in.recv_dir[k] >= Q) { P.error = "halo receive entry is not a ghost cell"; return false; }
      P.recv_index[k] = P.pop_index(in.recv_dir[k], P.ref2dev[c]);
    }
  }
  {
    // m_vars fix-ups: keep only the last writer of every (cell, variable), in the reference's order
    std::unordered_map<int64_t, size_t> last;
    for(size_t i = 0; i < P.varfix.size(); ++i) last[static_cast<int64_t>(P.varfix[i].cell) * 8 + P.varfix[i].var] = i;
    std::vector<VarFix> kept;
    for(size_t i = 0; i < P.varfix.size(); ++i)
      if(last[static_cast<int64_t>(P.varfix[i].cell) * 8 + P.varfix[i].var] == i) kept.push_back(P.varfix[i]);
    P.varfix.swap(kept);
  }
  return true;
}

} // namespace lbm
