"""Partition of the SFC-ordered cell list into contiguous ranges, one per GPU (host logic, numpy).

The reference has no domain decomposition (`partitionLevel`, halo/window cell properties and the
load-balancing weights are declared but every implementation is a stub, src/cartesiangrid.h:709-710,
src/loadbalancing_weights.h:6-31).  This is the B200-native design of SURVEY.md section 8e: equal-count contiguous ranges
of the curve (uniform weights, the reference's only WeightMethod), ghost copies of the remote cells an owned cell pushes
to or pulls from, and per-peer lists of exactly the (cell, direction) populations that cross the cut.

Both sides of every exchange derive their list from the tables alone, in (global cell id, direction) order, so no
set-up communication is needed and the partitioned run is bit-identical to the single-GPU run.
"""
from dataclasses import dataclass, field

import numpy as np

OPP = {9: [1, 0, 3, 2, 6, 7, 4, 5, 8],
       19: [1, 0, 3, 2, 5, 4, 9, 8, 7, 6, 13, 12, 11, 10, 17, 16, 15, 14, 18],
       27: [1, 0, 3, 2, 5, 4, 9, 8, 7, 6, 13, 12, 11, 10, 17, 16, 15, 14, 25, 24, 23, 22, 21, 20, 19, 18, 26]}


def bounds(n, world):
    """Equal-count contiguous ranges: rank r owns [b[r], b[r+1])."""
    return np.array([r * n // world for r in range(world + 1)], dtype=np.int64)


class TableRows:
    """Row provider over a full neighbour table held in memory (reference-sized cases)."""

    def __init__(self, nghbr, ndist):
        self.nghbr = np.ascontiguousarray(nghbr, dtype=np.int64)
        self.n = self.nghbr.shape[0]
        self.qm = ndist - 1
        # inverse of the push table; the table need not be symmetric, so it is inverted explicitly
        self.pull = np.full((self.n, self.qm), -1, dtype=np.int64)
        src = np.arange(self.n, dtype=np.int64)
        for j in range(self.qm):
            t = self.nghbr[:, j]
            ok = t >= 0
            self.pull[t[ok], j] = src[ok]

    def rows(self, ids):
        return self.nghbr[ids][:, :self.qm]

    def sources(self, ids):
        return self.pull[ids]


class BoxRows:
    """Row provider for the synthetic benchmark box: rows are generated on demand (lbm_b200_box_rows), the full table of a
    multi-GPU box never exists on any rank.  A box table is symmetric, so pull sources are the opposite-direction rows."""

    def __init__(self, shape, periodic, ndist):
        self.shape, self.periodic = tuple(shape), tuple(periodic)
        self.n = int(np.prod(self.shape))
        self.qm = ndist - 1
        self.opp = OPP[ndist][:self.qm]

    def rows(self, ids):
        from .capi import box_rows
        if getattr(self, "_last_ids", None) is ids:  # plan_rank asks for rows and sources of the same range
            return self._last_rows
        self._last_ids, self._last_rows = ids, box_rows(self.shape, self.periodic, ids)[0][:, :self.qm]
        return self._last_rows

    def sources(self, ids):
        return self.rows(ids)[:, self.opp]

    def centers(self, ids):
        from .capi import box_rows
        return box_rows(self.shape, self.periodic, ids, want_center=True)[1]


class GridRows:
    """Row provider over a configuration's single-level grid, rows generated on demand (lbm_b200.host_api.UniformGrid): box / sphere /
    step geometries of any size without a table over the whole domain.  Composed diagonal steps make the table asymmetric next to
    cut cells, so the pull sources come from the grid object (it evaluates the source's own composition)."""

    def __init__(self, grid, ndist):
        self.grid = grid
        self.n = int(grid.n)
        self.qm = ndist - 1

    def rows(self, ids):
        if getattr(self, "_last_ids", None) is ids:
            return self._last_rows
        self._last_ids, self._last_rows = ids, np.ascontiguousarray(self.grid.rows(ids)[0][:, :self.qm])
        return self._last_rows

    def sources(self, ids):
        return np.ascontiguousarray(self.grid.sources(ids)[:, :self.qm])

    def centers(self, ids):
        return self.grid.rows(ids, want_center=True)[1]


@dataclass
class LocalProblem:
    rank: int
    world: int
    lo: int
    hi: int
    ghosts: np.ndarray                 # global ids of the ghost cells, ascending
    nghbr: np.ndarray                  # [n_owned + n_ghost, stride] local ids
    peers: list = field(default_factory=list)
    send_count: list = field(default_factory=list)
    recv_count: list = field(default_factory=list)
    send_cell: np.ndarray = None       # local ids, concatenated in peer order
    send_dir: np.ndarray = None
    recv_cell: np.ndarray = None
    recv_dir: np.ndarray = None
    # velocity halo of the pressure boundary condition (bnd_pressure.h:78-84 reads m_vars of the two inward neighbours):
    # per peer, the owned cells whose velocity the peer's pressure cells extrapolate from / the ghost cells that receive one
    vsend_count: list = field(default_factory=list)
    vrecv_count: list = field(default_factory=list)
    vsend_cell: np.ndarray = None      # local ids of owned cells
    vrecv_cell: np.ndarray = None      # local ids of ghost cells

    @property
    def n_owned(self):
        return self.hi - self.lo

    @property
    def n_ghost(self):
        return len(self.ghosts)

    def to_local(self, gids):
        gids = np.asarray(gids, dtype=np.int64)
        out = np.full(gids.shape, -1, dtype=np.int64)
        own = (gids >= self.lo) & (gids < self.hi)
        out[own] = gids[own] - self.lo
        rest = ~own & (gids >= 0)
        if rest.any() and len(self.ghosts):
            pos = np.searchsorted(self.ghosts, gids[rest])
            pos = np.minimum(pos, len(self.ghosts) - 1)
            hit = self.ghosts[pos] == gids[rest]
            out[rest] = np.where(hit, self.n_owned + pos, -1)
        return out

    def restrict(self, cells, normals):
        """Boundary-condition entries of the cells this rank owns, order kept (application order is per cell)."""
        cells = np.asarray(cells, dtype=np.int64)
        m = (cells >= self.lo) & (cells < self.hi)
        return cells[m] - self.lo, np.asarray(normals)[m]

    def apply_halo(self, solver):
        solver.set_ghosts(self.n_ghost)
        solver.set_halo(self.peers, self.send_count, self.send_cell, self.send_dir, self.recv_count, self.recv_cell, self.recv_dir)
        if sum(self.vsend_count) + sum(self.vrecv_count) > 0:
            solver.set_vars_halo(self.vsend_count, self.vsend_cell, self.vrecv_count, self.vrecv_cell)


def inward_direction(normals):
    """LBMBnd_Pressure::apply (bnd_pressure.h:58-66): the inside direction is the axis direction against the first non-zero
    component of the normal."""
    normals = np.asarray(normals, dtype=np.float64)
    ins = np.full(len(normals), -1, dtype=np.int64)
    for d in range(normals.shape[1] - 1, -1, -1):
        ins = np.where(normals[:, d] < 0, 2 * d + 1, np.where(normals[:, d] > 0, 2 * d, ins))
    return ins


def pressure_stencil(provider, pressure):
    """For every entry of every pressure surface (application order): the cell and its two inward neighbours, global ids."""
    cs, n1s, n2s = [], [], []
    for cells, normals in pressure:
        cells = np.asarray(cells, dtype=np.int64)
        if len(cells) == 0:
            continue
        ins = inward_direction(normals)
        if (ins < 0).any():
            raise ValueError("pressure boundary: zero normal")
        n1 = np.ascontiguousarray(provider.rows(cells)[np.arange(len(cells)), ins])
        if (n1 < 0).any():
            raise ValueError("pressure boundary: cell without two inward neighbours")
        n2 = np.ascontiguousarray(provider.rows(n1)[np.arange(len(cells)), ins])
        if (n2 < 0).any():
            raise ValueError("pressure boundary: cell without two inward neighbours")
        cs.append(cells)
        n1s.append(n1)
        n2s.append(n2)
    if not cs:
        z = np.zeros(0, dtype=np.int64)
        return z, z, z
    c, n1, n2 = np.concatenate(cs), np.concatenate(n1s), np.concatenate(n2s)
    # the reference applies the entries one after the other and reads m_vars of n1 / n2: a neighbour that an earlier entry
    # has already rewritten makes the result order-dependent (plan.hpp rejects the same thing inside one rank)
    first = {}
    for k, cell in enumerate(c.tolist()):
        first.setdefault(cell, k)
    for k in range(len(c)):
        for nb in (int(n1[k]), int(n2[k])):
            if first.get(nb, len(c)) < k:
                raise ValueError("pressure boundary: inward neighbour is itself a pressure boundary cell (order-dependent in the reference)")
    return c, n1, n2


def plan_rank(provider, rank, world, stride, pressure=None):
    """pressure: [(cells, normals), ...] = the GLOBAL cell lists of all pressure boundary conditions in application order (every
    rank passes the same lists); needed only when such a surface exists, so that the velocity of inward neighbours that
    another rank owns travels with the halo."""
    n, qm = provider.n, provider.qm
    b = bounds(n, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    own = np.arange(lo, hi, dtype=np.int64)
    rows_own = provider.rows(own)
    src_own = provider.sources(own)
    owner_of = lambda g: np.searchsorted(b, g, side="right") - 1

    def outside(a):
        a = a[a >= 0]
        return a[(a < lo) | (a >= hi)]

    # velocity halo: items (entry, which of n1 / n2) whose pressure cell and inward neighbour have different owners, in entry order
    pc, pn1, pn2 = pressure_stencil(provider, pressure or [])
    item_cell = np.stack([pc, pc], axis=1).ravel()
    item_nb = np.stack([pn1, pn2], axis=1).ravel()
    item_cown, item_nown = owner_of(item_cell), owner_of(item_nb)
    v_recv = (item_cown == rank) & (item_nown != rank)   # I own the pressure cell, a peer owns the neighbour
    v_send = (item_nown == rank) & (item_cown != rank)   # I own the neighbour, a peer owns the pressure cell

    ghosts = np.unique(np.concatenate([outside(rows_own.ravel()), outside(src_own.ravel()), item_nb[v_recv]]))
    lp = LocalProblem(rank=rank, world=world, lo=lo, hi=hi, ghosts=ghosts, nghbr=None)
    n_local = (hi - lo) + len(ghosts)
    table = np.full((n_local, stride), -1, dtype=np.int64)
    table[:hi - lo, :qm] = lp.to_local(rows_own)
    if len(ghosts):
        rows_g = provider.rows(ghosts)
        into_me = (rows_g >= lo) & (rows_g < hi)
        table[hi - lo:, :qm] = np.where(into_me, rows_g - lo, -1)
    lp.nghbr = table
    # send: my cell s pushes (direction j) into a cell another rank owns
    s_idx, s_dir = np.nonzero((rows_own >= 0) & ((rows_own < lo) | (rows_own >= hi)))
    s_owner = owner_of(rows_own[s_idx, s_dir])
    # receive: a ghost pushes (direction j) into a cell I own
    if len(ghosts):
        g_idx, g_dir = np.nonzero(table[hi - lo:, :qm] >= 0)
        g_owner = owner_of(ghosts[g_idx])
    else:
        g_idx = g_dir = g_owner = np.zeros(0, dtype=np.int64)
    # the pressure boundary condition finds n1 = N(c, inside), n2 = N(n1, inside) in the table: give it the links of the
    # ghost cells it walks over (ghost rows otherwise only keep links into owned cells)
    if len(pc):
        mine = item_cown[0::2] == rank
        ins_dir = np.zeros(len(pc), dtype=np.int64)
        k0 = 0
        for cells, normals in pressure:
            ins_dir[k0:k0 + len(cells)] = inward_direction(normals)
            k0 += len(cells)
        l1, l2 = lp.to_local(pn1[mine]), lp.to_local(pn2[mine])
        assert (l1 >= 0).all() and (l2 >= 0).all()
        table[l1, ins_dir[mine]] = l2
    peers = sorted(set(s_owner.tolist()) | set(g_owner.tolist()) | set(item_cown[v_send].tolist()) | set(item_nown[v_recv].tolist()))
    vs, vr = [], []
    for q in peers:
        m = v_send & (item_cown == q)
        vs.append(item_nb[m] - lo)
        lp.vsend_count.append(int(m.sum()))
        m = v_recv & (item_nown == q)
        vr.append(lp.to_local(item_nb[m]))
        lp.vrecv_count.append(int(m.sum()))
    sc, sd, rc, rd = [], [], [], []
    for q in peers:
        m = s_owner == q
        sc.append(s_idx[m])
        sd.append(s_dir[m])
        lp.send_count.append(int(m.sum()))
        m = g_owner == q
        rc.append(g_idx[m] + (hi - lo))
        rd.append(g_dir[m])
        lp.recv_count.append(int(m.sum()))
    cat = lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dt)
    lp.peers = [int(q) for q in peers]
    lp.send_cell, lp.send_dir = cat(sc, np.int64), cat(sd, np.int32)
    lp.recv_cell, lp.recv_dir = cat(rc, np.int64), cat(rd, np.int32)
    lp.vsend_cell, lp.vrecv_cell = cat(vs, np.int64), cat(vr, np.int64)
    return lp
