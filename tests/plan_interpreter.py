"""numpy interpreter of the device plan (test helper): executes the GATHER of the fused kernel on the CPU from the arrays
lbm_b200_debug_plan exports, so that the plan (lbm_b200/csrc/plan.hpp: inverted push table, boundary-condition resolution,
chunk templates, wall descriptors, ghost blocks, halo lists) can be checked against the oracle without a GPU.  It never runs
in the product."""
import numpy as np

from lbm_b200.capi import pop_scatter, pop_slot

OPP = {9: [1, 0, 3, 2, 6, 7, 4, 5, 8],
       19: [1, 0, 3, 2, 5, 4, 9, 8, 7, 6, 13, 12, 11, 10, 17, 16, 15, 14, 18],
       27: [1, 0, 3, 2, 5, 4, 9, 8, 7, 6, 13, 12, 11, 10, 17, 16, 15, 14, 25, 24, 23, 22, 21, 20, 19, 18, 26]}
D3C = np.array([[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1], [-1, -1, 0], [-1, 1, 0], [1, -1, 0], [1, 1, 0],
                [-1, 0, -1], [-1, 0, 1], [1, 0, -1], [1, 0, 1], [0, -1, -1], [0, -1, 1], [0, 1, -1], [0, 1, 1], [-1, -1, -1],
                [-1, -1, 1], [-1, 1, -1], [-1, 1, 1], [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1], [0, 0, 0]], float)
D2C = np.array([[-1, 0], [1, 0], [0, -1], [0, 1], [1, 1], [1, -1], [-1, -1], [-1, 1], [0, 0]], float)


def weights(q):
    if q == 9:
        return np.array([1 / 9] * 4 + [1 / 36] * 4 + [4 / 9])
    if q == 19:
        return np.array([1 / 18] * 6 + [1 / 36] * 12 + [1 / 3])
    return np.array([2 / 27] * 6 + [1 / 54] * 12 + [1 / 216] * 8 + [8 / 27])


def to_device(plan, aos, q):
    """reference-order AoS [n, Q] -> device SoA [Q, npad]"""
    dev = np.zeros((q, plan["npad"]))
    pop_scatter(plan, dev, plan["ref2dev"].astype(np.int64), aos)   # per-direction in-chunk layouts (include/lbm_b200.h)
    return dev


def gather(plan, A, q, values=None, uext=None):
    """m_fold of every owned cell, in REFERENCE order, from device populations A [Q, npad] (population layout: direction j of
    device cell c sits at pop_slot(plan, j, c); the result fold_dev is a plain per-cell array)"""
    def P(j, cells):
        return pop_slot(plan, j, cells)
    qm, ch, nsel = q - 1, plan["chunk"], plan["nsel"]
    opp = OPP[q]
    w = weights(q)
    c = (D2C if q == 9 else D3C[:q - 1].tolist() + [[0, 0, 0]])
    c = np.array(c, float)
    values = plan["values"] if values is None else values
    fold_dev = np.zeros((q, plan["npad"]))
    # fast chunks: template + neighbour-chunk bases, wall selectors bounce back with the chunk's wall descriptor
    for k in range(plan["n_fast_chunks"]):
        nb = plan["chunk_nb"][k]
        base = k * ch
        cells = base + np.arange(ch)
        wid = nb[nsel]
        for j in range(qm):
            t = plan["tmpl"][j].astype(np.int64)
            sel, off = t >> 10, t & 1023
            nbv = nb[sel]
            val = np.empty(ch)
            pull = nbv >= 0
            val[pull] = A[j, nbv[pull] + off[pull]]                # template offsets are layout positions already
            for s_miss in np.unique(sel[~pull]) if (~pull).any() else []:
                m = (~pull) & (sel == s_miss)                     # the slots whose source lies in this missing neighbour chunk
                v = A[opp[j], P(opp[j], cells[m])].copy()
                e = plan["wall_desc"][wid, s_miss, j]
                if e[3] < 0:
                    # chunk on a pressure face: anti-bounce-back with the cell's own pressure entry (bnd_pressure.h:100)
                    ent = plan["chunk_abb"][int(plan["chunk_abb_base"][k])][m].astype(np.int64)
                    assert (ent >= 0).all()
                    u = uext[ent]
                    cu = np.zeros(len(ent))
                    for d in range(u.shape[1]):  # Phys::cu_rt / vsq: sums in ascending dimension order
                        cu = cu + u[:, d] * c[opp[j]][d]
                    vs = u[:, 0] * u[:, 0]
                    for d in range(1, u.shape[1]):
                        vs = vs + u[:, d] * u[:, d]
                    cs = 1.0 / 3.0
                    se = w[opp[j]] * plan["abb_p"][ent] * (1.0 + cu * cu / (2.0 * cs * cs) - vs / (2.0 * cs))
                    v = -v + 2 * se
                for a in range(max(int(e[3]), 0)):
                    v = v + e[a]
                val[m] = v
            fold_dev[j, cells] = val
        fold_dev[qm, cells] = A[qm, cells]
        assert plan["layout"][qm] == 0
    # generic range: link codes
    g0 = plan["gen_begin"]
    for g in range(plan["n_gen"]):
        cell = g0 + g
        for j in range(qm):
            code = int(plan["codes"][j, g])
            if code >= 0:
                fold_dev[j, cell] = A[j, P(j, code)]
                continue
            kind, pl = (code >> 28) & 7, code & 0x0FFFFFFF
            if kind == 0:
                sc, sd = plan["copytab"][pl]
                fold_dev[j, cell] = A[sd, P(sd, sc)]
            elif kind == 1:
                fold_dev[j, cell] = A[opp[j], P(opp[j], cell)]
            elif kind == 2:
                v = A[opp[j], P(opp[j], cell)]
                e = plan["addtab"][pl]
                for a in range(int(e[3])):
                    v = v + e[a]
                fold_dev[j, cell] = v
            elif kind == 3:
                u = uext[pl]
                p = plan["abb_p"][pl]
                cu = float(np.dot(u, c[opp[j]][:len(u)]))
                vs = float(np.dot(u, u))
                cs = 1.0 / 3.0
                se = w[opp[j]] * p * (1.0 + cu * cu / (2.0 * cs * cs) - vs / (2.0 * cs))
                fold_dev[j, cell] = -A[opp[j], P(opp[j], cell)] + 2 * se
            else:
                fold_dev[j, cell] = values[pl]
        fold_dev[qm, cell] = A[qm, cell]
    return fold_dev[:, plan["ref2dev"][:plan["n_owned"]]].T


def stale_values(plan, init_fold_rows, q):
    """value table with the slots nothing ever writes filled from the initial m_fold (what lbm_b200_init does on the device);
    init_fold_rows[k] = initial m_fold row of the plan's reference cell k"""
    values = plan["values"].copy()
    dev2ref = np.full(plan["npad"], -1)
    dev2ref[plan["ref2dev"]] = np.arange(plan["n"])
    for k, sr in enumerate(plan["stale_ref"].astype(np.int64)):
        values[k + 1] = init_fold_rows[dev2ref[sr // q], sr % q]
    return values


def extrapolated_velocity(plan, vel_of_dev, vrecv, ndim):
    """u_ext of every anti-bounce-back entry: 1.5 u(n1) - 0.5 u(n2) (bnd_pressure.h:78-84); n < 0 = slot of the received halo"""
    ab = plan["abb_cells"].astype(np.int64)
    if len(ab) == 0:
        return None

    def vel(n):
        out = np.empty((len(n), ndim))
        loc = n >= 0
        out[loc] = vel_of_dev(n[loc])
        if (~loc).any():
            out[~loc] = vrecv[-n[~loc] - 1]
        return out
    return 1.5 * vel(ab[:, 1]) - 0.5 * vel(ab[:, 2])


def partitioned_fold(r, plans, lps, f_glob, vars_glob, init_fold_glob, q, ndim):
    """m_fold of the cells rank r owns, computed from ITS plan: owned populations from the global post-collision state, ghost
    populations and remote velocities only through what the peers' send lists put on the wire (emulates k_halo_pack /
    k_velocity_pack -> NCCL -> k_halo_unpack without a GPU)."""
    plan, lp = plans[r], lps[r]
    own = np.arange(lp.lo, lp.hi)
    glob_of_local = np.concatenate([own, lp.ghosts])
    dev2glob = np.full(plan["npad"], -1)
    dev2glob[plan["ref2dev"][:lp.n_owned]] = own
    A = np.zeros((q, plan["npad"]))
    pop_scatter(plan, A, plan["ref2dev"][:lp.n_owned].astype(np.int64), f_glob[own])
    vrecv = np.zeros((plan["n_vrecv"], ndim))
    ro = vo = 0
    for k, peer in enumerate(lp.peers):
        pq, lq = plans[peer], lps[peer]
        kq = lq.peers.index(r)
        ownq = np.arange(lq.lo, lq.hi)
        q2glob = np.full(pq["npad"], -1)
        q2glob[pq["ref2dev"][:lq.n_owned]] = ownq
        Aq = np.zeros((q, pq["npad"]))
        pop_scatter(pq, Aq, pq["ref2dev"][:lq.n_owned].astype(np.int64), f_glob[ownq])
        so, ns = sum(lq.send_count[:kq]), lq.send_count[kq]
        assert ns == lp.recv_count[k]
        A.reshape(-1)[plan["recv_index"][ro:ro + ns].astype(np.int64)] = Aq.reshape(-1)[pq["send_index"][so:so + ns].astype(np.int64)]
        ro += ns
        vso, vns = sum(lq.vsend_count[:kq]), lq.vsend_count[kq]
        assert vns == lp.vrecv_count[k]
        vrecv[vo:vo + vns] = vars_glob[q2glob[pq["vsend_cells"][vso:vso + vns].astype(np.int64)], :ndim]
        vo += vns
    uext = extrapolated_velocity(plan, lambda n: vars_glob[dev2glob[n], :ndim], vrecv, ndim)
    values = stale_values(plan, init_fold_glob[glob_of_local], q)
    return gather(plan, A, q, values=values, uext=uext)
