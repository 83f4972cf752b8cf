"""numpy interpreter of the device plan (test helper): executes the GATHER of the fused kernel on the CPU from the arrays
lbm_b200_debug_plan exports, so that the plan (lbm_b200/csrc/plan.hpp: inverted push table, boundary-condition resolution,
chunk templates, wall descriptors, ghost blocks, halo lists) can be checked against the oracle without a GPU.  It never runs
in the product."""
import numpy as np

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
    dev[:, plan["ref2dev"]] = aos.T
    return dev


def gather(plan, A, q, values=None, uext=None):
    """m_fold of every owned cell, in REFERENCE order, from device populations A [Q, npad]"""
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
            val[pull] = A[j, nbv[pull] + off[pull]]
            if (~pull).any():
                v = A[opp[j], cells[~pull]].copy()
                e = plan["wall_desc"][wid, j]
                for a in range(int(e[3])):
                    v = v + e[a]
                val[~pull] = v
            fold_dev[j, cells] = val
        fold_dev[qm, cells] = A[qm, cells]
    # generic range: link codes
    g0 = plan["gen_begin"]
    for g in range(plan["n_gen"]):
        cell = g0 + g
        for j in range(qm):
            code = int(plan["codes"][j, g])
            if code >= 0:
                fold_dev[j, cell] = A[j, code]
                continue
            kind, pl = (code >> 28) & 7, code & 0x0FFFFFFF
            if kind == 0:
                sc, sd = plan["copytab"][pl]
                fold_dev[j, cell] = A[sd, sc]
            elif kind == 1:
                fold_dev[j, cell] = A[opp[j], cell]
            elif kind == 2:
                v = A[opp[j], cell]
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
                fold_dev[j, cell] = -A[opp[j], cell] + 2 * se
            else:
                fold_dev[j, cell] = values[pl]
        fold_dev[qm, cell] = A[qm, cell]
    return fold_dev[:, plan["ref2dev"][:plan["n_owned"]]].T
