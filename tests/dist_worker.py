"""One rank of a domain-decomposed run; launched by tests/test_distributed.py under torch.distributed.run.

--mode oracle (CPU, gloo): every rank steps the CPU oracle on its partition (owned + ghost cells) and exchanges exactly the
    (cell, direction) populations of the halo lists through torch.distributed; checks that the owned cells equal the
    single-domain oracle run bit for bit.  This covers the host logic of the N>1 path: partition, ghosts, halo lists.
--mode gpu (NCCL): the same through the C ABI: lbm_b200 with ghosts + ncclSend/ncclRecv inside lbm_b200_step.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from gridgen import box_grid  # noqa: E402
from lbm_b200 import partition  # noqa: E402
from oracle import oracle  # noqa: E402

OMEGA = 1.0 / 0.6


def bcs_for(g, ndim, pressure=False):
    names = ["-x", "+x", "-y", "+y", "-z", "+z"][:2 * ndim]
    out = []
    for nm in sorted(names):
        cells, normals = g["surfaces"][nm]
        if len(cells) == 0:
            continue
        if pressure and nm in ("-x", "+x"):
            # anti-bounce-back pressure in- and outlet (sphere_ns.json / step_ns.json of the reference); the wall surfaces that
            # come first in the order own the edge cells' wall slots, the pressure surface rewrites their x slots afterwards
            out.append(("pressure", cells, normals, 1.0005 if nm == "-x" else 1.0))
        elif nm == names[-1]:
            v = np.zeros(ndim)
            v[0] = 0.05
            out.append(("dirichlet", cells, normals, v))
        else:
            out.append(("wall", cells, normals, 0.0))
    return out


def add_bcs(solver, bcs, lp=None):
    for kind, cells, normals, val in bcs:
        if lp is not None:
            cells, normals = lp.restrict(cells, normals)
            if len(cells) == 0:
                continue
        if kind == "dirichlet":
            solver.add_dirichlet_bb(cells, normals, val)
        elif kind == "pressure":
            solver.add_pressure(cells, normals, val)
        else:
            solver.add_wall_bb(cells, normals, val)


def exchange_host(lp, f):
    """halo exchange of an AoS array f[n_local, Q] through torch.distributed (CPU tensors)"""
    ops, recv_bufs = [], []
    so = ro = 0
    for k, q in enumerate(lp.peers):
        ns, nr = lp.send_count[k], lp.recv_count[k]
        if ns:
            sb = torch.from_numpy(np.ascontiguousarray(f[lp.send_cell[so:so + ns], lp.send_dir[so:so + ns]]))
            ops.append(dist.P2POp(dist.isend, sb, q))
        if nr:
            rb = torch.empty(nr, dtype=torch.float64)
            recv_bufs.append((rb, ro, nr))
            ops.append(dist.P2POp(dist.irecv, rb, q))
        so += ns
        ro += nr
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    for rb, ro, nr in recv_bufs:
        f[lp.recv_cell[ro:ro + nr], lp.recv_dir[ro:ro + nr]] = rb.numpy()


def exchange_velocity_host(lp, vars_, ndim):
    """velocity halo of the pressure boundary condition: m_vars velocity of the listed owned cells -> the peers' ghost cells"""
    ops, recv_bufs = [], []
    so = ro = 0
    for k, q in enumerate(lp.peers):
        ns, nr = lp.vsend_count[k], lp.vrecv_count[k]
        if ns:
            sb = torch.from_numpy(np.ascontiguousarray(vars_[lp.vsend_cell[so:so + ns], :ndim]))
            ops.append(dist.P2POp(dist.isend, sb, q))
        if nr:
            rb = torch.empty((nr, ndim), dtype=torch.float64)
            recv_bufs.append((rb, ro, nr))
            ops.append(dist.P2POp(dist.irecv, rb, q))
        so += ns
        ro += nr
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    for rb, ro, nr in recv_bufs:
        vars_[lp.vrecv_cell[ro:ro + nr], :ndim] = rb.numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="oracle")
    ap.add_argument("--shape", default="16,12,10")
    ap.add_argument("--ndist", type=int, default=19)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--case", default="box", help="box, or a 3D case of tests/cases3d.py: sphere3d, step3d (BASELINE.json configs[3], [4])")
    ap.add_argument("--level", type=int, default=5)
    ap.add_argument("--collision", default="bgk", choices=["bgk", "trt", "mrt"])
    ap.add_argument("--check-residual", action="store_true", dest="check_residual")
    ap.add_argument("--p2p", action="store_true", help="peer-to-peer halo (CUDA IPC mailboxes, copy engines) instead of ncclSend / ncclRecv")
    ap.add_argument("--bc", default="walls", help="walls: periodic x, walls, moving lid; pressure: pressure in-/outlet on -x/+x")
    args = ap.parse_args()
    shape = tuple(int(x) for x in args.shape.split(","))
    ndim = len(shape) if args.case == "box" else 3
    coll = {"bgk": 0, "trt": 1, "mrt": 2}[args.collision]
    with_pressure = args.bc == "pressure"
    periodic = (not with_pressure,) + (False,) * (ndim - 1)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.mode == "gpu":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")

    if args.case == "box":
        g = box_grid(shape, periodic)
        bcs = bcs_for(g, ndim, with_pressure)
    else:
        from cases3d import build_case, mrt_rates
        spec = build_case(args.case, args.level)
        args.ndist = spec.ndist
        g = {"nghbr": spec.nghbr}
        kinds = {"wall_bb": "wall", "pressure": "pressure", "dirichlet_bb": "dirichlet"}
        bcs = [(kinds[bc["kind"]], bc["cells"], bc["normals"], bc.get("pressure", bc.get("value", bc.get("tangential")))) for bc in spec.bcs]
    stride = g["nghbr"].shape[1]
    pressure = [(cells, normals) for kind, cells, normals, _ in bcs if kind == "pressure"]
    rates = OMEGA * (1.0 + 0.01 * np.arange(27) / 27)
    om_minus = 1.0 / 0.8

    def make_oracle(table):
        orc = oracle.Oracle(ndim, args.ndist, table, OMEGA)
        orc.set_collision(coll, om_minus, rates)
        return orc
    # single-domain reference run (CPU oracle)
    ref = make_oracle(g["nghbr"])
    add_bcs(ref, bcs)
    ref.init()
    ref.step(args.steps)

    # partition: the table-driven provider and the on-demand box provider must give the same plan
    lp = partition.plan_rank(partition.TableRows(g["nghbr"], args.ndist), rank, world, stride, pressure)
    if args.case == "box":
        lp2 = partition.plan_rank(partition.BoxRows(shape, [int(p) for p in periodic], args.ndist), rank, world, stride, pressure)
        assert np.array_equal(lp.nghbr, lp2.nghbr) and np.array_equal(lp.ghosts, lp2.ghosts)
        assert lp.peers == lp2.peers and np.array_equal(lp.send_cell, lp2.send_cell) and np.array_equal(lp.recv_dir, lp2.recv_dir)
        assert np.array_equal(lp.vsend_cell, lp2.vsend_cell) and np.array_equal(lp.vrecv_cell, lp2.vrecv_cell)
    # send counts of mine must equal the receive counts of the peer
    counts = [None] * world
    dist.all_gather_object(counts, {q: (lp.send_count[k], lp.recv_count[k], lp.vsend_count[k], lp.vrecv_count[k]) for k, q in enumerate(lp.peers)})
    for k, q in enumerate(lp.peers):
        assert counts[q][rank] == (lp.recv_count[k], lp.send_count[k], lp.vrecv_count[k], lp.vsend_count[k]), "halo lists of the two sides do not match"
    nvel = [None] * world
    dist.all_gather_object(nvel, sum(lp.vrecv_count))
    if with_pressure and world > 1 and args.case == "box":
        assert sum(nvel) > 0, "test set-up: no pressure cell is separated from its inward neighbours by the cut"

    if args.mode == "oracle":
        o = make_oracle(lp.nghbr)
        add_bcs(o, bcs, lp)
        o.init()
        for _ in range(args.steps):
            o.step_collide()
            exchange_host(lp, o.f)
            exchange_velocity_host(lp, o.vars, ndim)
            o.step_stream()
        mine_f, mine_fold = o.f[:lp.n_owned].copy(), o.fold[:lp.n_owned].copy()
    else:
        import lbm_b200
        from lbm_b200.capi import comm_unique_id
        uid = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        s = lbm_b200.Solver(ndim, args.ndist, lp.nghbr, OMEGA, device=local, collision=coll, omega_minus=om_minus, mrt_rates=rates)
        add_bcs(s, bcs, lp)
        lp.apply_halo(s)
        s.comm_init(uid[0], rank, world)
        s.init()
        if args.p2p:
            s.p2p_connect(dist)
        s.step(args.steps)
        s.synchronize()
        mine_f, mine_fold = s.f[:lp.n_owned], s.fold[:lp.n_owned]
        # residual of the whole domain: ncclAllReduce over the ranks' owned cells (order of the sum differs from the serial one)
        if args.check_residual:
            res, bad = s.residual()
            ref_res, _ = ref.residual()
            assert not bad and np.allclose(res, ref_res, rtol=1e-11, atol=1e-300), (res, ref_res)
        st = s.stats()
        assert st["cells_ghost"] == lp.n_ghost and (world == 1 or st["halo_bytes"] > 0)
        torch.cuda.synchronize()
        dist.barrier()
        s.close()
    ok = np.array_equal(mine_f, ref.f[lp.lo:lp.hi]) and np.array_equal(mine_fold, ref.fold[lp.lo:lp.hi])
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if rank == 0:
        print("PARTITION_PARITY", "OK" if all(flags) else f"FAILED {flags}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
