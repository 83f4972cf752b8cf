"""Parity of the CUDA product (through the C ABI) against the CPU oracle and the reference's golden dumps.

fp64 STRICT arithmetic must be BIT-IDENTICAL to the oracle (which is itself bit-identical to the reference,
tests/test_oracle_golden.py).  fp64 FAST must agree within 1e-12 relative (BASELINE.json north_star).
"""
import numpy as np
import pytest

import lbm_b200
from casebuilder import CaseSpec, load_golden
from gridgen import box_grid

pytestmark = pytest.mark.gpu

GOLDEN_CASES = ["couette", "couette_bnd", "couette_bnd_bbDirichlet", "poiseuille", "poiseuille_bnd", "step_ns", "sphere_ns",
                # wet-node wall family (reference-order pipeline on the GPU, lbm_b200/csrc/sequential.cuh)
                "couette_bnd_eq", "couette_bnd_eq2", "couette_bnd_eq_aligned", "couette_bnd_NEEM", "couette_bnd_NEBB",
                "poiseuille_bnd_eq", "poiseuille_bnd_NEEM", "poiseuille_bnd_NEBB", "poiseuille_bnd_pressure",
                "poiseuille_bnd_pressure_neem2"]


def rel_err(a, b):
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0))


def run_pair(spec, oracle_mod, steps, **kw):
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    g = spec.apply_to(lbm_b200.Solver(spec.ndim, spec.ndist, spec.nghbr, spec.omega, **kw))
    o.init()
    g.init()
    return o, g


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_strict_fp64_is_bit_identical_on_reference_cases(name, oracle_mod):
    spec = load_golden(name)
    o, g = run_pair(spec, oracle_mod, 0)
    assert np.array_equal(g.f, o.f) and np.array_equal(g.fold, o.fold), "initial condition differs"
    done = 0
    for s in spec.golden["steps"]:
        s = int(s)
        o.step(s - done)
        g.step(s - done)
        done = s
        for arr in ("f", "fold", "vars", "varsold"):
            a, b = getattr(g, arr), getattr(o, arr)
            assert np.array_equal(a, b), f"{name} step {s}: {arr} differs, max abs {np.max(np.abs(a - b))}"
            key = f"{arr}_{s}"
            if key in spec.golden:
                assert np.array_equal(a, spec.golden[key]), f"{name} step {s}: {arr} differs from the reference dump"
        ro, _ = o.residual()
        rg, bad = g.residual()
        assert not bad
        assert np.allclose(rg, ro, rtol=1e-12, atol=1e-300)
    o.update_moments()
    assert np.array_equal(g.moments(), o.vars)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fast_fp64_within_1e12(name, oracle_mod):
    spec = load_golden(name)
    o, g = run_pair(spec, oracle_mod, 0, arithmetic=lbm_b200.FAST)
    n = int(spec.golden["steps"][-1])
    o.step(n)
    g.step(n)
    assert rel_err(g.f, o.f) < 1e-12
    assert rel_err(g.fold, o.fold) < 1e-12
    vo, vg = o.vars, g.vars
    assert rel_err(vg[:, -1], vo[:, -1]) < 1e-12          # density
    # velocities: these flows run at Ma 1e-4..1e-2, so u is a 1e-6..1e-3 difference of O(0.1) populations and cannot
    # be more accurate than the populations' own rounding; the tolerance is 1e-12 of the lattice speed of sound
    assert np.max(np.abs(vg[:, :-1] - vo[:, :-1])) < 1e-12 / np.sqrt(3.0)


def box_spec(shape, ndist, periodic, omega=1.0 / 0.6, lid=None):
    g = box_grid(shape, periodic)
    ndim = len(shape)
    spec = CaseSpec(name=f"box{shape}", ndim=ndim, ndist=ndist, nghbr=g["nghbr"], omega=omega, center=g["center"],
                    bbmin=g["bbmin"], bbmax=g["bbmax"], cell_length=g["cell_length"])
    names = ["-x", "+x", "-y", "+y", "-z", "+z"][:2 * ndim]
    for nm in sorted(names):  # lexicographic like the reference
        cells, normals = g["surfaces"][nm]
        if len(cells) == 0:
            continue
        if lid is not None and nm == lid[0]:
            spec.bcs.append(dict(kind="dirichlet_bb", cells=cells, normals=normals, value=np.array(lid[1], float)))
        else:
            spec.bcs.append(dict(kind="wall_bb", cells=cells, normals=normals, tangential=0.0))
    return spec


BOXES = [
    # 3D Couette-type benchmark shape: periodic x, walls y, moving lid +z
    ((32, 32, 32), 19, (True, False, False), ("+z", (0.05, 0.0, 0.0))),
    ((32, 32, 32), 27, (True, False, False), ("+z", (0.05, 0.0, 0.0))),
    ((32, 16, 24), 19, (True, True, False), ("+z", (0.05, 0.02, 0.0))),
    ((16, 16, 16), 19, (True, True, True), None),
    ((128, 128), 9, (True, False), ("+y", (0.05, 0.0))),
    ((96, 64), 9, (False, False), ("+y", (0.05, 0.0))),
]


@pytest.mark.parametrize("shape,ndist,periodic,lid", BOXES)
def test_strict_fp64_boxes_fast_chunks(shape, ndist, periodic, lid, oracle_mod):
    spec = box_spec(shape, ndist, periodic, lid=lid)
    o, g = run_pair(spec, oracle_mod, 0)
    # perturb the start so that all populations differ: a few steps of lid-driven flow first on the oracle
    o.step(3)
    g.set_populations(o.f, o.fold)
    for chunk in (1, 1, 5, 20):
        o.step(chunk)
        g.step(chunk)
        assert np.array_equal(g.f, o.f), f"f differs after {chunk}"
        assert np.array_equal(g.fold, o.fold)
    st = g.stats()
    if all(periodic):
        assert st["cells_fast"] == st["ncells"]
    side = 8 if len(shape) == 3 else 32
    interior = [s // side - (0 if p else 2) for s, p in zip(shape, periodic)]
    if all(s % side == 0 for s in shape) and min(interior) > 0:
        # interior chunks always qualify; chunks on a single wall qualify too (bounce-back selectors)
        assert st["cells_fast"] >= int(np.prod(interior)) * side ** len(shape), "template-indexed chunk path not taken"


@pytest.mark.parametrize("collision", [lbm_b200.TRT, lbm_b200.MRT])
@pytest.mark.parametrize("ndist,shape,periodic,lid", [(19, (32, 32, 32), (True, False, False), ("+z", (0.05, 0, 0))),
                                                      (9, (64, 64), (True, False), ("+y", (0.05, 0)))])
def test_trt_mrt_strict_match_oracle(collision, ndist, shape, periodic, lid, oracle_mod):
    spec = box_spec(shape, ndist, periodic, lid=lid)
    rates = np.linspace(1.1, 1.7, 27)
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    o.set_collision(collision, 1.3, rates)
    g = spec.apply_to(lbm_b200.Solver(spec.ndim, spec.ndist, spec.nghbr, spec.omega, collision=collision,
                                     omega_minus=1.3, mrt_rates=rates))
    o.init()
    g.init()
    o.step(30)
    g.step(30)
    assert np.array_equal(g.f, o.f)


def test_fp32_opt_in_tolerance(oracle_mod):
    """fp32 is an opt-in; stated tolerance: 2e-6 relative on populations after 100 steps of the couette case."""
    spec = load_golden("couette")
    o, g = run_pair(spec, oracle_mod, 0, precision=lbm_b200.FP32, arithmetic=lbm_b200.FAST)
    o.step(100)
    g.step(100)
    assert rel_err(g.f, o.f) < 2e-6
    assert np.max(np.abs(g.vars[:, 0] - o.vars[:, 0])) < 5e-5 * 0.1  # wall speed 0.1


def test_nan_in_the_state_is_reported_as_divergence(oracle_mod):
    """convergenceCondition (src/lbm/solver.cpp:254-260): a NaN / Inf in the residual sets m_diverged.  A NaN planted in m_fold spreads
    with the next steps; lbm_b200_residual must say diverged = 1 (and so does the oracle)."""
    spec = box_spec((24, 16, 16), 19, (True, False, False), lid=("+z", (0.05, 0, 0)))
    g = spec.apply_to(lbm_b200.Solver(spec.ndim, spec.ndist, spec.nghbr, spec.omega, track_vars=1))
    g.init()
    g.step(3)
    res, bad = g.residual()
    assert not bad and np.isfinite(res).all()
    fold = g.fold.copy()
    fold[fold.shape[0] // 2, 5] = np.nan
    g.set_populations(None, fold)
    g.step(2)
    res, bad = g.residual()
    assert bad, "a NaN in the state was not reported as divergence"


@pytest.mark.parametrize("case", ["box3d", "sphere3d"])
def test_fp32_opt_in_tolerance_3d(case, oracle_mod):
    """fp32 opt-in on 3D cases (a D3Q19 moving-lid box on the chunk path; the 3D sphere case, D3Q27, with cut cells on the link-code path):
    populations within 1e-5 relative (max |difference| / max |population|) of the fp64 oracle after 100 steps.  The 2D case holds 2e-6; in
    the closed 3D cases the float round-off of 19 / 27 populations per cell accumulates as a slow drift of the density (measured: box
    3.2e-6, sphere 5.6e-6)."""
    if case == "box3d":
        spec, steps = box_spec((32, 24, 24), 19, (True, False, False), lid=("+z", (0.05, 0, 0))), 100
    else:
        from cases3d import build_case
        spec, steps = build_case("sphere3d", 5), 100
    o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    o.init()
    o.step(steps)
    g = spec.apply_to(lbm_b200.Solver(spec.ndim, spec.ndist, spec.nghbr, spec.omega, precision=lbm_b200.FP32, arithmetic=lbm_b200.FAST))
    g.init()
    g.step(steps)
    assert rel_err(g.f, o.f) < 1e-5


@pytest.mark.parametrize("case", ["box3d", "couette"])
def test_device_side_output_encoding(case, oracle_mod):
    """lbm_b200_encode_output: the fields of the kept cells exactly as the reference's binary VTK writer stores them -- every value through
    the 15-decimal text round trip (string_helper.h:93-107, IO.h:479), base64(uint64 header = 8 * count || doubles), '=' padding -- with
    the filter gather, the rounding and the base64 done on the device.  Reference here: the same bytes built in Python from the moments."""
    import base64
    import struct
    spec = box_spec((24, 16, 16), 19, (True, False, False), lid=("+z", (0.05, 0, 0))) if case == "box3d" else load_golden("couette")
    g = spec.apply_to(lbm_b200.Solver(spec.ndim, spec.ndist, spec.nghbr, spec.omega))
    g.init()
    g.step(25)
    m = g.moments()
    rng = np.random.default_rng(3)
    for keep in (None, rng.random(m.shape[0]) < 0.37):
        got = g.encode_output(keep)
        sel = m if keep is None else m[keep]
        for v in range(m.shape[1]):
            col = np.array([float(f"{x:.15f}") for x in sel[:, v]])
            want = base64.b64encode(struct.pack("<Q", 8 * len(col)) + col.astype("<f8").tobytes())
            assert got[v] == want, f"field {v} differs"
    for _ in range(2):                                                      # an empty selection is the reference's encodeLE error, every time
        with pytest.raises(lbm_b200.LbmB200Error, match="length = 0"):
            g.encode_output(np.zeros(m.shape[0], dtype=bool))
    assert g.encode_output(keep) == got                                     # and leaves the encoder usable
    pinned = lbm_b200.HostBuffer(sum(len(t) for t in got))                  # page-locked destination (lbm_b200_host_alloc): same text
    again = g.encode_output(keep, out=pinned.array, raw=True)
    assert [bytes(a) for a in again] == got
    pinned.close()
