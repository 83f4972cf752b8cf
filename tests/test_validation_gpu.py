"""Physics validation of the 3D / TRT / MRT extensions, whose parity is not pinned by the reference:
analytic 3D Couette profile, exact projection of z-invariant D3Q19 flow onto D2Q9, TRT/MRT with equal rates == BGK."""
import numpy as np
import pytest

import lbm_b200
from lbm_b200.capi import box_topology

pytestmark = pytest.mark.gpu

NAMES = ["-x", "+x", "-y", "+y", "-z", "+z"]


def surfaces(nghbr, ndim):
    out = {}
    for d in range(2 * ndim):
        cells = np.nonzero(nghbr[:, d] < 0)[0].astype(np.int64)
        n = np.zeros(ndim)
        n[d // 2] = -1.0 if d % 2 == 0 else 1.0
        out[NAMES[d]] = (cells, np.tile(n, (len(cells), 1)))
    return out


def test_couette_3d_analytic_profile():
    """Plates at z = 0 and z = H (half-way bounce-back), upper plate moving with U in x, periodic in x and y:
    u_x(z_k) = U (k + 1/2) / Nz exactly at steady state."""
    shape, U = (8, 8, 16), 0.05
    nghbr, center, coords = box_topology(shape, (1, 1, 0), True, True)
    s = lbm_b200.Solver(3, 19, nghbr, 1.0 / 0.6, track_vars=0)
    srf = surfaces(nghbr, 3)
    s.add_dirichlet_bb(*srf["+z"], np.array([U, 0.0, 0.0]))
    s.add_wall_bb(*srf["-z"], 0.0)
    s.init()
    s.step(40000)
    m = s.moments()
    exact = U * (coords[:, 2] + 0.5) / shape[2]
    assert np.max(np.abs(m[:, 0] - exact)) < 1e-10
    assert np.max(np.abs(m[:, 1])) < 1e-12 and np.max(np.abs(m[:, 2])) < 1e-12  # rounding-level drift only
    assert np.max(np.abs(m[:, 3] - 1.0)) < 1e-10  # 40 000 steps of rounding in the wall addends


def channel(ndim, ndist, shape, periodic):
    nghbr, center, coords = box_topology(shape, periodic, True, True)
    s = lbm_b200.Solver(ndim, ndist, nghbr, 1.2, track_vars=0)
    srf = surfaces(nghbr, ndim)
    for nm in sorted(k for k in srf if len(srf[k][0])):  # lexicographic application order, like the reference
        if nm == "+x":
            s.add_pressure(*srf[nm], 1.0)
        elif nm == "-x":
            s.add_pressure(*srf[nm], 1.0000008)
        else:
            s.add_wall_bb(*srf[nm], 0.0)
    s.init()
    return s, coords


def test_z_invariant_d3q19_projects_onto_d2q9():
    """A D3Q19 flow that does not depend on z is, summed over c_z, exactly the D2Q9 flow of the same case
    (pressure in/outlet, bounce-back walls): checks 3D streaming, 3D bounce-back and the 3D anti-bounce-back."""
    s2, c2 = channel(2, 9, (32, 16), (0, 0))
    s3, c3 = channel(3, 19, (32, 16, 4), (0, 0, 1))
    s2.step(2000)
    s3.step(2000)
    m2, m3 = s2.moments(), s3.moments()
    lut = np.full((32, 16), -1)
    lut[c2[:, 0], c2[:, 1]] = np.arange(len(c2))
    j = lut[c3[:, 0], c3[:, 1]]
    assert np.max(np.abs(m3[:, 0] - m2[j, 0])) < 1e-13 and np.max(np.abs(m3[:, 1] - m2[j, 1])) < 1e-13
    assert np.max(np.abs(m3[:, 2])) < 1e-16
    assert np.max(np.abs(m3[:, 3] - m2[j, 2])) < 1e-13
    assert np.max(np.abs(m2[:, 0])) > 1e-8  # there is a flow to compare


@pytest.mark.parametrize("collision", [lbm_b200.TRT, lbm_b200.MRT])
def test_equal_rates_reduce_to_bgk(collision):
    nghbr, _, _ = box_topology((16, 16, 16), (1, 0, 0), False, False)
    srf = surfaces(nghbr, 3)

    def make(coll):
        s = lbm_b200.Solver(3, 19, nghbr, 1.3, collision=coll, omega_minus=1.3, mrt_rates=np.full(27, 1.3), track_vars=0)
        s.add_dirichlet_bb(*srf["+z"], np.array([0.05, 0.01, 0.0]))
        for nm in ("-z", "-y", "+y"):
            s.add_wall_bb(*srf[nm], 0.0)
        s.init()
        s.step(200)
        return s.f

    a, b = make(lbm_b200.BGK), make(collision)
    assert np.max(np.abs(a - b)) / np.max(np.abs(a)) < 1e-13
