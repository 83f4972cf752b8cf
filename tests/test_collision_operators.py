"""TRT and MRT are extensions (the reference is BGK only), so nothing pins them but their own defining properties.
Checked here on the CPU oracle (the GPU kernels are bit-identical to it: tests/test_gpu_parity.py, tests/test_zz_baseline_configs_gpu.py):

  * the generated moment bases are orthogonal, conserved rows first, non-conserved rows orthogonal to 1 and c (tools/gen_mrt_tables.py);
  * MRT conserves mass and momentum to rounding with EVERY non-conserved rate different; TRT does too;
  * MRT / TRT with all rates equal are BGK (1e-13);
  * the shear rate alone sets the viscosity: a decaying shear wave decays with nu = cs^2 (1/s_shear - 1/2) whatever the bulk and
    ghost rates are, and differently when the shear rate changes;
  * TRT with the magic parameter 3/16: on the reference's own Poiseuille case the bounce-back wall sits closer to half-way than with BGK
    at the same viscosity (exactly half-way would need a body force, which the reference does not have).
"""
import re

import numpy as np
import pytest

from gridgen import D2_DIRS, D3_DIRS, box_grid

ROOT = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
LATT = {9: (2, "D2Q9"), 19: (3, "D3Q19"), 27: (3, "D3Q27")}


def dirs(q):
    if q == 9:
        return np.vstack([D2_DIRS, [[0, 0]]]).astype(float)
    return np.vstack([D3_DIRS[:q - 1], [[0, 0, 0]]]).astype(float)


def basis(q):
    text = open(f"{ROOT}/oracle/mrt_tables.h").read()
    name = LATT[q][1]
    body = text[text.index(f"#define LBM_MRT_{name}_M"):text.index(f"#define LBM_MRT_{name}_NORM")]
    rows = re.findall(r"\{([-0-9, ]+)\}", body)
    M = np.array([[int(x) for x in r.split(",")] for r in rows])
    norm = np.array([int(x) for x in re.search(r"LBM_MRT_%s_NORM \{([^}]*)\}" % name, text).group(1).split(",")])
    kind = np.array([int(x) for x in re.search(r"LBM_MRT_%s_KIND \{([^}]*)\}" % name, text).group(1).split(",")])
    return M, norm, kind


def test_product_and_oracle_tables_are_the_same_generated_text():
    a = open(f"{ROOT}/oracle/mrt_tables.h").read().split("\n", 1)[1].replace("LBM_ORACLE_MRT_TABLES_H", "X")
    b = open(f"{ROOT}/lbm_b200/csrc/mrt_tables.h").read().split("\n", 1)[1].replace("LBM_B200_MRT_TABLES_H", "X")
    assert a == b


@pytest.mark.parametrize("q", [9, 19, 27])
def test_moment_basis(q):
    M, norm, kind = basis(q)
    d = LATT[q][0]
    c = dirs(q)
    assert M.shape == (q, q) and np.linalg.matrix_rank(M) == q
    G = M @ M.T
    assert np.array_equal(G, np.diag(norm)), "rows are not orthogonal"
    assert np.array_equal(M[0], np.ones(q)) and all(np.array_equal(M[1 + a], c[:, a]) for a in range(d)), "conserved rows: 1, c"
    assert list(kind[:d + 1]) == [0] * (d + 1) and (kind[d + 1:] > 0).all()
    # shear rows span exactly the traceless second-order polynomials, the bulk row is c^2 minus its mean
    second = [c[:, a] * c[:, b] for a in range(d) for b in range(a, d)]
    shear_bulk = M[(kind == 1) | (kind == 2)]
    assert len(shear_bulk) == len(second)
    for poly in second:
        r = poly - poly.mean()
        coef, res, *_ = np.linalg.lstsq(shear_bulk.T.astype(float), r, rcond=None)
        assert np.allclose(shear_bulk.T @ coef, r, atol=1e-12)


def distinct_rates(q):
    r = np.full(27, 1.0)
    r[:q] = 0.6 + 1.3 * (np.arange(q) * 0.6180339887498949 % 1.0)   # every moment its own rate in (0.6, 1.9)
    return r


def perturbed_state(o, seed):
    """a strongly non-equilibrium m_fold: equilibrium weights times random factors"""
    rng = np.random.default_rng(seed)
    o.init()
    o.fold[:] = o.fold * (1.0 + 0.2 * rng.standard_normal(o.fold.shape))


@pytest.mark.parametrize("q,model", [(9, 2), (19, 2), (27, 2), (9, 1), (19, 1), (27, 1)])
def test_mass_and_momentum_are_conserved_with_all_rates_distinct(q, model, oracle_mod):
    d = LATT[q][0]
    g = box_grid((6,) * d, (True,) * d)
    o = oracle_mod.Oracle(d, q, g["nghbr"], 1.7)
    o.set_collision(model, 1.23, distinct_rates(q))
    perturbed_state(o, 7 + q)
    before = o.fold.copy()
    o.step_collide()
    after = o.f
    c = dirs(q)
    assert not np.allclose(after, before, rtol=1e-3), "the collision did nothing"
    assert np.max(np.abs(after.sum(1) - before.sum(1))) < 2e-15 * q
    assert np.max(np.abs(after @ c - before @ c)) < 2e-15 * q


@pytest.mark.parametrize("q,model", [(9, 2), (19, 2), (27, 2), (9, 1), (19, 1)])
def test_equal_rates_are_bgk(q, model, oracle_mod):
    d = LATT[q][0]
    g = box_grid((5,) * d, (True,) * d)
    omega = 1.37
    out = []
    for m in (0, model):
        o = oracle_mod.Oracle(d, q, g["nghbr"], omega)
        o.set_collision(m, omega, np.full(27, omega))
        perturbed_state(o, 3)
        o.step_collide()
        out.append(o.f.copy())
    assert np.max(np.abs(out[0] - out[1])) < 1e-13


def shear_wave_decay(oracle_mod, q, n, steps, model, omega, rates):
    d = LATT[q][0]
    shape = (4, n) if d == 2 else (4, n, 4)
    g = box_grid(shape, (True,) * d)
    o = oracle_mod.Oracle(d, q, g["nghbr"], omega)
    o.set_collision(model, omega, rates)
    o.init()
    # u_x = U sin(2 pi y / n): start from the equilibrium of that velocity field (second order in u)
    y = g["coords"][:, 1]
    U = 1e-3
    ux = U * np.sin(2 * np.pi * (y + 0.5) / n)
    c = dirs(q)
    w = o.fold[0].copy()                                    # init: rho = 1, u = 0 -> the weights
    cu = ux[:, None] * c[None, :, 0]
    o.fold[:] = w[None, :] * (1 + 3 * cu + 4.5 * cu * cu - 1.5 * (ux * ux)[:, None])
    amp = []
    for s in range(steps + 1):
        o.update_moments()
        amp.append(2 * np.mean(o.vars[:, 0] * np.sin(2 * np.pi * (y + 0.5) / n)))
        o.step(1)
    amp = np.array(amp)
    k2 = (2 * np.pi / n) ** 2
    return -np.log(amp[steps] / amp[steps // 2]) / (k2 * (steps - steps // 2))   # measured viscosity


@pytest.mark.parametrize("q", [9, 19, 27])
def test_the_shear_rate_alone_sets_the_viscosity(q, oracle_mod):
    M, norm, kind = basis(q)
    n, steps = 32, 200
    results = {}
    for s_shear in (1.2, 1.7):
        nu = (1.0 / s_shear - 0.5) / 3.0
        nu_bgk = shear_wave_decay(oracle_mod, q, n, steps, 0, s_shear, np.full(27, s_shear))
        rates = np.full(27, s_shear)
        rates[:q][kind == 2] = 1.05
        ghost = np.nonzero(kind == 3)[0]
        rates[ghost] = np.array([1.1, 1.4, 1.95, 0.9])[np.arange(len(ghost)) % 4]
        nu_mrt = shear_wave_decay(oracle_mod, q, n, steps, 2, s_shear, rates)
        assert abs(nu_bgk / nu - 1) < 5e-3, (nu_bgk, nu)
        assert abs(nu_mrt / nu - 1) < 5e-3, (nu_mrt, nu)
        results[s_shear] = nu_mrt
    assert results[1.2] / results[1.7] == pytest.approx(((1 / 1.2 - 0.5) / (1 / 1.7 - 0.5)), rel=1e-2)


@pytest.mark.parametrize("omega", [1.0, 1.6])
def test_trt_magic_parameter_on_the_reference_poiseuille_case(omega, oracle_mod):
    """The reference's own test/poiseuille/poiseuille.json (32^2, forcing through the in-/outlet equilibria, bounce-back walls) with the
    TRT operator.  With a body force, Lambda = 3/16 puts the bounce-back wall exactly half-way between the nodes; the reference drives
    the channel by a density jump instead (solver.cpp:626-696), so the position is not exact here, but it must be clearly closer to
    half-way than BGK's at the same viscosity, on both sides of the viscosity at which BGK happens to have Lambda = 3/16.  The wall
    position is where the parabola through the three central nodes of the steady profile vanishes."""
    from casebuilder import load_golden
    from lbm_b200.cases import trt_omega_minus
    spec = load_golden("poiseuille")
    x, y, h = spec.center[:, 0], spec.center[:, 1], spec.cell_length
    offset = {}
    for model in (0, 1):
        o = spec.apply_to(oracle_mod.Oracle(spec.ndim, spec.ndist, spec.nghbr, omega))
        o.set_collision(model, trt_omega_minus(omega), np.full(27, omega))
        o.init()
        o.step(20000)
        o.update_moments()
        col = np.abs(x - np.sort(x)[len(x) // 2]) < 1e-12
        idx = np.argsort(y[col])
        yy, uu = y[col][idx], o.vars[col, 0][idx]
        m = len(yy) // 2
        coef = np.polyfit(yy[m - 1:m + 2], uu[m - 1:m + 2], 2)
        assert np.max(np.abs(uu - np.polyval(coef, yy))) < 1e-4 * uu.max(), "the steady profile is not a parabola"
        roots = np.sort(np.roots(coef))
        offset[model] = max(abs(roots[0] - spec.bbmin[1]), abs(spec.bbmax[1] - roots[1])) / h   # in cells
    assert offset[1] < 0.6 * offset[0] and offset[1] < 4e-3, offset
