"""Boundary-value expressions of the host mirror (lbm_b200/host/expr.hpp) against the values the reference's exprtk evaluation
produced: test/poisson/poisson2D_helmholtz.json has three expression-valued Dirichlet boundaries; the fixture holds what the
reference wrote into m_vars of their cells (tests/golden/make_golden.py).  Bit for bit."""
import numpy as np
import pytest

from casebuilder import load_golden
from lbm_b200 import host_api


def test_expressions_equal_the_reference_evaluation():
    spec = load_golden("poisson2D_helmholtz")
    bnd = spec.config["solver"]["boundary"]["line"]
    g = spec.golden
    checked = 0
    for k, nm in enumerate(g["surface_names"]):
        val = bnd[str(nm).split("_", 1)[1]]["value"]
        if not isinstance(val, str):
            continue
        cells = g[f"surf{k}_cells"].astype(np.int64)
        mine = host_api.eval_expression(val, spec.center[cells])
        assert np.array_equal(mine, g[f"surf{k}_values"]), f"{nm}: {val}"
        checked += len(cells)
    assert checked == 768


@pytest.mark.parametrize("text,x,want", [
    ("1", (0.3, 0.7), 1.0), ("-x", (0.25, 0.0), -0.25), ("2*x+3*y", (0.5, 0.25), 1.75), ("x^2", (3.0, 0.0), 9.0), ("-x^2", (3.0, 0.0), -9.0),
    ("2^3^2", (0, 0), 512.0), ("(1-y)/(1+y)", (0.0, 0.5), 1.0 / 3.0), ("sqrt(4+pi^2)", (0, 0), float(np.sqrt(4 + np.pi * np.pi))),
    ("cos(pi*x)", (1.0, 0.0), -1.0), ("exp(log(x))", (2.0, 0), float(np.exp(np.log(2.0)))), ("abs(-3.5e0)", (0, 0), 3.5), ("z", (1, 2), 0.0),
])
def test_grammar(text, x, want):
    assert host_api.eval_expression(text, np.array([x], dtype=float))[0] == want


@pytest.mark.parametrize("text", ["", "1+", "foo(1)", "x y", "(1", "sin 1", "1/*2"])
def test_malformed_expressions_are_reported(text):
    with pytest.raises(ValueError) as e:
        host_api.eval_expression(text, np.zeros((1, 2)))
    assert "Invalid math expression" in str(e.value)
