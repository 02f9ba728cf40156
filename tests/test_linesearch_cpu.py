"""Zoom line search (core/linesearch.py, the optax.scale_by_zoom_linesearch link of planner.py:2373-2374):
the returned step satisfies the strong Wolfe conditions on analytic objectives."""
import math

import pytest

from pyrddlgym_jax_b200.core.linesearch import zoom_kwargs, zoom_linesearch


def _wolfe(phi, eta, f0, g0, slope_rtol, curv_rtol):
    f, g = phi(eta)
    return f <= f0 + slope_rtol * eta * g0 + 1e-12 and abs(g) <= curv_rtol * abs(g0) + 1e-12


CASES = {
    "quadratic": (lambda e: ((e - 3.0) ** 2, 2.0 * (e - 3.0))),
    # More & Thuente test functions (also Nocedal & Wright's examples)
    "rational": (lambda e: (-e / (e * e + 2.0), (e * e - 2.0) / (e * e + 2.0) ** 2)),
    "quartic": (lambda e: ((e + 0.2) ** 5 - 2.0 * (e + 0.2) ** 4,
                           5.0 * (e + 0.2) ** 4 - 8.0 * (e + 0.2) ** 3)),
    "steep": (lambda e: (math.exp(4.0 * e) - 10.0 * e, 4.0 * math.exp(4.0 * e) - 10.0)),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("init", [1e-3, 1.0, 50.0])
@pytest.mark.parametrize("curv_rtol", [0.9, 0.1])
def test_zoom_returns_a_wolfe_point(name, init, curv_rtol):
    phi = CASES[name]
    f0, g0 = phi(0.0)
    assert g0 < 0
    calls = []

    def counted(e):
        calls.append(e)
        return phi(e)
    eta, f, evals, ok = zoom_linesearch(counted, f0, g0, init, max_linesearch_steps=40,
                                        curv_rtol=curv_rtol)
    assert ok, (name, init, eta, evals)
    assert evals == len(calls) <= 40
    assert eta > 0 and _wolfe(phi, eta, f0, g0, 1e-4, curv_rtol)
    assert abs(f - phi(eta)[0]) < 1e-12


def test_zoom_budget_and_failure_modes():
    phi = CASES["rational"]
    f0, g0 = phi(0.0)
    # a budget of one evaluation: the safe step is a point that decreases the value (or zero)
    eta, f, evals, ok = zoom_linesearch(phi, f0, g0, 100.0, max_linesearch_steps=1, curv_rtol=1e-5)
    assert evals == 1 and not ok and f <= f0
    # an almost flat start (slope -5e-7): the curvature test is out of reach of the step-size precision;
    # the search stops inside its budget at a point that decreased the value
    flat = lambda e: ((e + 0.004) ** 5 - 2.0 * (e + 0.004) ** 4,                          # noqa: E731
                      5.0 * (e + 0.004) ** 4 - 8.0 * (e + 0.004) ** 3)
    f0f, g0f = flat(0.0)
    eta, f, evals, ok = zoom_linesearch(flat, f0f, g0f, 1.0, max_linesearch_steps=40)
    assert evals <= 40 and f < f0f - 2.0 and abs(eta - 1.596) < 1e-3
    # not a descent direction: no step
    assert zoom_linesearch(phi, f0, +1.0, 1.0, max_linesearch_steps=5)[0] == 0.0
    # a non-finite trial is treated as a failed decrease and the interval shrinks towards zero
    bad = lambda e: (float("nan"), float("nan")) if e > 0.5 else CASES["quadratic"](e)   # noqa: E731
    f0, g0 = bad(0.0)
    eta, f, evals, ok = zoom_linesearch(bad, f0, g0, 4.0, max_linesearch_steps=30)
    assert 0.0 < eta <= 0.5 and f < f0
    # max_learning_rate caps the step
    eta, *_ = zoom_linesearch(CASES["quadratic"], 9.0, -6.0, 1.0, max_linesearch_steps=20,
                              max_learning_rate=0.25, curv_rtol=0.1)
    assert eta <= 0.25
    with pytest.raises(TypeError):
        zoom_kwargs({"max_steps": 3})
    with pytest.raises(TypeError):
        zoom_kwargs({})
    assert zoom_kwargs({"max_linesearch_steps": 7})["curv_rtol"] == 0.9
