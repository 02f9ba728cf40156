"""Zoom line search link of the optimiser chain (planner.py:2373-2374: optax.scale_by_zoom_linesearch).

optax is a third-party dependency of the reference (optax >= 0.2.7, not vendored), so
this restates the published algorithm it implements -- the interval / zoom search of Nocedal & Wright,
Numerical Optimization, Algorithms 3.5 and 3.6, with the approximate-decrease test of Hager & Zhang --
under optax's keyword names and defaults.  The optimiser's update u (already scaled by -learning_rate) is
the search direction; the link returns the step size eta for which  phi(eta) = loss(params + eta u)
satisfies

    sufficient decrease   phi(eta) <= phi(0) + slope_rtol * eta * phi'(0)          (Armijo), or the
    approximate decrease  phi'(eta) <= (2 slope_rtol - 1) phi'(0)  and  phi(eta) <= phi(0) + approx_dec_rtol |phi(0)|
    small curvature       |phi'(eta)| <= curv_rtol * |phi'(0)|

Every trial costs one relaxed rollout + adjoint on the device (same noise: the reference passes the same
sim_state to value_fn, planner.py:2888-2892); the scalar logic below runs on the host between them.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

ZOOM_DEFAULTS = {"max_linesearch_steps": None, "max_learning_rate": None, "tol": 0.0,
                 "increase_factor": 2.0, "slope_rtol": 1e-4, "curv_rtol": 0.9,
                 "approx_dec_rtol": 1e-6, "stepsize_precision": 1e-5,
                 "initial_guess_strategy": "keep", "verbose": False}


def zoom_kwargs(kw: Dict) -> Dict:
    extra = set(kw) - set(ZOOM_DEFAULTS)
    if extra:
        raise TypeError(f"line_search_kwargs: unknown arguments {sorted(extra)}")
    out = dict(ZOOM_DEFAULTS)
    out.update(kw)
    if out["max_linesearch_steps"] is None:
        raise TypeError("line_search_kwargs: max_linesearch_steps is required")
    if out["initial_guess_strategy"] not in ("keep", "one"):
        raise ValueError("initial_guess_strategy must be 'keep' or 'one'")
    return out


def _cubic_min(a, fa, ga, b, fb, c, fc) -> Optional[float]:
    """Minimiser of the cubic through (a, fa) with slope ga, (b, fb) and (c, fc)."""
    db, dc = b - a, c - a
    denom = (db * dc) ** 2 * (db - dc)
    if denom == 0.0:
        return None
    d1b, d1c = fb - fa - ga * db, fc - fa - ga * dc
    A = (dc * dc * d1b - db * db * d1c) / denom
    B = (-dc ** 3 * d1b + db ** 3 * d1c) / denom
    rad = B * B - 3.0 * A * ga
    if A == 0.0 or rad < 0.0:
        return None
    return a + (-B + math.sqrt(rad)) / (3.0 * A)


def _quad_min(a, fa, ga, b, fb) -> Optional[float]:
    """Minimiser of the parabola through (a, fa) with slope ga and (b, fb)."""
    db = b - a
    B = (fb - fa - ga * db) / (db * db)
    if B <= 0.0:
        return None
    return a - ga / (2.0 * B)


def zoom_linesearch(phi: Callable[[float], Tuple[float, float]], f0: float, slope0: float,
                    init_stepsize: float, max_linesearch_steps: int,
                    max_learning_rate: Optional[float] = None, tol: float = 0.0,
                    increase_factor: float = 2.0, slope_rtol: float = 1e-4, curv_rtol: float = 0.9,
                    approx_dec_rtol: Optional[float] = 1e-6, stepsize_precision: float = 1e-5,
                    initial_guess_strategy: str = "keep", verbose: bool = False):
    """Returns (eta, f(eta), number of evaluations, converged).  ``phi(eta) -> (value, slope)``."""
    if not (slope0 < 0.0) or not math.isfinite(f0):
        # not a descent direction (or a broken objective): take no step, as a failed search does
        return 0.0, f0, 0, False

    def decrease_ok(eta, f, g):
        if f <= f0 + slope_rtol * eta * slope0:
            return True
        return approx_dec_rtol is not None and g <= (2.0 * slope_rtol - 1.0) * slope0 \
            and f <= f0 + approx_dec_rtol * abs(f0)

    def curvature_ok(g):
        return abs(g) <= curv_rtol * abs(slope0) or abs(g) <= tol

    evals = 0
    best = (0.0, f0)                    # safe step: the best point seen that decreases the value
    eta_prev, f_prev, g_prev = 0.0, f0, slope0
    eta = float(init_stepsize)
    if max_learning_rate is not None:
        eta = min(eta, float(max_learning_rate))
    lo = hi = None
    # ---- interval search (Algorithm 3.5)
    while evals < max_linesearch_steps:
        f, g = phi(eta)
        evals += 1
        ok = math.isfinite(f) and math.isfinite(g)
        if ok and f < best[1] and decrease_ok(eta, f, g):
            best = (eta, f)
        if not ok or not decrease_ok(eta, f, g) or (evals > 1 and f >= f_prev):
            lo, hi = (eta_prev, f_prev, g_prev), (eta, f if ok else math.inf, g if ok else 0.0)
            break
        if curvature_ok(g):
            return eta, f, evals, True
        if g >= 0.0:
            lo, hi = (eta, f, g), (eta_prev, f_prev, g_prev)
            break
        if max_learning_rate is not None and eta >= max_learning_rate:
            return eta, f, evals, False
        eta_prev, f_prev, g_prev = eta, f, g
        eta = eta * increase_factor
        if max_learning_rate is not None:
            eta = min(eta, float(max_learning_rate))
    if lo is None:
        return best[0], best[1], evals, False
    # ---- zoom (Algorithm 3.6): lo is the end with the lower value that satisfies the decrease test
    rec = None                          # the previous trial point, for the cubic fit
    while evals < max_linesearch_steps:
        a, fa, ga = lo
        b, fb, _ = hi
        width = abs(b - a)
        if width <= stepsize_precision:
            break
        cand = None
        if rec is not None and math.isfinite(fb) and math.isfinite(rec[1]):
            cand = _cubic_min(a, fa, ga, b, fb, rec[0], rec[1])
        lo_edge, hi_edge = min(a, b), max(a, b)
        if cand is None or not (lo_edge + 0.2 * width <= cand <= hi_edge - 0.2 * width):
            cand = _quad_min(a, fa, ga, b, fb) if math.isfinite(fb) else None
            if cand is None or not (lo_edge + 0.1 * width <= cand <= hi_edge - 0.1 * width):
                cand = 0.5 * (a + b)
        f, g = phi(cand)
        evals += 1
        ok = math.isfinite(f) and math.isfinite(g)
        if ok and f < best[1] and decrease_ok(cand, f, g):
            best = (cand, f)
        if not ok or not decrease_ok(cand, f, g) or f >= fa:
            rec = (b, fb)
            hi = (cand, f if ok else math.inf, g if ok else 0.0)
        else:
            if curvature_ok(g):
                return cand, f, evals, True
            if g * (b - a) >= 0.0:
                rec = (b, fb)
                hi = lo
            else:
                rec = (a, fa)
            lo = (cand, f, g)
    return best[0], best[1], evals, False
