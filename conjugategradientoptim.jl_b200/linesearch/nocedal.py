"""Strong-Wolfe bracketing + bisection zoom — host mirror of src/linesearch/nocedal.jl
(Algorithms 3.5 / 3.6 of Nocedal & Wright 2006).  Scalar-only: every vector touch is one
`evalϕdϕ_` kernel launch or a cached dot."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..cg_types import LineSearchConfig
from ..cg_utils import evalϕdϕ_
from ..device import dot

f64 = np.float64


@dataclass(frozen=True)
class StrongWolfeBisection(LineSearchConfig):
    """nocedal.jl:3-11"""
    c1: float
    c2: float
    a_max_growth_factor: float
    max_iters: int
    zoom_max_iters: int


def setupStrongWolfeBisection(c1, c2, *, a_max_growth_factor=2.0, max_iters=1000,
                              zoom_max_iters=100) -> StrongWolfeBisection:
    """nocedal.jl:14-30"""
    assert 0.0 < c1 < c2 < 1.0                                      # :22
    assert max_iters >= 0                                           # :24
    assert zoom_max_iters >= 0                                      # :25
    assert a_max_growth_factor > 1                                  # :26
    return StrongWolfeBisection(float(c1), float(c2), float(a_max_growth_factor), int(max_iters),
                                int(zoom_max_iters))


def linesearch_(info, config: StrongWolfeBisection, fdf_, f_x, df_x, a_initial):
    """linesearch! (nocedal.jl:33-158) -> (f_xp, a_star, fdf_evals_ran, status)"""
    max_iters, zoom_max_iters = config.max_iters, config.zoom_max_iters
    c1, c2 = f64(config.c1), f64(config.c2)
    a_max_growth_factor = f64(config.a_max_growth_factor)
    xp, df_xp, x, u = info.xp, info.df_xp, info.x, info.u
    a_initial = f64(a_initial)

    if not (0.0 < a_initial and np.isfinite(a_initial)):            # :49-52
        a_initial = f64(1.0)

    ϕ_0 = f64(f_x)
    info.hint_first_trial(a_initial)        # lets a deferred updatedir! ride on the first trial
    dϕ_0 = dot(df_x, u)                                             # :56
    if dϕ_0 > 0.0:                                                  # :57-63
        return ϕ_0, f64(0.0), 0, "non_descent_search_direction"

    a_prev = f64(0.0)
    ϕ_a_prev = ϕ_0
    a = a_initial
    ϕ_a = ϕ_0
    dϕ_a = dϕ_0
    a_max = a * a_max_growth_factor
    fdf_evals_ran = 0
    non_initial_iter = False
    with np.errstate(all="ignore"):
        for _ in range(max_iters):                                  # :76
            ϕ_a, dϕ_a = evalϕdϕ_(xp, df_xp, fdf_, a, x, u)          # :78
            fdf_evals_ran += 1

            chk1 = ϕ_a > ϕ_0 + c1 * a * dϕ_0                        # :81
            chk2 = ϕ_a >= ϕ_a_prev                                  # :82
            if chk1 or (chk2 and non_initial_iter):                 # :83-105  zoom(a_prev, a)
                return zoom_(xp, df_xp, fdf_, x, u, a_prev, a, ϕ_a_prev, ϕ_0, dϕ_0, c1, c2,
                             fdf_evals_ran, zoom_max_iters)

            if abs(dϕ_a) <= -c2 * dϕ_0:                             # :107-110
                return ϕ_a, a, fdf_evals_ran, "success"

            if dϕ_a >= 0:                                           # :112-134  zoom(a, a_prev)
                return zoom_(xp, df_xp, fdf_, x, u, a, a_prev, ϕ_a, ϕ_0, dϕ_0, c1, c2,
                             fdf_evals_ran, zoom_max_iters)

            a_prev = a                                              # :137-139
            ϕ_a_prev = ϕ_a
            non_initial_iter = True

            a_max = a * a_max_growth_factor                         # :143
            if a > a_max:                                           # :144-149
                return ϕ_a, a, fdf_evals_ran, "linesearch_a_max_overflow"
            a = (a_max + a) / 2                                     # :150

    return ϕ_a, a, fdf_evals_ran, "linesearch_max_iters_reached"    # :157


def zoom_(xp, df_xp, fdf_, x, u, a_lb, a_ub, ϕ_a_lb, ϕ_0, dϕ_0, c1, c2, fdf_evals_ran, max_iters):
    """zoom! (nocedal.jl:162-209)"""
    a = f64(0.0)
    ϕ_a = f64(0.0)
    dϕ_a = f64(0.0)
    for _ in range(max_iters):
        a = (a_lb + a_ub) / 2                                       # :187
        ϕ_a, dϕ_a = evalϕdϕ_(xp, df_xp, fdf_, a, x, u)              # :190
        fdf_evals_ran += 1
        if (ϕ_a > ϕ_0 + c1 * a * dϕ_0) or (ϕ_a >= ϕ_a_lb):          # :193
            a_ub = a
        else:
            if abs(dϕ_a) <= -c2 * dϕ_0:                             # :196
                return ϕ_a, a, fdf_evals_ran, "success"
            if dϕ_a * (a_ub - a_lb) >= 0:                           # :200
                a_ub = a_lb
            a_lb = a
            ϕ_a_lb = ϕ_a
    return ϕ_a, a, fdf_evals_ran, "zoom_max_iters_reached"          # :208
