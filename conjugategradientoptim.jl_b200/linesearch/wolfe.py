"""Weak-Wolfe / Yuan-Wei-Lu bisection line search — host mirror of src/linesearch/wolfe.jl,
including its quirks: the mid-search reset `u ← −df_x` that does not
recompute dϕ_0 (:123-129) and the un-returned tuple at :131."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np

from ..cg_types import LineSearchConfig
from ..cg_utils import evalϕdϕ_
from ..device import dot

f64 = np.float64


@dataclass(frozen=True)
class WolfeBisection(LineSearchConfig):
    """wolfe.jl:6-11"""
    condition: Any
    max_iters: int
    max_step_size: float
    feasibility_max_iters: int


@dataclass(frozen=True)
class YuanWeiLuWolfe:
    """wolfe.jl:213-217"""
    c1: float
    c2: float
    δ1: float


@dataclass(frozen=True)
class Wolfe:
    """wolfe.jl:259-262"""
    c1: float
    c2: float


def _jl_min(a, b):
    if a != a:
        return a
    if b != b:
        return b
    return a if a < b else b


def evalwolfeconditions(condition, ϕ_a, dϕ_a, a, u, ϕ_0, dϕ_0):
    """evalwolfeconditions (wolfe.jl:219-251 YuanWeiLuWolfe, :264-294 Wolfe) -> (chk1, chk2)"""
    if isinstance(condition, YuanWeiLuWolfe):
        c1, c2, δ1 = f64(condition.c1), f64(condition.c2), f64(condition.δ1)
        assert 0.0 < δ1 < c1 < c2 < 1.0                             # :233
        norm_u_sq = dot(u, u)                                       # :240
        RHS1 = ϕ_0 + c1 * a * dϕ_0 + a * _jl_min(-δ1 * dϕ_0, c1 * a * norm_u_sq / 2)   # :243
        chk1 = ϕ_a <= RHS1
        RHS2 = c2 * dϕ_0 + _jl_min(-δ1 * dϕ_0, c1 * a * norm_u_sq)                     # :247
        chk2 = dϕ_a >= RHS2
        return bool(chk1), bool(chk2)
    if isinstance(condition, Wolfe):
        c1, c2 = f64(condition.c1), f64(condition.c2)
        assert 0.0 < c1 < c2 < 1.0                                  # :278
        RHS1 = ϕ_0 + c1 * a * dϕ_0                                  # :285
        chk1 = ϕ_a <= RHS1
        RHS2 = c2 * dϕ_0                                            # :289
        chk2 = dϕ_a >= RHS2
        return bool(chk1), bool(chk2)
    raise TypeError(f"no evalwolfeconditions method for {type(condition).__name__}")


def findfeasiblestepsize_(df_xp, xp, fdf_, fdf_evals_ran, a, x, u, reduction_factor, lb, *, max_iters=300):
    """findfeasiblestepsize! (wolfe.jl:171-207) -> (ϕ_a, dϕ_a, a, fdf_evals_ran, flag)"""
    assert 0.0 < reduction_factor < 1.0                             # :185
    if lb > a:                                                      # :186-188
        return f64(0.0), f64(0.0), a, fdf_evals_ran, "bisection_lower_bound_larger_than_proposed_step"
    ϕ_a, dϕ_a = evalϕdϕ_(xp, df_xp, fdf_, a, x, u)                  # :191
    fdf_evals_ran += 1
    it = 1
    while a > lb and it < max_iters:                                # :195
        if np.isfinite(ϕ_a) and np.isfinite(dϕ_a):
            return ϕ_a, dϕ_a, a, fdf_evals_ran, "feasible"
        a = a * reduction_factor                                    # :200
        ϕ_a, dϕ_a = evalϕdϕ_(xp, df_xp, fdf_, a, x, u)
        fdf_evals_ran += 1
        it += 1
    return ϕ_a, dϕ_a, a, fdf_evals_ran, "infeasible"                # :206


def linesearch_(info, config: WolfeBisection, fdf_, f_x, df_x, a_initial):
    """linesearch! (wolfe.jl:13-165)"""
    reduction_factor = f64(0.5)                                     # :23
    growth_factor = f64(2)                                          # :24
    xp, df_xp, x, u = info.xp, info.df_xp, info.x, info.u
    condition, max_step_size = config.condition, f64(config.max_step_size)
    max_iters, feasibility_max_iters = config.max_iters, config.feasibility_max_iters
    a_initial = f64(a_initial)
    with np.errstate(all="ignore"):
        if not (max_step_size > a_initial > 0.0):                   # :30-32
            a_initial = min(f64(1.0), max_step_size / 2)

        ϕ_0 = f64(f_x)
        if not np.isfinite(ϕ_0):                                    # :36-38
            return ϕ_0, f64(0.0), 0, "accepted_non_finite_iterate"

        info.hint_first_trial(a_initial)
        dϕ_0 = dot(df_x, u)                                         # :40
        if dϕ_0 > 0.0:
            return ϕ_0, f64(0.0), 0, "non_descent_search_direction"

        a = a_initial
        fdf_evals_ran = 0
        lb = f64(0.0)
        ub = f64(np.inf)

        ϕ_a, dϕ_a, a, fdf_evals_ran, status_flag = findfeasiblestepsize_(   # :51-62
            df_xp, xp, fdf_, fdf_evals_ran, a, x, u, reduction_factor, f64(0.0),
            max_iters=feasibility_max_iters)
        if status_flag != "feasible":
            return ϕ_0, f64(0.0), 0, "cannot_find_initial_feasible_step"

        for _ in range(max_iters):                                  # :67
            valid_large_step, valid_small_step = evalwolfeconditions(
                condition, ϕ_a, dϕ_a, a, u, ϕ_0, dϕ_0)              # :70-78
            if (not valid_large_step) or (not valid_small_step):
                if not valid_large_step:
                    ub = a                                          # :86
                    a = (lb + ub) / 2                               # :95
                else:
                    lb = a                                          # :98
                    if not np.isfinite(ub):
                        a = growth_factor * a                       # :102
                        if a > max_step_size:                       # :104-112
                            return ϕ_0, f64(0.0), 0, "max_step_length_reached"
                    else:
                        a = (lb + ub) / 2                           # :114
                if not (lb < a < ub):                               # :122
                    # !isapprox(norm(u+df_x), 0) with rtol=√eps, atol=0  ⇔  norm != 0 (or NaN)
                    if not (info.norm_u_plus_g() == 0.0):           # :123
                        lb = f64(0.0)
                        ub = f64(np.inf)
                        a = a_initial
                        info.reset_direction()                      # :129  u[:] = -df_x
                    # else: wolfe.jl:131 builds a tuple and does not return it
                ϕ_a, dϕ_a, a, fdf_evals_ran, status_flag = findfeasiblestepsize_(   # :141-152
                    df_xp, xp, fdf_, fdf_evals_ran, a, x, u, reduction_factor, lb,
                    max_iters=feasibility_max_iters)
                if status_flag != "feasible":
                    return ϕ_0, f64(0.0), 0, "cannot_find_feasible_step"   # :157
            else:
                return ϕ_a, a, fdf_evals_ran, "success"             # :160
    return ϕ_a, a, fdf_evals_ran, "linesearch_max_iters_reached"    # :164
