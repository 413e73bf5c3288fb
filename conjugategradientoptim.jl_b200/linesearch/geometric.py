"""Armijo geometric grow/shrink line search — host mirror of src/linesearch/geometric.jl.

Replicated bug-for-bug: geometricsearch! returns the PREVIOUS (ϕ, a) as
:success while info.xp / info.df_xp hold the rejected trial, which the engine then adopts
(optim.jl:136-139); and the redundant second evaluation at :78."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np

from ..cg_types import LineSearchConfig
from ..cg_utils import evalϕdϕ_
from ..device import dot
from .wolfe import findfeasiblestepsize_

f64 = np.float64


class GeometricStepTrait:            # geometric.jl:3
    pass


class DivideGeometricStep(GeometricStepTrait):     # :4
    pass


class MultiplyGeometricStep(GeometricStepTrait):   # :5
    pass


def getgeometricstep(method: GeometricStepTrait, a, ρ):
    """geometric.jl:7-13"""
    if isinstance(method, DivideGeometricStep):
        return a / ρ
    return a * ρ


@dataclass(frozen=True)
class Backtracking(LineSearchConfig):
    """geometric.jl:15-20"""
    condition: Any
    discount_factor: float
    max_iters: int
    feasibility_max_iters: int


@dataclass(frozen=True)
class Armijo:
    """geometric.jl:159-162"""
    c1: float


def evalbacktrackcondition(condition: Armijo, ϕ_a, a, ϕ_0, dϕ_0) -> bool:
    """geometric.jl:164-186"""
    c1 = f64(condition.c1)
    assert 0.0 < c1 < 1.0                                           # :174
    if not np.isfinite(ϕ_0) or not np.isfinite(ϕ_a) or not np.isfinite(a):   # :177-179
        return False
    LHS1 = ϕ_0 - ϕ_a
    return bool(LHS1 >= -c1 * a * dϕ_0)                             # :182-183


def linesearch_(info, config: Backtracking, fdf_, f_x, df_x, a_initial):
    """linesearch! (geometric.jl:22-100)"""
    xp, df_xp, x, u = info.xp, info.df_xp, info.x, info.u
    discount_factor, max_iters = f64(config.discount_factor), config.max_iters
    condition = config.condition
    feasibility_max_iters = config.feasibility_max_iters
    with np.errstate(all="ignore"):
        ϕ_0 = f64(f_x)
        if not np.isfinite(ϕ_0):                                    # :39-41
            return ϕ_0, f64(0.0), 0, "accepted_non_finite_iterate"
        a = f64(a_initial)
        if np.isfinite(a):
            info.hint_first_trial(a)
        dϕ_0 = dot(df_x, u)                                         # :43
        if dϕ_0 > 0.0:
            return ϕ_0, f64(0.0), 0, "non_descent_search_direction"
        fdf_evals_ran = 0
        if not np.isfinite(a):                                      # :50-53
            a = abs(ϕ_0) / dot(u, u)
        if not np.isfinite(a):                                      # :54-57
            a = f64(1.0)
        reduction_factor = f64(0.5)
        ϕ_a, dϕ_a, a, fdf_evals_ran, status_flag = findfeasiblestepsize_(   # :60-71
            df_xp, xp, fdf_, fdf_evals_ran, a, x, u, reduction_factor, f64(0.0),
            max_iters=feasibility_max_iters)
        if status_flag != "feasible":
            return ϕ_0, f64(0.0), 0, "cannot_find_initial_feasible_step"    # :74

        ϕ_a, _ = evalϕdϕ_(xp, df_xp, fdf_, a, x, u)                 # :78 (redundant re-evaluation)
        fdf_evals_ran += 1
        valid_step = evalbacktrackcondition(condition, ϕ_a, a, ϕ_0, dϕ_0)   # :81
        method = DivideGeometricStep() if valid_step else MultiplyGeometricStep()   # :83-97
        return geometricsearch_(xp, df_xp, fdf_, a, x, u, condition, max_iters, discount_factor,
                                method, fdf_evals_ran, ϕ_a, ϕ_0, dϕ_0)


def geometricsearch_(xp, df_xp, fdf_, a, x, u, condition, max_iters, discount_factor, geta_method,
                     fdf_evals_ran, ϕ_a, ϕ_0, dϕ_0):
    """geometricsearch! (geometric.jl:102-152)"""
    a_prev = a
    ϕ_a_prev = ϕ_a
    for _ in range(max_iters):
        a = getgeometricstep(geta_method, a, discount_factor)       # :127
        if not np.isfinite(a):                                      # :128-130
            return ϕ_a_prev, a_prev, fdf_evals_ran, "non_finite_step_proposed"
        if a == a_prev:                                             # :132-134
            return ϕ_a_prev, a_prev, fdf_evals_ran, "proposed_step_same_as_current_step"
        ϕ_a, _ = evalϕdϕ_(xp, df_xp, fdf_, a, x, u)                 # :137
        fdf_evals_ran += 1
        valid_step = evalbacktrackcondition(condition, ϕ_a, a, ϕ_0, dϕ_0)
        if not valid_step:                                          # :140-144
            return ϕ_a_prev, a_prev, fdf_evals_ran, "success"
        a_prev = a
        ϕ_a_prev = ϕ_a
    return ϕ_a, a, fdf_evals_ran, "linesearch_max_iters_reached"    # :151
