"""CG for nonlinear systems g(x) = 0 — host mirror of src/engine/solve_system.jl (Yuan, Wang &
Sheng 2019, Alg. 3.1).  `fdf!` returns a merit value f and writes g(x) where a gradient would go.

    reference (solve_system.jl)                        here
    :80-87   copy x0 twice, fdf!, norm                 DeviceLineSearchContainer + solvesys_begin()
    :28-56   linesearch!: evalϕdϕ! + norm per trial    info.eval_trial(a)             [1 launch set / trial]
    :171-177 updateiteratesolvesys! (:237-253)         info.solvesys_project(m)       [1 BLAS-1 launch
    :179     f_x_next = fdf!(info.df_xp, x_next)           + 1 launch set, fused pack]
    :196-208 swap x / x_next, df_x[:] = …, info.x[:] = … info.solvesys_accept()         [pointer swaps]
    :201     getβ                                      host arithmetic on the pack    [0]
    :212     updatedir!                                deferred into the next trial   [0, fused]

Quirks, kept by default (the drop-in must give what the reference gives):
  * linesearch! returns the 0-based index of the accepted trial as `fdf_evals_ran` (:50);
  * its failure return (:54) reads the loop variable outside the loop — an UndefVarError in Julia,
    so the reference never reaches its own :linesearch_failed branch (:131-142); the evident
    intent (that status) is what is returned here;
  * updateiteratesolvesys! is handed `x_next`, which after the first swap holds the iterate
    BEFORE the current one: from the second iteration on the method is not Alg. 3.1 any more and
    in practice diverges.  `fix_stale_iterate=True` (not in the reference) projects from the
    current iterate, as published.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from ..cg_flavours import getβ, initializeLineSearchContainer_, initializeβ, updatedir_
from ..cg_types import CGConfig, CGβConfig, Results, resizetrace_, setuptrace, updateresult_, updatetrace_
from ..cg_utils import evalϕdϕ_
from ..device import dot

f64 = np.float64


@dataclass(frozen=True)
class LinesearchSolveSys:
    """solve_system.jl:7-12 — parameters of eqn 18 of (Yuan 2019)"""
    ρ: float      # 0 < ρ < 1
    σ: float      # σ > 0
    s: float      # s > 0
    max_iters: int


def setupLinesearchSolveSys(s, *, σ=0.5, ρ=0.95, max_iters=None) -> LinesearchSolveSys:
    """solve_system.jl:14-26"""
    if max_iters is None:
        max_iters = round(math.log(1e-6) / math.log(ρ))            # round(Int, log(ρ, 1e-6))  :18
    assert 0.0 < ρ < 1.0                                            # :21
    assert ρ > 0.0                                                  # :22
    assert s > 0.0                                                  # :23
    return LinesearchSolveSys(float(ρ), float(σ), float(s), int(max_iters))


def linesearch_solvesys_(info, config: LinesearchSolveSys, fdf_):
    """linesearch! (solve_system.jl:28-56) -> (f_xp, norm_df_xp, a, i, success_flag)"""
    max_iters = config.max_iters
    ρ, σ, a0 = config.ρ, f64(config.σ), config.s
    xp, df_xp, x, u = info.xp, info.df_xp, info.x, info.u
    info.hint_first_trial(a0)               # the first trial is always a = s·ρ⁰
    norm_u_sq = dot(u, u)                                           # :39
    with np.errstate(all="ignore"):
        for i in range(max_iters):                                  # :41
            a = f64(a0 * math.pow(ρ, i))                            # :42  a0*ρ^i
            f_xp, dϕ_xp = evalϕdϕ_(xp, df_xp, fdf_, a, x, u)        # :44
            norm_df_xp = info.norm_df_xp()                          # :47
            if not (-dϕ_xp < σ * a * norm_df_xp * norm_u_sq):       # :48
                return f_xp, norm_df_xp, a, i, True                 # :50
    return None, f64(np.nan), f64(np.nan), max_iters - 1, False     # :54 (see module docstring)


def solvesystem(fdf_, x_initial, config: CGConfig, linesearch_config: LinesearchSolveSys, *,
                fix_stale_iterate: bool = False, fuse_direction: bool = True, beta_form: str = "fused") -> Results:
    """solvesystem (src/engine/solve_system.jl:64-239)."""
    if not hasattr(fdf_, "make_workspace"):
        raise TypeError("fdf! must be a device objective handle: this package has no CPU fallback")
    assert isinstance(config, CGConfig) and isinstance(linesearch_config, LinesearchSolveSys)
    assert isinstance(config.β_config, CGβConfig)                   # BT <: CGβConfig  :69
    max_iters, β_config = config.max_iters, config.β_config        # :74-75
    # ## allocate + Step 1: x, x_next = copies of x_initial; f_x = fdf!(df_x, x); norm   :80-87
    info = fdf_.make_workspace(x_initial, lbfgs_m=0, fuse_direction=fuse_direction, beta_form=beta_form)
    info.solvesys_begin()
    x, df_x = info.x, info.df_x
    f_x = info.f_x0
    norm_df_x = info.norm_df_x0
    β = initializeβ(β_config)                                       # :88
    ret = Results(f_x, x, df_x, 0, "incomplete", setuptrace(config.trace_status))      # :93-101
    resizetrace_(ret.trace, max_iters)
    initializeLineSearchContainer_(info, β_config, df_x, x)        # :104-105

    def finish(minimizer_from_trial=False):
        ret.minimizer, ret.gradient = info.download_trial() if minimizer_from_trial else info.download()
        info.close()
        return ret

    with np.errstate(all="ignore"):
        for n in range(1, max_iters + 1):                           # :109
            if norm_df_x < config.ϵ:                                # :112-123
                updateresult_(ret, x, df_x, f_x, n - 1, "success")
                return finish()
            f_xp, norm_df_xp, a_star, fdf_evals_ran, success_flag = linesearch_solvesys_(   # :126-130
                info, linesearch_config, fdf_)
            if not success_flag:                                    # :131-142
                updateresult_(ret, x, df_x, f_x, n - 1, "linesearch_failed")
                return finish()
            if norm_df_xp < config.ϵ:                               # :146-168  return the line-search point
                updateresult_(ret, info.xp, info.df_xp, f_xp, n, "success")
                updatetrace_(ret.trace, f_xp, info.norm_df_xp(), a_star, fdf_evals_ran, n)
                return finish(minimizer_from_trial=True)
            # updateiteratesolvesys!(x_next, info.df_xp, norm_df_xp, a_star, info.u)   :171-177, :237-253
            m = a_star * info.dot_df_xp_u() / (norm_df_xp * norm_df_xp)
            f_x_next, norm_next = info.solvesys_project(m, fix_stale_iterate)           # :246-250, :179
            if not np.isfinite(f_x_next) or not np.isfinite(norm_next):                 # :180-194
                updateresult_(ret, x, df_x, f_x, n - 1, "non_finite_objective_or_gradient_proposed")
                return finish()
            β = getβ(β_config, info.df_xp, df_x, info.u)            # :201-206 (before the swap: same operands)
            info.solvesys_accept(fix_stale_iterate)                 # :196, :207-208
            f_x = f_x_next                                          # :197
            norm_df_x = norm_next                                   # :209
            updatedir_(info.u, df_x, β)                             # :212
            updatetrace_(ret.trace, f_x, norm_df_x, a_star, fdf_evals_ran, n)           # :215-222
    updateresult_(ret, x, df_x, f_x, max_iters, "max_iters_reached")                    # :225-233
    return finish()
