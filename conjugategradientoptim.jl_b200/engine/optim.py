"""Unconstrained minimiser loop + restart wrapper — host mirror of src/engine/optim.jl.

The loop body, the status Symbols, the `iters_ran = n − 1` conventions and the trace are the
reference's; every line that touched a vector there is one C-ABI call here:

    reference (optim.jl)                         here
    :20-26  copy x0, fdf!, norm                  DeviceLineSearchContainer(...)  [1 launch]
    :46     initializeLineSearchContainer!       info.reset_direction()          [1 launch]
    :83     linesearch! → evalϕdϕ! per trial     info.eval_trial(a)              [1 launch / trial]
    :107    norm(info.df_xp)                     sqrt(pack[g⁺·g⁺])               [0]
    :130    getβ (3 temporaries, 3–6 dots)       host arithmetic on the pack     [0]
    :136-140 three n-vector copies               info.accept(): pointer swaps    [0]
    :145    updatedir!                           deferred into the next trial    [0, fused]
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from ..cg_flavours import getβ, initializeLineSearchContainer_, initializeβ, updatedir_
from ..cg_types import (CGConfig, LineSearchConfig, Results, resizetrace_, setuptrace,
                        updateresult_, updatetrace_, βConfig)
from ..linesearch import geometric, nocedal, wolfe
from ..qn_flavours import LBFGS
from .._capi import P_PHI

f64 = np.float64


def linesearch_(info, config: LineSearchConfig, fdf_, f_x, df_x, a_initial):
    """linesearch! — dispatch on the line-search config type (nocedal.jl:33, wolfe.jl:13,
    geometric.jl:22) -> (f_xp, a_star, fdf_evals_ran, status)"""
    if isinstance(config, nocedal.StrongWolfeBisection):
        return nocedal.linesearch_(info, config, fdf_, f_x, df_x, a_initial)
    if isinstance(config, wolfe.WolfeBisection):
        return wolfe.linesearch_(info, config, fdf_, f_x, df_x, a_initial)
    if isinstance(config, geometric.Backtracking):
        return geometric.linesearch_(info, config, fdf_, f_x, df_x, a_initial)
    raise TypeError(f"no linesearch! method for {type(config).__name__}")


class MinimizerRun:
    """State of one `minimizeobjective` call, steppable one iteration at a time (bench.py times
    individual iterations; `minimizeobjective` just drives it to completion)."""

    def __init__(self, fdf_, x_initial, config: CGConfig, linesearch_config: LineSearchConfig,
                 fuse_direction: bool = True, beta_form: str = "fused", quadratic_linesearch: bool = False):
        if not hasattr(fdf_, "make_workspace"):
            raise TypeError(
                "fdf! must be a device objective handle (RosenbrockGPU, SparseLSGPU, LogRegGPU, ..., or "
                "UserObjectiveGPU(n, fdf) for your own fdf!(g, x) written on CUDA tensors): host callbacks "
                "cannot run on the GPU and this package has no CPU fallback")
        assert isinstance(config, CGConfig) and isinstance(linesearch_config, LineSearchConfig)
        # ## parse.                                                         optim.jl:14-17
        self.fdf_ = fdf_
        self.config, self.linesearch_config = config, linesearch_config
        self.max_iters = config.max_iters
        self.β_config = config.β_config
        lbfgs_m = self.β_config.m if isinstance(self.β_config, LBFGS) else 0
        # ## allocate + Step 1: x = copy(x_initial); f_x = fdf!(df_x, x); norm(df_x)   :20-26
        kw = {"quadratic_linesearch": True} if quadratic_linesearch else {}
        if quadratic_linesearch and isinstance(linesearch_config, geometric.Backtracking):
            raise TypeError("the quadratic-aware path needs a line search that accepts its last trial "
                            "(Backtracking adopts a rejected one, geometric.jl:141-144)")
        self.info = fdf_.make_workspace(x_initial, lbfgs_m=lbfgs_m, fuse_direction=fuse_direction,
                                        beta_form=beta_form, **kw)
        info = self.info
        self.x, self.df_x = info.x, info.df_x
        self.f_x = info.f_x0
        self.norm_df_x = info.norm_df_x0
        self.norm_df_xp = f64(np.nan)
        self.β = initializeβ(self.β_config)                                 # :29
        self.fdf_evals_ran = -1
        self.f_x0 = self.f_x                                                # :31
        # ## return container.                                              :34-42
        self.ret = Results(self.f_x, self.x, self.df_x, 0, "incomplete", setuptrace(config.trace_status))
        resizetrace_(self.ret.trace, self.max_iters)
        # ## line search.                                                   :45-47
        initializeLineSearchContainer_(info, self.β_config, self.df_x, self.x)
        self.a_initial = f64(np.nan)
        self.n = 0
        self.done = False

    def _finish(self, i: int, status: str) -> Results:
        updateresult_(self.ret, self.x, self.df_x, self.f_x, i, status)
        self.done = True
        return self.ret

    def step(self):
        """One pass of the `for n = 1:max_iters` body (optim.jl:50-160).  Returns Results when
        the run terminated, else None."""
        assert not self.done
        config, info = self.config, self.info
        self.n += 1
        n = self.n
        if n > self.max_iters:                                              # :162-170
            return self._finish(self.max_iters, "max_iters_reached")

        # check stopping conditions.                                        :53-80
        if np.isfinite(self.f_x) and np.isfinite(self.norm_df_x):
            if self.norm_df_x < config.ϵ:
                if self.f_x <= self.f_x0:
                    return self._finish(n - 1, "success")
                return self._finish(n - 1, "increasing_objective")

        # step 3: linesearch.                                               :83-90
        f_xp, a_star, self.fdf_evals_ran, status_symbol = linesearch_(
            info, self.linesearch_config, self.fdf_, self.f_x, self.df_x, self.a_initial)
        self.a_initial = a_star                                             # :92
        if status_symbol != "success":                                      # :93-104
            return self._finish(n - 1, status_symbol)

        # numerically valid proposed iterate?                               :107-121
        self.norm_df_xp = info.norm_df_xp()
        if getattr(info, "quadratic", False):
            # the trials were evaluated from the parabola ½r·r + a r·v + ½a² v·v, which cancels badly once
            # f has dropped by many orders of magnitude; the accepted step's own ½ Σ r² is now known
            f_xp = f64(info.pack[P_PHI])
        if not np.isfinite(f_xp) or not np.isfinite(self.norm_df_xp):
            return self._finish(n - 1, "non_finite_objective_or_gradient_proposed")

        # step 4 & 5: update iterate and objective-related evaluations.     :130-141
        self.β = getβ(self.β_config, info.df_xp, self.df_x, info.u)
        info.accept()            # x[:] = info.xp; df_x[:] = info.df_xp; info.x[:] = x
        self.f_x = f_xp
        self.norm_df_x = self.norm_df_xp

        # step 5: update search direction for next iteration.               :145
        updatedir_(info.u, self.df_x, self.β)

        updatetrace_(self.ret.trace, self.f_x, self.norm_df_x, a_star, self.fdf_evals_ran, n)   # :152-159
        return None

    def run(self) -> Results:
        while True:
            r = self.step()
            if r is not None:
                return r

    def results(self) -> Results:
        """Materialise Results.minimizer / Results.gradient on the host (one D2H each)."""
        x, g = self.info.download()
        self.ret.minimizer, self.ret.gradient = x, g
        return self.ret


def minimizeobjective(fdf_, x_initial, config: CGConfig, linesearch_config: LineSearchConfig, *,
                      fuse_direction: bool = True, beta_form: str = "fused",
                      quadratic_linesearch: bool = False, keep_workspace: bool = False) -> Results:
    """minimizeobjective (src/engine/optim.jl:6-171).

    `fdf_` is a device objective handle; `x_initial` a host vector (this rank's shard); the
    returned Results hold host copies of the minimiser and gradient.  Keyword knobs (not in the
    reference): `fuse_direction` defers updatedir! into the next trial kernel (bitwise identical
    result, 24n fewer bytes per iteration); `beta_form="literal"` evaluates the HZ / YWS β
    exactly as cg_flavours.jl:71-76 writes it (one extra pass) instead of the algebraically
    equal pack form; `quadratic_linesearch` (CSR least squares) evaluates the trials
    of each line search from v = A u — one SpMV and one SpMVᵀ per iteration whatever the number of
    trials; ϕ then comes from ½r·r + a r·v + ½a² v·v, so decisions agree with the plain path up to
    rounding.  `x_initial` may be a `DeviceStart` (the iterate of a live workspace: no upload); `keep_workspace`
    leaves the run's device state alive in `ret.workspace` (caller closes it) so that a restart can begin from it.
    """
    run = MinimizerRun(fdf_, x_initial, config, linesearch_config, fuse_direction, beta_form,
                       quadratic_linesearch)
    run.run()
    ret = run.results()
    ret.h2d_bytes = getattr(run.info, "h2d_bytes", 0)
    if keep_workspace:
        ret.workspace = run.info
    else:
        run.info.close()
    return ret


def minimizeobjectivererun(fdf_, x_initial, config: CGConfig, linesearch_config: LineSearchConfig,
                           *rerun_config_tuples: Tuple[CGConfig, LineSearchConfig], **kw) -> List[Results]:
    """minimizeobjectivererun (src/engine/optim.jl:173-208)."""
    # the next attempt starts from rets[end].minimizer (:197): that vector is still on the device, in the previous
    # attempt's workspace — start from it there (DeviceStart: one D2D copy) instead of uploading the host copy
    from ..device import DeviceStart
    keep = kw.pop("keep_workspace", False)
    live = []

    def attempt(x0, cfg, ls):
        ret = minimizeobjective(fdf_, x0, cfg, ls, keep_workspace=True, **kw)
        live.append(ret.workspace)
        for ws in live[:-1]:
            ws.close()
        del live[:-1]
        return ret

    try:
        rets = [attempt(x_initial, config, linesearch_config)]                        # :183-188
        for rerun_config, backup_linesearch_config in rerun_config_tuples:            # :191
            if rets[-1].status != "success":
                rets.append(attempt(DeviceStart(live[-1]), rerun_config, backup_linesearch_config))   # :195-200
            else:
                break                                                                 # :203
    except BaseException:
        for ws in live:
            ws.close()
        raise
    for r in rets[:-1]:
        r.workspace = None
    if keep:
        rets[-1].workspace = live[-1]
    else:
        live[-1].close()
        rets[-1].workspace = None
    return rets
