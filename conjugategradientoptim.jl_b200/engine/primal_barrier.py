"""Primal barrier method — host mirror of src/engine/primal_barrier.jl (Alg. 11.1 of Boyd &
Vandenberghe 2004): an outer loop of centering steps, each one `minimizeobjectivererun` on
t·f0(x) + ψ(x), t growing geometrically.

The reference takes the constraints as a host callback `hdh!(fi_evals, dfi_evals, x)` with one dense
gradient vector per constraint (primal_barrier.jl:37-61).  A host callback cannot run on the GPU, and
M dense n-vectors do not exist at n = 1e8; what the device path offers is the constraint set the
reference's own example uses (examples/constrained.jl:17-47): the box lbs < x < ubs, whose 2n
constraint gradients are ±e_d.  `hdh!` is therefore a `BoxConstraint(lbs, ubs)` description, and the
barrier objective (evalbarrier!, :112-133) is the device objective `BoxBarrierGPU`.

Kept as written: every centering step starts from `x_initial` (the loop at :215-247 never updates
`x`), so the method only "continues" through t.  `update_iterate=True` (not in the reference) starts
each step from the previous centre, as Alg. 11.1 does.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List

import numpy as np

from ..cg_types import CGConfig, LineSearchConfig, Results
from .optim import minimizeobjectivererun

f64 = np.float64


@dataclass
class PrimalBarrierResults:
    """primal_barrier.jl:1-7"""
    centering_results: List[List[Results]]
    status: str
    iters_ran: int
    t_final: float
    total_objective_evals: int


def assembleresults_(rets, status: str, iter: int, t) -> PrimalBarrierResults:
    """assembleresults! (primal_barrier.jl:9-34)"""
    del rets[iter:]                                                 # resize!(rets, iter)  :16
    total_objective_evals = 0
    for step in rets:                                               # :18-25
        for r in step:
            total_objective_evals += int(np.sum(r.trace.objective_evals))
    return PrimalBarrierResults(rets, status, iter, t, total_objective_evals)


@dataclass
class CvxInequalityConstraint:
    """primal_barrier.jl:37-41.  The reference's buffers (fi_evals, one dense dfi vector per
    constraint, grad) live inside the barrier kernel here; only the sizes remain."""
    M: int
    D: int


def setupCvxInequalityConstraint(M: int, D: int) -> CvxInequalityConstraint:
    """primal_barrier.jl:51-58 (the element type argument is fixed to Float64)"""
    return CvxInequalityConstraint(int(M), int(D))


def getNconstraints(X: CvxInequalityConstraint) -> int:            # :60-62
    return X.M


@dataclass(frozen=True)
class BoxConstraint:
    """The `hdh!` of examples/constrained.jl:17-47: fi = [x − ubs; lbs − x] (this rank's shard)."""
    lbs: Any
    ubs: Any


@dataclass(frozen=True)
class PrimalBarrierConfig:
    """primal_barrier.jl:135-141"""
    barrier_tol: float
    barrier_growth_factor: float
    max_iters: int
    t_initial: float
    inf_f0_lb: float


def setupPrimalBarrierConfig(barrier_tol, barrier_growth_factor, max_iters: int, *,
                             t_initial=float("nan")) -> PrimalBarrierConfig:
    """primal_barrier.jl:143-155"""
    return PrimalBarrierConfig(float(barrier_tol), float(barrier_growth_factor), int(max_iters),
                               float(t_initial), 0.0)


def verifyt0(t0, x0, f0df0_, μ, inf_f0_lb):
    """verifyt0 (primal_barrier.jl:259-277): t0 = (f0(x0) − inf f0) μ when none is given"""
    if not np.isfinite(t0) or t0 < 0.0:
        ws = f0df0_.make_workspace(x0, fuse_direction=False)        # f_x0 = f0df0!(df_x0, x0)  :270
        f_x0 = ws.f_x0
        ws.close()
        return f64((f_x0 - inf_f0_lb) * μ)                          # :271
    return f64(t0)


def primalbarriermethod_(constraints: CvxInequalityConstraint, f0df0_, hdh_: BoxConstraint, x_initial,
                         centering_config: CGConfig, linesearch_config: LineSearchConfig,
                         barrier_config: PrimalBarrierConfig, *rerun_config_tuples,
                         update_iterate: bool = False, make_barrier=None, **kw) -> PrimalBarrierResults:
    """primalbarriermethod! (src/engine/primal_barrier.jl:158-255).

    `f0df0_` is a device objective, `hdh_` a BoxConstraint, `constraints` the size descriptor
    (`setupCvxInequalityConstraint(2*D, D)` for a box).  `make_barrier(f0df0_, lbs, ubs, t)` builds
    the barrier objective (default: BoxBarrierGPU; tests inject the numpy stand-in)."""
    if make_barrier is None:
        from ..device import BoxBarrierGPU
        make_barrier = BoxBarrierGPU
    assert isinstance(hdh_, BoxConstraint), "the device path implements box constraints (examples/constrained.jl:17-47)"
    barrier_tol = barrier_config.barrier_tol                       # :170-174
    barrier_growth_factor = barrier_config.barrier_growth_factor
    max_iters = barrier_config.max_iters
    t_initial = barrier_config.t_initial
    inf_f0_lb = barrier_config.inf_f0_lb
    N_constraints = getNconstraints(constraints)                   # :176
    x = np.array(x_initial, dtype=np.float64)                      # :179
    rets: List[List[Results]] = [None] * max_iters                 # :184
    live = []                                                      # device workspaces kept for DeviceStart

    fdf_ = make_barrier(f0df0_, hdh_.lbs, hdh_.ubs, 1.0)           # (t is set below)  :205-213
    try:
        # check if the initial iterate is feasible.                 :187-198
        if fdf_.infeasible_count(x) > 0:
            return assembleresults_(rets, "infeasible_start", 0, t_initial)
        t = verifyt0(t_initial, x, f0df0_, barrier_growth_factor, inf_f0_lb)   # :200
        # every centering step starts from a vector that is already on the device: x_initial (the reference
        # restarts from it every time, :217) is uploaded ONCE into a workspace that is never stepped, the minimiser
        # of the previous step (update_iterate) stays in that step's workspace — DeviceStart, no H2D per step
        device_start = hasattr(fdf_, "make_workspace") and hasattr(fdf_, "h")
        start = x
        if device_start:
            from ..device import DeviceStart
            fdf_.set_t(t)
            anchor = fdf_.make_workspace(x, fuse_direction=False)
            live.append(anchor)
            start = DeviceStart(anchor)
        for i in range(1, max_iters + 1):                           # :215
            fdf_.set_t(t)
            rets[i - 1] = minimizeobjectivererun(fdf_, start, centering_config, linesearch_config,   # :217-223
                                                 *rerun_config_tuples, keep_workspace=device_start and update_iterate, **kw)
            if rets[i - 1][-1].status != "success":                 # :224-232
                return assembleresults_(rets, "centering_step_issue", i, t)
            if N_constraints / t < barrier_tol:                     # :235-243
                return assembleresults_(rets, "success", i, t)
            if update_iterate:
                if device_start:
                    ws = rets[i - 1][-1].workspace
                    rets[i - 1][-1].workspace = None
                    for old in live:
                        old.close()
                    live[:] = [ws]
                    start = DeviceStart(ws)
                else:
                    start = np.array(rets[i - 1][-1].minimizer, dtype=np.float64)
            t = barrier_growth_factor * t                           # :246
        return assembleresults_(rets, "max_iters_reached", max_iters, t)   # :249-254
    finally:
        for ws in live:
            ws.close()
        for rr in rets:
            if rr:
                for r_ in rr:
                    w_ = getattr(r_, "workspace", None)
                    if w_ is not None:
                        w_.close()
                        r_.workspace = None
        fdf_.close()
