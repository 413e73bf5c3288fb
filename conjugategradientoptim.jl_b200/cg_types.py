"""Config / result / trace types — host mirror of src/types.jl (citations: file:line there).

Julia's multiple dispatch on config types becomes isinstance dispatch; Symbols become str.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List

import numpy as np


class LineSearchConfig:          # types.jl:1
    pass


class βConfig:                   # types.jl:4
    pass


class CGβConfig(βConfig):        # types.jl:5
    pass


class QNβConfig(βConfig):        # types.jl:6
    pass


class TraceTrait:                # types.jl:9
    pass


class EnableTrace(TraceTrait):   # types.jl:10
    pass


class DisableTrace(TraceTrait):  # types.jl:11
    pass


@dataclass
class TraceContainer:
    """types.jl:17-23 — only stores a full, successful cg update."""
    objective: np.ndarray
    grad_norm: np.ndarray
    step_size: np.ndarray
    objective_evals: np.ndarray
    status: TraceTrait


def setuptrace(status: TraceTrait) -> TraceContainer:                     # types.jl:25-33
    return TraceContainer(np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.int64), status)


def resizetrace_(t: TraceContainer, N: int) -> None:                      # types.jl:35-54
    if isinstance(t.status, EnableTrace):
        t.objective = np.resize(t.objective, N) if N else t.objective[:0]
        t.grad_norm = np.resize(t.grad_norm, N) if N else t.grad_norm[:0]
        t.step_size = np.resize(t.step_size, N) if N else t.step_size[:0]
        t.objective_evals = np.resize(t.objective_evals, N) if N else t.objective_evals[:0]


def updatetrace_(t: TraceContainer, f_x, df_x_norm, step_size, objective_evals: int, n: int) -> None:
    """types.jl:56-79 (n is 1-based like the reference)."""
    if isinstance(t.status, EnableTrace):
        t.objective[n - 1] = f_x
        t.grad_norm[n - 1] = df_x_norm
        t.step_size[n - 1] = step_size
        t.objective_evals[n - 1] = objective_evals


@dataclass
class Results:
    """types.jl:107-114.  minimizer / gradient are host vectors filled by one D2H at exit
    (this rank's shard when the objective is sharded)."""
    objective: float
    minimizer: Any
    gradient: Any
    iters_ran: int
    status: str
    trace: TraceContainer
    # not in the reference: host → device bytes the run's start cost (0 for a DeviceStart), and — on request
    # (keep_workspace) — the run's live device state, from which a restart can begin without a host round trip
    h2d_bytes: int = 0
    workspace: Any = None


def updateresult_(ret: Results, x, df_x, f_x, i: int, status: str) -> None:
    """types.jl:134-151.  In the reference ret.minimizer / ret.gradient alias the engine's x /
    df_x, so the copies are self-copies; here x / df_x are the device vectors and the host
    copies are taken when the engine returns."""
    ret.objective = f_x
    ret.minimizer = x
    ret.gradient = df_x
    ret.iters_ran = i
    ret.status = status
    resizetrace_(ret.trace, i)


@dataclass(frozen=True)
class CGConfig:
    """types.jl:156-168"""
    ϵ: float
    β_config: βConfig
    max_iters: int
    verbose: bool
    trace_status: TraceTrait

    @property
    def eps(self):
        return self.ϵ


def setupCGConfig(ϵ: float, β_config: βConfig, trace_status: TraceTrait, *, max_iters: int = 1000,
                  verbose: bool = False) -> CGConfig:
    """types.jl:171-203"""
    assert isinstance(β_config, βConfig) and isinstance(trace_status, TraceTrait)
    assert 0.0 < ϵ < 1.0                                                  # types.jl:187
    return CGConfig(float(ϵ), β_config, int(max_iters), bool(verbose), trace_status)
