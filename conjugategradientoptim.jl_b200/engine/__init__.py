"""Engines — host mirror of src/engine/optim.jl, solve_system.jl and primal_barrier.jl."""
from .optim import MinimizerRun, minimizeobjective, minimizeobjectivererun  # noqa: F401
from .solve_system import LinesearchSolveSys, setupLinesearchSolveSys, solvesystem  # noqa: F401
from .primal_barrier import (BoxConstraint, CvxInequalityConstraint, PrimalBarrierConfig,  # noqa: F401
                             PrimalBarrierResults, getNconstraints, primalbarriermethod_,
                             setupCvxInequalityConstraint, setupPrimalBarrierConfig, verifyt0)
