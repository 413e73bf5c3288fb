"""Engines — host mirror of src/engine/optim.jl and src/engine/solve_system.jl."""
from .optim import MinimizerRun, minimizeobjective, minimizeobjectivererun  # noqa: F401
from .solve_system import LinesearchSolveSys, setupLinesearchSolveSys, solvesystem  # noqa: F401
