"""Engines — host mirror of src/engine/optim.jl."""
from .optim import MinimizerRun, minimizeobjective, minimizeobjectivererun  # noqa: F401
