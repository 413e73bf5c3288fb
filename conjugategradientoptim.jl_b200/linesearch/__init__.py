"""Line-search plugins — host mirrors of src/linesearch/*.jl (scalar state machines only)."""
from .nocedal import StrongWolfeBisection, setupStrongWolfeBisection  # noqa: F401
from .wolfe import Wolfe, WolfeBisection, YuanWeiLuWolfe  # noqa: F401
from .geometric import Armijo, Backtracking  # noqa: F401
