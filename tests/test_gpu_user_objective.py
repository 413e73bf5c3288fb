"""The user-objective plug-in (include/cgoptim.h cgo_obj_user_create; UserObjectiveGPU): ANY fdf!(g, x) -> f, the
reference's own callback signature (src/engine/optim.jl:6-11), evaluated on the device by the caller's code while
the line search, β and every dot stay the library's."""
import numpy as np
import pytest

import cgoptim_b200 as cg
from oracle import oracle as O

from helpers import make_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cg.Context(0)
    yield c
    c.close()


def test_booth_through_a_user_callback(ctx):
    """boothfdf! (examples/helpers/test_funcs.jl:3-12) written by the 'user' in torch, run with the canonical
    configuration of examples/min.jl:16-43: converges to the known minimiser [1, 3] (test/runtests.jl:18-21)."""
    import torch

    def boothfdf(g, x):                                    # fdf!(g, x) -> f
        a, b = x[0] + 2 * x[1] - 7, 2 * x[0] + x[1] - 5
        g[0] = 2 * a + 4 * b
        g[1] = 4 * a + 2 * b
        return a * a + b * b

    obj = cg.UserObjectiveGPU(2, boothfdf, ctx)
    _, cfg, ls = make_pair("HagerZhang", "StrongWolfeBisection", eps=1e-5, max_iters=1000)
    ret = cg.minimizeobjective(obj, np.array([0.43, 1.23]), cfg, ls)
    assert ret.status == "success" and np.allclose(ret.minimizer, [1.0, 3.0], atol=1e-5) and ret.objective < 1e-9
    ora = O.minimize(O.Objective.booth(), np.array([0.43, 1.23]), make_pair("HagerZhang")[0])
    assert ret.iters_ran == ora.iters_ran and np.array_equal(ret.trace.step_size, ora.trace_step_size)
    np.testing.assert_allclose(ret.trace.objective, ora.trace_objective, rtol=1e-9, atol=1e-300)
    obj.close()


@pytest.mark.parametrize("flavour,linesearch", [("HagerZhang", "StrongWolfeBisection"), ("LBFGS", "Wolfe"),
                                                 ("LiuStorrey", "Backtracking"), ("BroydenFamily", "StrongWolfeBisection")])
def test_user_rosenbrock_matches_the_builtin(ctx, flavour, linesearch):
    """extended Rosenbrock written by the user in torch against the built-in device objective: the gradient is
    elementwise (bit-identical), f differs only by torch's order of summation, so the runs take the same decisions."""
    import torch
    n = 20_000

    def fdf(g, x):
        x1, x2 = x[0::2], x[1::2]
        t, om = x2 - x1 * x1, 1.0 - x1
        g[0::2] = (-400.0 * x1) * t - 2.0 * om
        g[1::2] = 200.0 * t
        return torch.sum((100.0 * t) * t + om * om)

    user, ref = cg.UserObjectiveGPU(n, fdf, ctx), cg.RosenbrockGPU(n, ctx)
    x0 = ref.default_x0(24, 0.1)
    wu, wr = user.make_workspace(x0, fuse_direction=False), ref.make_workspace(x0, fuse_direction=False)
    assert np.array_equal(wu.download()[1], wr.download()[1]) and abs(wu.f_x0 - wr.f_x0) <= 1e-13 * wr.f_x0
    wu.close(); wr.close()
    if flavour == "BroydenFamily":
        cfg = cg.setupCGConfig(1e-5, cg.setupBroydenFamily(0.5, n), cg.EnableTrace(), max_iters=25)
        ls = make_pair("HagerZhang", linesearch)[2]
        cfg_sd = cfg
    else:
        _, cfg, ls = make_pair(flavour, linesearch, max_iters=25)
    a, b = cg.minimizeobjective(user, x0, cfg, ls), cg.minimizeobjective(ref, x0, cfg, ls)
    k = min(len(a.trace.objective), len(b.trace.objective), 12)
    assert k >= 5 and np.array_equal(a.trace.step_size[:k], b.trace.step_size[:k])
    np.testing.assert_allclose(a.trace.objective[:k], b.trace.objective[:k], rtol=1e-9)
    if flavour == "BroydenFamily":                          # the reference's quasi-Newton slot is steepest descent
        u = user.make_workspace(x0)
        u.reset_direction()
        assert np.array_equal(u.download_vector("u"), -u.download()[1])
        u.close()
    user.close(); ref.close()


def test_callback_errors_do_not_cross_the_abi(ctx):
    def bad(g, x):
        raise RuntimeError("user bug")
    obj = cg.UserObjectiveGPU(4, bad, ctx)
    with pytest.raises(cg.CgoError, match="user objective callback returned 1"):
        obj.make_workspace(np.zeros(4))
    assert isinstance(obj.last_error, RuntimeError)
    obj.close()
    with pytest.raises(cg.CgoError):
        cg.UserObjectiveGPU(3, bad, ctx)                    # odd n
