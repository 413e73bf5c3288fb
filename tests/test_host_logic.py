"""CPU tests of the host mirror (engine, three line searches, five flavours, rerun wrapper)
against the C oracle, through the test-only numpy workspace.  Both sides use the canonical
reduction order, so whole runs must agree bit for bit."""
import numpy as np
import pytest

import cgoptim_b200 as cg
from oracle import oracle as O

from helpers import FLAVOURS, LINESEARCHES, assert_same_run, make_pair
from numpy_workspace import NumpyObjective


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("linesearch", LINESEARCHES)
@pytest.mark.parametrize("beta_form", ["fused", "literal"])
def test_booth_all_flavours_all_linesearches(flavour, linesearch, beta_form):
    """Booth (examples/helpers/test_funcs.jl:3-12) from x0 of examples/min.jl:38."""
    if beta_form == "literal" and flavour not in ("HagerZhang", "YuanWangSheng"):
        pytest.skip("literal form only differs for the HZ family")
    ocfg, cfg, ls = make_pair(flavour, linesearch, beta_form=beta_form)
    x0 = np.array([0.43, 1.23])
    ora = O.minimize(O.Objective.booth(), x0, ocfg)
    ret = cg.minimizeobjective(NumpyObjective(O.Objective.booth()), x0, cfg, ls, beta_form=beta_form)
    assert_same_run(ret, ora, what=f"{flavour}/{linesearch}/{beta_form}")


@pytest.mark.parametrize("flavour", FLAVOURS)
@pytest.mark.parametrize("linesearch", LINESEARCHES)
def test_rosenbrock_small(flavour, linesearch):
    n = 10
    ocfg, cfg, ls = make_pair(flavour, linesearch, max_iters=300)
    x0 = O.rosenbrock_x0(n, 24, 0.1)
    ora = O.minimize(O.Objective.rosenbrock(n), x0, ocfg)
    ret = cg.minimizeobjective(NumpyObjective(O.Objective.rosenbrock(n)), x0, cfg, ls)
    assert_same_run(ret, ora, what=f"{flavour}/{linesearch}")


@pytest.mark.parametrize("linesearch", LINESEARCHES)
@pytest.mark.parametrize("flavour", ["HagerZhang", "LBFGS"])
def test_barrier_nonfinite_paths(flavour, linesearch):
    """Objective that is non-finite outside |x_i| < 1: exercises findfeasiblestepsize!
    (wolfe.jl:171-207), the non-finite exits (optim.jl:108-121) and failure statuses."""
    n = 6
    ocfg, cfg, ls = make_pair(flavour, linesearch, max_iters=200)
    x0 = np.linspace(-0.5, 0.5, n)
    ora = O.minimize(O.Objective.barrier(n), x0, ocfg)
    ret = cg.minimizeobjective(NumpyObjective(O.Objective.barrier(n)), x0, cfg, ls)
    assert_same_run(ret, ora, what=f"{flavour}/{linesearch}")


def test_sparse_ls_small_bit_exact():
    n = 3000
    ocfg, cfg, ls = make_pair(max_iters=100)
    x0 = np.zeros(n)
    ora = O.minimize(O.Objective.sparse_ls(n, 10, 64, 24, 0), x0, ocfg)
    ret = cg.minimizeobjective(NumpyObjective(O.Objective.sparse_ls(n, 10, 64, 24, 0)), x0, cfg, ls)
    assert ora.status == "success"
    assert_same_run(ret, ora)


def test_rerun_wrapper():
    """minimizeobjectivererun (optim.jl:173-208): primary config capped at 3 iterations fails
    with max_iters_reached, the backups restart from the last minimiser."""
    x0 = np.array([0.43, 1.23])
    o1, c1, l1 = make_pair("HagerZhang", "StrongWolfeBisection", max_iters=3)
    o2, c2, l2 = make_pair("LiuStorrey", "Wolfe", max_iters=4)
    o3, c3, l3 = make_pair("SallehAlhawarat", "StrongWolfeBisection", max_iters=500)
    o4, c4, l4 = make_pair("YuanWangSheng", "StrongWolfeBisection", max_iters=500)
    oras = O.minimize_rerun(O.Objective.booth(), x0, [o1, o2, o3, o4])
    rets = cg.minimizeobjectivererun(NumpyObjective(O.Objective.booth()), x0, c1, l1, (c2, l2), (c3, l3), (c4, l4))
    assert [r.status for r in rets] == ["max_iters_reached", "max_iters_reached", "success"]
    assert len(rets) == len(oras) == 3
    for r, o in zip(rets, oras):
        assert_same_run(r, o)


def test_config_asserts():
    with pytest.raises(AssertionError):
        cg.setupCGConfig(1.5, cg.HagerZhang(), cg.EnableTrace())          # types.jl:187
    with pytest.raises(AssertionError):
        cg.setupStrongWolfeBisection(0.9, 0.1)                            # nocedal.jl:22
    with pytest.raises(AssertionError):
        cg.setupStrongWolfeBisection(1e-4, 0.9, a_max_growth_factor=1.0)  # nocedal.jl:26


def test_host_callback_rejected():
    """No CPU fallback: a Python callable is not an objective."""
    cfg = cg.setupCGConfig(1e-5, cg.HagerZhang(), cg.EnableTrace())
    ls = cg.setupStrongWolfeBisection(1e-5, 0.8)
    with pytest.raises(TypeError):
        cg.minimizeobjective(lambda g, x: 0.0, np.zeros(2), cfg, ls)


def test_disable_trace():
    ocfg, cfg, ls = make_pair(trace=False)
    x0 = np.array([0.43, 1.23])
    ret = cg.minimizeobjective(NumpyObjective(O.Objective.booth()), x0, cfg, ls)
    assert ret.status == "success" and len(ret.trace.objective) == 0


def test_broyden_family_is_what_the_reference_computes():
    """src/qn_flavours.jl:72-90 restated with the dense matrix: with `s = B\\y` the update leaves B where it is
    (B starts as the identity, :31-34), so `u = B\\(-df_x)` (:17) is −df_x — which is what the kept
    `BroydenFamily` flavour does without the n×n matrix (β = 0)."""
    rng = np.random.default_rng(5)
    n = 12
    for θ in (0.0, 0.3, 1.0, 2.5):
        B = np.eye(n)
        for _ in range(20):                                   # getβ, qn_flavours.jl:78-88
            g, g_next = rng.standard_normal(n), rng.standard_normal(n)
            y = g_next - g
            s = np.linalg.solve(B, y)
            Bs = B @ s
            sBs = s @ Bs
            v = y / (s @ y) - Bs / sBs
            B = B - np.outer(Bs, Bs) / sBs + np.outer(y, y) / (s @ y) + θ * sBs * np.outer(v, v)
            assert np.abs(B - np.eye(n)).max() < 1e-12
            assert np.allclose(np.linalg.solve(B, -g_next), -g_next, rtol=0, atol=1e-11)   # updatedir!, :17
    cfgθ = cg.setupBroydenFamily(2.5, 2)                      # only `0 <= θ` is asserted (:57)
    with pytest.raises(AssertionError):
        cg.setupBroydenFamily(-0.1, 2)
    # the flavour itself: every direction is −df_x, i.e. the run equals a CG run whose β is always 0
    from cgoptim_b200 import cg_flavours
    x0 = np.array([0.43, 1.23])
    _, _, ls = make_pair("HagerZhang")
    cfg = cg.setupCGConfig(1e-5, cfgθ, cg.EnableTrace(), max_iters=400)
    run = cg.MinimizerRun(NumpyObjective(O.Objective.booth()), x0, cfg, ls)
    for _ in range(5):
        assert run.step() is None
        assert run.β == 0.0 or float(run.β) == 0.0
        run.info.dot_g_u()                                   # materialises the deferred updatedir!
        assert np.array_equal(run.info.u_, -run.info.g_)
    ret = cg.minimizeobjective(NumpyObjective(O.Objective.booth()), x0, cfg, ls)
    assert ret.status == "success" and np.allclose(ret.minimizer, [1.0, 3.0], atol=1e-4)
