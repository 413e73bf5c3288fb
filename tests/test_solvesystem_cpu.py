"""CPU tests of the solvesystem host mirror (conjugategradientoptim.jl_b200/engine/solve_system.py)
against the C oracle's restatement of src/engine/solve_system.jl, through the test-only numpy
workspace: same canonical reduction order on both sides, so whole runs agree bit for bit — in the
as-written variant (stale x_next, solve_system.jl:171-177) and in Alg. 3.1 as published."""
import numpy as np
import pytest

import cgoptim_b200 as cg
from oracle import oracle as O

from helpers import assert_same_run, make_pair
from numpy_workspace import NumpyObjective

CG_FLAVOURS = ["HagerZhang", "YuanWangSheng", "SallehAlhawarat", "LiuStorrey"]


def _both(make_obj, x0, flavour, fix, max_iters=300, s=1.0, sigma=0.5, rho=0.95, ls_iters=None, eps=1e-5):
    ocfg, cfg, _ = make_pair(flavour, eps=eps, max_iters=max_iters)
    ora = O.solvesystem(make_obj(), x0, ocfg, O.solvesys_ls(s, sigma, rho, ls_iters, fix_stale_iterate=fix))
    ls = cg.setupLinesearchSolveSys(s, σ=sigma, ρ=rho, max_iters=ls_iters)
    ret = cg.solvesystem(NumpyObjective(make_obj()), x0, cfg, ls, fix_stale_iterate=fix)
    return ret, ora


def test_setup_defaults():
    ls = cg.setupLinesearchSolveSys(1.0)
    assert (ls.ρ, ls.σ, ls.s) == (0.95, 0.5, 1.0) and ls.max_iters == 269      # round(log(0.95, 1e-6))
    with pytest.raises(AssertionError):
        cg.setupLinesearchSolveSys(1.0, ρ=1.5)
    with pytest.raises(AssertionError):
        cg.setupLinesearchSolveSys(-1.0)


@pytest.mark.parametrize("flavour", CG_FLAVOURS)
@pytest.mark.parametrize("fix", [False, True])
def test_booth_dev_example(flavour, fix):
    """dev/solve_sys.jl:26-59: Booth, x0 = [0.43, 1.23], s = 1, σ = 0.5, ρ = 0.95, ϵ = 1e-5."""
    ret, ora = _both(O.Objective.booth, np.array([0.43, 1.23]), flavour, fix, max_iters=120)
    assert_same_run(ret, ora, what=f"booth/{flavour}/fix={fix}")
    if fix and flavour == "YuanWangSheng":            # the method of the paper, as published, converges
        assert ret.status == "success" and np.allclose(ret.minimizer, [1.0, 3.0], atol=1e-4)
    if not fix:                                       # as written it does not (stale iterate)
        assert ret.status != "success"


@pytest.mark.parametrize("flavour", ["HagerZhang", "YuanWangSheng"])
@pytest.mark.parametrize("fix", [False, True])
def test_sparse_ls_normal_equations(flavour, fix):
    """g(x) = Aᵀ(Ax − b) = 0: a monotone (SPD linear) system, the class Alg. 3.1 is stated for."""
    n = 400
    mk = lambda: O.Objective.sparse_ls(n, 10, 64, 24, 0)
    ret, ora = _both(mk, np.zeros(n), flavour, fix, max_iters=90 if fix else 25)
    assert_same_run(ret, ora, what=f"sparse_ls/{flavour}/fix={fix}")
    if fix and flavour == "YuanWangSheng":
        assert ret.status == "success"


def test_rosenbrock_and_status_paths():
    n = 10
    x0 = O.rosenbrock_x0(n, 24, 0.1)
    ret, ora = _both(lambda: O.Objective.rosenbrock(n), x0, "HagerZhang", True, max_iters=40)
    assert_same_run(ret, ora)
    ret, ora = _both(lambda: O.Objective.rosenbrock(n), x0, "HagerZhang", True, max_iters=40, ls_iters=3)
    assert_same_run(ret, ora)
    assert ret.status == "linesearch_failed"
    ret, ora = _both(O.Objective.booth, np.array([1.0, 3.0]), "HagerZhang", False)   # starts at the root
    assert_same_run(ret, ora)
    assert ret.status == "success" and ret.iters_ran == 0
    ret, ora = _both(lambda: O.Objective.barrier(6), np.linspace(-0.5, 0.5, 6), "HagerZhang", True, s=4.0)
    assert_same_run(ret, ora)


def test_lbfgs_rejected():
    _, cfg, _ = make_pair("LBFGS")
    with pytest.raises(AssertionError):
        cg.solvesystem(NumpyObjective(O.Objective.booth()), np.zeros(2), cfg, cg.setupLinesearchSolveSys(1.0))
