"""GPU parity tests of solvesystem (src/engine/solve_system.jl) on the device path: the line search is
the fused trial kernel, updateiteratesolvesys! + the re-evaluation at x_next are
cgo_solvesys_project, the swaps are pointer swaps.  Whole runs are compared bit for bit with the
oracle's restatement, as written (stale x_next) and as published (fix_stale_iterate)."""
import numpy as np
import pytest

import cgoptim_b200 as cg
from oracle import oracle as O

from helpers import assert_same_run, make_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cg.Context(0)
    yield c
    c.close()


def _run(obj, ora_obj, x0, flavour, fix, max_iters, ls_iters=None):
    ocfg, cfg, _ = make_pair(flavour, max_iters=max_iters)
    ora = O.solvesystem(ora_obj, x0, ocfg, O.solvesys_ls(1.0, max_iters=ls_iters, fix_stale_iterate=fix))
    ls = cg.setupLinesearchSolveSys(1.0, max_iters=ls_iters)
    ret = cg.solvesystem(obj, x0, cfg, ls, fix_stale_iterate=fix)
    return ret, ora


@pytest.mark.parametrize("flavour", ["HagerZhang", "YuanWangSheng", "SallehAlhawarat", "LiuStorrey"])
@pytest.mark.parametrize("fix", [False, True])
def test_rosenbrock_bit_exact(ctx, flavour, fix):
    n = 4096
    x0 = O.rosenbrock_x0(n, 24, 0.1)
    obj = cg.RosenbrockGPU(n, ctx)
    ret, ora = _run(obj, O.Objective.rosenbrock(n), x0, flavour, fix, 30)
    assert_same_run(ret, ora, what=f"rosenbrock/{flavour}/fix={fix}")
    assert ret.iters_ran >= 1
    obj.close()


@pytest.mark.parametrize("flavour", ["HagerZhang", "YuanWangSheng"])
@pytest.mark.parametrize("fix", [False, True])
def test_sparse_ls_normal_equations_bit_exact(ctx, flavour, fix):
    """g(x) = Aᵀ(Ax − b) = 0 through the CSR kernels."""
    n = 20_000
    obj = cg.SparseLSGPU(n, 10, 256, 24, 0, ctx)
    ret, ora = _run(obj, O.Objective.sparse_ls(n, 10, 256, 24, 0), np.zeros(n), flavour, fix, 60 if fix else 12)
    assert_same_run(ret, ora, what=f"sparse_ls/{flavour}/fix={fix}")
    if fix and flavour == "YuanWangSheng":
        assert ret.trace.grad_norm[-1] < 1e-2 * ret.trace.grad_norm[0]
    obj.close()


def test_early_exit_returns_linesearch_point_and_failure_status(ctx):
    n = 64
    x0 = O.rosenbrock_x0(n, 24, 0.0)
    obj = cg.RosenbrockGPU(n, ctx)
    ret, ora = _run(obj, O.Objective.rosenbrock(n), x0, "HagerZhang", True, 50, ls_iters=2)
    assert_same_run(ret, ora)
    assert ret.status == "linesearch_failed"
    ones = np.ones(n)                                   # start at the root: zero iterations
    ret, ora = _run(obj, O.Objective.rosenbrock(n), ones, "HagerZhang", False, 50)
    assert_same_run(ret, ora)
    assert ret.status == "success" and ret.iters_ran == 0
    obj.close()


def test_large_run_reproducible(ctx):
    n = 2_000_000
    obj = cg.SparseLSGPU(n, 10, None, 24, 30, ctx)
    _, cfg, _ = make_pair("YuanWangSheng", max_iters=6)
    ls = cg.setupLinesearchSolveSys(1.0)
    a = cg.solvesystem(obj, np.zeros(n), cfg, ls, fix_stale_iterate=True)
    b = cg.solvesystem(obj, np.zeros(n), cfg, ls, fix_stale_iterate=True)
    assert np.array_equal(a.trace.objective, b.trace.objective) and np.array_equal(a.minimizer, b.minimizer)
    assert a.trace.grad_norm[-1] < a.trace.grad_norm[0]
    obj.close()
