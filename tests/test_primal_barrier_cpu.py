"""CPU tests of the primal-barrier host mirror (engine/primal_barrier.py) against the oracle's
restatement of src/engine/primal_barrier.jl (oracle.primalbarrier + the C box-barrier objective),
through the numpy stand-in: identical reduction order and the same libm, so bit for bit."""
import numpy as np
import pytest

import cgoptim_b200 as cg
from oracle import oracle as O

from helpers import make_pair
from numpy_workspace import NumpyBoxBarrier, NumpyObjective


def _same(b, ob):
    assert b.status == ob.status and b.iters_ran == ob.iters_ran
    assert b.total_objective_evals == ob.total_objective_evals
    assert (b.t_final == ob.t_final) or (np.isnan(b.t_final) and np.isnan(ob.t_final))
    assert len(b.centering_results) == len(ob.centering_results)
    for step, ostep in zip(b.centering_results, ob.centering_results):
        assert len(step) == len(ostep)
        for r, o in zip(step, ostep):
            assert r.status == o.status and r.iters_ran == o.iters_ran
            assert np.array_equal(r.trace.objective, o.trace_objective)
            assert np.array_equal(r.trace.step_size, o.trace_step_size)
            assert np.array_equal(r.minimizer, o.minimizer, equal_nan=True)
            assert np.array_equal(r.gradient, o.gradient, equal_nan=True)


def _run(make_f0, lbs, ubs, x0, pairs, tol=1e-8, growth=10.0, iters=100, t0=float("nan"), update=False):
    ocfgs = [p[0] for p in pairs]
    ob = O.primalbarrier(make_f0(), lbs, ubs, x0, ocfgs, tol, growth, iters, t0, update_iterate=update)
    D = len(x0)
    b = cg.primalbarriermethod_(cg.setupCvxInequalityConstraint(2 * D, D), NumpyObjective(make_f0()),
                                cg.BoxConstraint(np.asarray(lbs, float), np.asarray(ubs, float)), np.asarray(x0, float),
                                pairs[0][1], pairs[0][2], cg.setupPrimalBarrierConfig(tol, growth, iters, t_initial=t0),
                                *[(p[1], p[2]) for p in pairs[1:]], update_iterate=update, make_barrier=NumpyBoxBarrier)
    _same(b, ob)
    return b


def test_barrier_objective_value_and_gradient():
    """t·f0 + ψ and t·∇f0 + dψ against the formulas of primal_barrier.jl:78, :83, :130-132."""
    n = 6
    x = np.linspace(-0.8, 0.9, n)
    lbs, ubs = -np.ones(n) * 2, np.ones(n) * 1.5
    f0 = O.Objective.rosenbrock(n)
    bar = O.Objective.box_barrier(f0, lbs, ubs, 3.0)
    f, g = bar.fdf(x)
    f0v, g0 = f0.fdf(x)
    psi = -(np.sum(np.log(ubs - x)) + np.sum(np.log(x - lbs)))
    assert abs(f - (3.0 * f0v + psi)) <= 1e-13 * abs(f)
    np.testing.assert_allclose(g, 3.0 * g0 + 1.0 / (ubs - x) - 1.0 / (x - lbs), rtol=1e-14)
    fo, go = bar.fdf(np.array([1.5, 0, 0, 0, 0, 0.0]))                # on the boundary: log(0)
    assert fo == np.inf and not np.isfinite(go[0])


@pytest.mark.parametrize("update", [False, True])
def test_booth_example(update):
    """examples/constrained.jl:10-199: Booth in the box [−10, 10]², x0 = [0.43, 1.23], HZ + Wolfe(1e-3, 0.9)
    with YuanWeiLuWolfe / Armijo backups, barrier_tol 1e-8, growth 10."""
    pairs = [make_pair("HagerZhang", "Wolfe", sum_mode="cgo"),
             make_pair("LiuStorrey", "Backtracking", sum_mode="cgo"),
             make_pair("SallehAlhawarat", "YuanWeiLuWolfe", sum_mode="cgo")]
    b = _run(O.Objective.booth, [-10.0, -10.0], [10.0, 10.0], [0.43, 1.23], pairs, update=update)
    assert b.iters_ran >= 1 and b.total_objective_evals > 0


@pytest.mark.parametrize("flavour,ls", [("HagerZhang", "StrongWolfeBisection"), ("LBFGS", "Wolfe")])
def test_rosenbrock_active_bound(flavour, ls):
    """the unconstrained minimiser ones(n) lies outside the box: the barrier path ends on the bound"""
    n = 8
    pairs = [make_pair(flavour, ls, max_iters=400, eps=1e-4)]
    b = _run(lambda: O.Objective.rosenbrock(n), -2.0 * np.ones(n), 0.8 * np.ones(n), np.zeros(n), pairs,
             tol=1e-3, growth=20.0, iters=12, t0=1.0, update=True)
    assert b.status in ("success", "centering_step_issue", "max_iters_reached")
    if b.status == "success":
        x = b.centering_results[-1][-1].minimizer
        assert np.all(x < 0.8) and np.max(x) > 0.75


def test_infeasible_start_and_config():
    pairs = [make_pair("HagerZhang", "Wolfe")]
    b = _run(O.Objective.booth, [-10.0, -10.0], [0.4, 10.0], [0.43, 1.23], pairs)
    assert b.status == "infeasible_start" and b.iters_ran == 0 and b.centering_results == []
    cfg = cg.setupPrimalBarrierConfig(1e-8, 10.0, 100)
    assert np.isnan(cfg.t_initial) and cfg.inf_f0_lb == 0.0
    assert cg.getNconstraints(cg.setupCvxInequalityConstraint(4, 2)) == 4
