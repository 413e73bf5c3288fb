"""GPU parity tests of the box-constraint log barrier (BoxBarrierGPU: evalbarrier!,
src/engine/primal_barrier.jl:112-133) and of primalbarriermethod_ on the device path.  CUDA's log
and glibc's may differ in the last place, so — as for logistic regression — values are compared at
1e-13 per evaluation and whole runs at north_star's tolerances with identical decisions."""
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


@pytest.mark.parametrize("inner", ["rosenbrock", "sparse_ls"])
def test_barrier_value_gradient_and_pack(ctx, inner):
    n = 5000
    if inner == "rosenbrock":
        f0, of0 = cg.RosenbrockGPU(n, ctx), O.Objective.rosenbrock(n)
    else:
        f0, of0 = cg.SparseLSGPU(n, 10, 128, 24, 0, ctx), O.Objective.sparse_ls(n, 10, 128, 24, 0)
    rng = np.random.default_rng(4)
    lbs, ubs = -2.0 - rng.random(n), 1.5 + rng.random(n)
    x = -1.5 + 2.5 * rng.random(n)
    bar, obar = cg.BoxBarrierGPU(f0, lbs, ubs, 7.0), O.Objective.box_barrier(of0, lbs, ubs, 7.0)
    obar.set_sum_mode("cgo")
    ws = bar.make_workspace(x, fuse_direction=False)
    f, g = obar.fdf(x)
    assert abs(ws.f_x0 - f) <= 1e-13 * abs(f)
    gd = ws.download()[1]
    assert np.max(np.abs(gd - g)) <= 1e-13 * np.max(np.abs(g))
    ws.reset_direction()
    phi, dphi = ws.eval_trial(1e-4)
    xp = x + 1e-4 * (-gd)
    fp, gp = obar.fdf(xp)
    assert abs(phi - fp) <= 1e-12 * abs(fp) and abs(dphi - gp @ (-gd)) <= 1e-10 * abs(dphi)
    ws.close()
    bar.set_t(0.5)
    obar.set_t(0.5)
    ws = bar.make_workspace(x, fuse_direction=False)
    assert abs(ws.f_x0 - obar.fdf(x)[0]) <= 1e-13 * abs(ws.f_x0)
    ws.close()
    # outside the box: +Inf objective, non-finite gradient norm
    xo = x.copy(); xo[17] = ubs[17] + 0.1
    assert bar.infeasible_count(xo) == 1 and bar.infeasible_count(x) == 0
    ws = bar.make_workspace(xo, fuse_direction=False)
    assert ws.f_x0 == np.inf and not np.isfinite(ws.norm_df_x0)
    ws.close()
    bar.close(); f0.close()


@pytest.mark.parametrize("flavour,ls", [("HagerZhang", "StrongWolfeBisection"), ("LBFGS", "Wolfe"),
                                        ("HagerZhang", "Backtracking")])
def test_centering_run_matches_oracle(ctx, flavour, ls):
    """one centering step = minimizeobjective on the barrier objective (the line searches must back
    off from the walls through findfeasiblestepsize!, wolfe.jl:171-207)"""
    n = 2000
    lbs, ubs = -2.0 * np.ones(n), 0.8 * np.ones(n)
    ocfg, cfg, lsc = make_pair(flavour, ls, max_iters=60, eps=1e-6)
    f0 = cg.RosenbrockGPU(n, ctx)
    bar = cg.BoxBarrierGPU(f0, lbs, ubs, 5.0)
    ret = cg.minimizeobjective(bar, np.zeros(n), cfg, lsc)
    ora = O.minimize(O.Objective.box_barrier(O.Objective.rosenbrock(n), lbs, ubs, 5.0), np.zeros(n), ocfg)
    k = min(40, len(ora.trace_objective), len(ret.trace.objective))
    assert k >= 3 and ret.status == ora.status and abs(ret.iters_ran - ora.iters_ran) <= 2
    assert np.array_equal(ret.trace.step_size[:k], ora.trace_step_size[:k])
    assert np.array_equal(ret.trace.objective_evals[:k], ora.trace_objective_evals[:k])
    np.testing.assert_allclose(ret.trace.objective[:k], ora.trace_objective[:k], rtol=1e-10)
    np.testing.assert_allclose(ret.trace.grad_norm[:k], ora.trace_grad_norm[:k], rtol=1e-8)
    assert np.all(ret.minimizer < 0.8) and np.all(ret.minimizer > -2.0)
    bar.close(); f0.close()


@pytest.mark.parametrize("update", [False, True])
def test_primal_barrier_method_matches_oracle(ctx, update):
    n = 512
    lbs, ubs = -2.0 * np.ones(n), 0.8 * np.ones(n)
    pair = make_pair("HagerZhang", "StrongWolfeBisection", max_iters=500, eps=1e-4)
    ob = O.primalbarrier(O.Objective.rosenbrock(n), lbs, ubs, np.zeros(n), [pair[0]], 1e-2 * n, 20.0, 8, 1.0,
                         update_iterate=update)
    f0 = cg.RosenbrockGPU(n, ctx)
    b = cg.primalbarriermethod_(cg.setupCvxInequalityConstraint(2 * n, n), f0, cg.BoxConstraint(lbs, ubs), np.zeros(n),
                                pair[1], pair[2], cg.setupPrimalBarrierConfig(1e-2 * n, 20.0, 8, t_initial=1.0),
                                update_iterate=update)
    assert b.status == ob.status and b.iters_ran == ob.iters_ran and b.t_final == ob.t_final
    # device-resident restarts: no centering step (and no rerun attempt) uploads its starting point
    assert all(r.h2d_bytes == 0 and r.workspace is None for step in b.centering_results for r in step)
    for step, ostep in zip(b.centering_results, ob.centering_results):
        assert [r.status for r in step] == [o.status for o in ostep]
        assert abs(step[-1].iters_ran - ostep[-1].iters_ran) <= 2
        assert abs(step[-1].objective - ostep[-1].objective) <= 1e-8 * abs(ostep[-1].objective)
    if b.status == "success":
        x = b.centering_results[-1][-1].minimizer
        assert np.all(x < 0.8) and np.allclose(x, ob.centering_results[-1][-1].minimizer, atol=1e-5)
    # verifyt0 and the feasibility test
    b2 = cg.primalbarriermethod_(cg.setupCvxInequalityConstraint(2 * n, n), f0, cg.BoxConstraint(lbs, ubs),
                                 np.ones(n), pair[1], pair[2], cg.setupPrimalBarrierConfig(1e-2 * n, 20.0, 3))
    assert b2.status == "infeasible_start" and b2.iters_ran == 0
    t0 = cg.verifyt0(float("nan"), np.zeros(n), f0, 20.0, 0.0)
    assert t0 == O.Objective.rosenbrock(n).fdf(np.zeros(n))[0] * 20.0 or abs(t0 - n / 2 * 20.0) < 1e-9
    f0.close()


def test_rerun_starts_on_the_device(ctx):
    """minimizeobjectivererun (optim.jl:173-208): the second attempt starts from the first one's minimiser on the
    device (DeviceStart, one D2D copy) and gives bit for bit what a restart from the downloaded host copy gives."""
    n = 4096
    obj = cg.RosenbrockGPU(n, ctx)
    x0 = obj.default_x0(24, 0.1)
    _, cfg1, ls1 = make_pair("HagerZhang", "StrongWolfeBisection", max_iters=7)          # ends :max_iters_reached
    _, cfg2, ls2 = make_pair("LiuStorrey", "Wolfe", max_iters=40)
    rets = cg.minimizeobjectivererun(obj, x0, cfg1, ls1, (cfg2, ls2))
    assert len(rets) == 2 and rets[0].status == "max_iters_reached"
    assert rets[0].h2d_bytes == 8 * n and rets[1].h2d_bytes == 0
    a = cg.minimizeobjective(obj, x0, cfg1, ls1)
    b = cg.minimizeobjective(obj, a.minimizer, cfg2, ls2)                               # the host round trip
    assert np.array_equal(rets[0].minimizer, a.minimizer) and rets[1].status == b.status
    assert np.array_equal(rets[1].trace.objective, b.trace.objective) and np.array_equal(rets[1].minimizer, b.minimizer)
    ws = obj.make_workspace(x0)
    c = cg.minimizeobjective(obj, cg.DeviceStart(ws), cfg1, ls1)                        # DeviceStart by hand
    ws.close()
    assert c.h2d_bytes == 0 and np.array_equal(c.minimizer, a.minimizer)
    obj.close()
