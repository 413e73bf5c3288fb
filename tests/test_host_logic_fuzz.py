"""Property-based cross-check of the two independent restatements of the reference's scalar logic: the
Python host mirror (engine + line searches + flavours, driven through the test-only numpy workspace)
and the C oracle.  They share the objective code and the reduction order, so for ANY configuration
and start every run must agree bit for bit — statuses, counts, step sizes, traces, minimiser."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import cgoptim_b200 as cg
from oracle import oracle as O

from helpers import FLAVOURS, LINESEARCHES, assert_same_run, make_pair
from numpy_workspace import NumpyBoxBarrier, NumpyObjective

CG_FLAVOURS = [f for f in FLAVOURS if f != "LBFGS"]
finite = dict(allow_nan=False, allow_infinity=False)


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.too_slow], derandomize=True)
@given(flavour=st.sampled_from(FLAVOURS), linesearch=st.sampled_from(LINESEARCHES),
       x=st.floats(-6.0, 6.0, **finite), y=st.floats(-6.0, 6.0, **finite),
       c1=st.floats(1e-6, 0.3, **finite), gap=st.floats(0.05, 0.6, **finite),
       eps=st.sampled_from([1e-3, 1e-5, 1e-8]), max_iters=st.integers(1, 60), mu=st.floats(0.01, 0.9, **finite))
def test_minimizeobjective_booth_any_config(flavour, linesearch, x, y, c1, gap, eps, max_iters, mu):
    c2 = min(c1 + gap, 0.99)
    kw = dict(eps=eps, max_iters=max_iters, mu=mu, c1=c1, c2=c2)
    if linesearch == "YuanWeiLuWolfe":
        kw["delta1"] = c1 / 10
    ocfg, cfg, ls = make_pair(flavour, linesearch, **kw)
    x0 = np.array([x, y])
    ora = O.minimize(O.Objective.booth(), x0, ocfg)
    ret = cg.minimizeobjective(NumpyObjective(O.Objective.booth()), x0, cfg, ls)
    assert_same_run(ret, ora, what=f"{flavour}/{linesearch} x0={x0} {kw}")


@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.too_slow], derandomize=True)
@given(flavour=st.sampled_from(FLAVOURS), linesearch=st.sampled_from(LINESEARCHES), n2=st.integers(1, 12),
       seed=st.integers(0, 1000), perturb=st.floats(0.0, 0.5, **finite), max_iters=st.integers(1, 40))
def test_minimizeobjective_rosenbrock_any_size(flavour, linesearch, n2, seed, perturb, max_iters):
    n = 2 * n2
    ocfg, cfg, ls = make_pair(flavour, linesearch, max_iters=max_iters)
    x0 = O.rosenbrock_x0(n, seed, perturb)
    ora = O.minimize(O.Objective.rosenbrock(n), x0, ocfg)
    ret = cg.minimizeobjective(NumpyObjective(O.Objective.rosenbrock(n)), x0, cfg, ls)
    assert_same_run(ret, ora, what=f"{flavour}/{linesearch} n={n} seed={seed}")


@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.too_slow], derandomize=True)
@given(flavour=st.sampled_from(CG_FLAVOURS), fix=st.booleans(), x=st.floats(-4.0, 4.0, **finite),
       y=st.floats(-4.0, 4.0, **finite), s=st.floats(0.1, 4.0, **finite), sigma=st.floats(0.01, 2.0, **finite),
       rho=st.floats(0.3, 0.97, **finite), max_iters=st.integers(1, 40))
def test_solvesystem_any_config(flavour, fix, x, y, s, sigma, rho, max_iters):
    ocfg, cfg, _ = make_pair(flavour, max_iters=max_iters)
    x0 = np.array([x, y])
    ora = O.solvesystem(O.Objective.booth(), x0, ocfg, O.solvesys_ls(s, sigma, rho, 40, fix_stale_iterate=fix))
    ret = cg.solvesystem(NumpyObjective(O.Objective.booth()), x0, cfg,
                         cg.setupLinesearchSolveSys(s, σ=sigma, ρ=rho, max_iters=40), fix_stale_iterate=fix)
    assert_same_run(ret, ora, what=f"solvesystem {flavour} fix={fix} x0={x0}")


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.too_slow], derandomize=True)
@given(linesearch=st.sampled_from(LINESEARCHES), update=st.booleans(), half=st.floats(3.5, 20.0, **finite),
       growth=st.floats(2.0, 50.0, **finite), t0=st.sampled_from([float("nan"), 0.5, 5.0]))
def test_primal_barrier_any_config(linesearch, update, half, growth, t0):
    pair = make_pair("HagerZhang", linesearch, max_iters=200)
    lbs, ubs, x0 = np.array([-half, -half]), np.array([half, half]), np.array([0.43, 1.23])
    ob = O.primalbarrier(O.Objective.booth(), lbs, ubs, x0, [pair[0]], 1e-4, growth, 8, t0, update_iterate=update)
    b = cg.primalbarriermethod_(cg.setupCvxInequalityConstraint(4, 2), NumpyObjective(O.Objective.booth()),
                                cg.BoxConstraint(lbs, ubs), x0, pair[1], pair[2],
                                cg.setupPrimalBarrierConfig(1e-4, growth, 8, t_initial=t0),
                                update_iterate=update, make_barrier=NumpyBoxBarrier)
    assert b.status == ob.status and b.iters_ran == ob.iters_ran and b.total_objective_evals == ob.total_objective_evals
    for step, ostep in zip(b.centering_results, ob.centering_results):
        for r, o in zip(step, ostep):
            assert r.status == o.status and np.array_equal(r.trace.objective, o.trace_objective)
            assert np.array_equal(r.minimizer, o.minimizer, equal_nan=True)
