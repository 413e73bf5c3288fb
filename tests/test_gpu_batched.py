"""GPU parity tests of the batched on-device solver (BASELINE.json configs[4]: many independent
n = 512 Rosenbrock problems, one CTA each).  The device restates the whole of minimizeobjective
(src/engine/optim.jl:6-171) + the three line searches (src/linesearch/nocedal.jl:33-209,
wolfe.jl:13-294, geometric.jl:22-186) + getβ (src/cg_flavours.jl); every problem must match the
oracle run on it alone, bit for bit, in the kernel's reduction order (B = cg.batched_lanes(n)
lanes, one tile: oracle.set_cgo_batched)."""
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


def _starts(nprob, n, seed=24, perturb=0.1):
    """x0 = standard start + 0.1·h(seed, problem, i)  (SURVEY.md §8d cfg 5)"""
    base = np.tile([-1.2, 1.0], n // 2)
    X = np.empty((nprob, n))
    for p in range(nprob):
        h = np.array([O.hash_u01(seed + p, i, 0) for i in range(n)])
        X[p] = base + perturb * (2.0 * h - 1.0)
    return X


def _check(ctx, X0, flavour, max_iters, linesearch="StrongWolfeBisection", **kw):
    ocfg, cfg, ls = make_pair(flavour, linesearch, max_iters=max_iters, **kw)
    res = cg.minimizeobjective_batched(X0, cfg, ls, ctx)
    n = X0.shape[1]
    O.set_cgo_batched(cg.batched_lanes(n))
    try:
        for p in range(X0.shape[0]):
            ora = O.minimize(O.Objective.rosenbrock(n), X0[p], ocfg)
            what = f"{flavour}/{linesearch} n={n} problem {p}"
            assert res.status[p] == ora.status, what
            assert res.iters_ran[p] == ora.iters_ran, what
            assert res.fdf_evals[p] == ora.fdf_evals_total, what
            assert res.objective[p] == ora.objective, what
            assert np.array_equal(res.minimizer[p], ora.minimizer), what
            assert res.grad_norm[p] == np.sqrt(O.dot(ora.gradient, ora.gradient, "cgo")), what
    finally:
        O.set_cgo_lanes(256)
    return res


def test_layout():
    assert [cg.batched_lanes(n) for n in (2, 64, 512, 514, 1024, 1026, 2048, 2050, 4096)] == \
        [32, 32, 32, 64, 64, 128, 128, 256, 256]


@pytest.mark.parametrize("flavour", ["HagerZhang", "YuanWangSheng", "SallehAlhawarat", "LiuStorrey"])
def test_n512_every_flavour_bit_exact(ctx, flavour):
    res = _check(ctx, _starts(12, 512), flavour, 1000)
    if flavour != "LiuStorrey":     # (LS ends in non_descent_search_direction on these starts; the oracle agrees)
        assert "success" in res.status


@pytest.mark.parametrize("n", [2, 10, 100, 130, 258, 510, 514, 1024, 1500, 2048, 3000, 4096])
def test_dimensions_bit_exact(ctx, n):
    """1, 2, 4 and 8 element pairs per lane on one warp, 2 / 4 / 8 warps beyond n = 512, partially
    filled lanes, a single pair."""
    _check(ctx, _starts(3, n, seed=7), "HagerZhang", 300)


def test_failure_statuses_match(ctx):
    """max_iters_reached, zoom_max_iters_reached and line-search limits are returned per problem
    exactly as the host engine would (optim.jl:93-121, :162-170)."""
    X0 = _starts(4, 64, seed=3)
    r1 = _check(ctx, X0, "HagerZhang", 5)
    assert set(r1.status) == {"max_iters_reached"}
    r2 = _check(ctx, X0, "LiuStorrey", 200, zoom_max_iters=1)
    assert "zoom_max_iters_reached" in r2.status
    r3 = _check(ctx, X0, "HagerZhang", 200, ls_max_iters=1, c2=1e-4, c1=1e-5)
    assert all(s in ("linesearch_max_iters_reached", "zoom_max_iters_reached", "success") for s in r3.status)


@pytest.mark.parametrize("linesearch", ["Wolfe", "YuanWeiLuWolfe", "Backtracking"])
@pytest.mark.parametrize("flavour", ["HagerZhang", "SallehAlhawarat"])
def test_other_linesearches_bit_exact(ctx, linesearch, flavour):
    """WolfeBisection with both conditions (wolfe.jl, including the mid-search reset u ← −df_x that
    keeps the old dϕ_0) and Backtracking{Armijo} (geometric.jl, including the adoption of the
    rejected trial point), restated on the device."""
    res = _check(ctx, _starts(6, 512, seed=13), flavour, 400, linesearch)
    res2 = _check(ctx, _starts(3, 1500, seed=17), flavour, 150, linesearch)
    assert len(set(res.status) | set(res2.status)) >= 1


def test_other_linesearch_failure_statuses(ctx):
    X0 = _starts(4, 64, seed=3)
    r = _check(ctx, X0, "HagerZhang", 200, "Wolfe", max_step_size=1e-3)
    assert set(r.status) <= {"max_step_length_reached", "success", "max_iters_reached", "linesearch_max_iters_reached"}
    r = _check(ctx, X0, "HagerZhang", 200, "Backtracking", ls_max_iters=2)
    assert len(r.status) == 4
    r = _check(ctx, X0, "HagerZhang", 50, "YuanWeiLuWolfe", ls_max_iters=3)
    assert len(r.status) == 4


def test_close_to_single_problem_device_path(ctx):
    """The batched kernel (one warp per problem) and the one-problem fused path (296 virtual CTAs of
    256 lanes) reduce in different orders: same decisions early on, objectives within 1e-10."""
    n = 512
    X0 = _starts(2, n, seed=11)
    _, cfg, ls = make_pair("HagerZhang", max_iters=12)
    res = cg.minimizeobjective_batched(X0, cfg, ls, ctx)
    for p in range(2):
        one = cg.minimizeobjective(cg.RosenbrockGPU(n, ctx), X0[p], cfg, ls)
        assert one.status == res.status[p] and one.iters_ran == res.iters_ran[p]
        assert abs(one.objective - res.objective[p]) <= 1e-10 * abs(one.objective)


@pytest.mark.parametrize("nprob", [32_768, 262_144])
def test_large_batch_properties(ctx, nprob):
    """32,768 problems (one GPU's share of cfg 5 at 8 GPUs) and the full 262,144 of BASELINE.json
    configs[4]: identical starts give identical results whichever CTA runs them, every problem
    converges to ones(n), and the run is reproducible."""
    n = 512
    X0 = np.tile(_starts(4, n, seed=5), (nprob // 4, 1))
    _, cfg, ls = make_pair("HagerZhang", max_iters=1000)
    a = cg.minimizeobjective_batched(X0, cfg, ls, ctx)
    b = cg.minimizeobjective_batched(X0, cfg, ls, ctx)
    assert np.array_equal(a.objective, b.objective) and np.array_equal(a.minimizer, b.minimizer)
    for k in range(4):
        assert np.all(a.objective[k::4] == a.objective[k]) and np.all(a.iters_ran[k::4] == a.iters_ran[k])
    assert set(a.status) == {"success"}
    assert np.allclose(a.minimizer, 1.0, atol=1e-4)


def test_config_and_shape_errors(ctx):
    _, cfg, ls = make_pair("HagerZhang")
    with pytest.raises(cg.CgoError):
        cg.minimizeobjective_batched(np.zeros((2, 3)), cfg, ls, ctx)          # odd n
    with pytest.raises(cg.CgoError):
        cg.minimizeobjective_batched(np.zeros((2, 4098)), cfg, ls, ctx)       # n > 4096
    _, cfg2, ls2 = make_pair("LBFGS")
    with pytest.raises(TypeError):
        cg.minimizeobjective_batched(np.zeros((2, 4)), cfg2, ls2, ctx)
    with pytest.raises(TypeError):
        cg.minimizeobjective_batched(np.zeros((2, 4)), cfg, object(), ctx)
