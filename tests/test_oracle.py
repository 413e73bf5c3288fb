"""CPU tests of the oracle itself (oracle/cgo_oracle.c), the thing every GPU parity test leans on.

PARITY UNPINNED against the reference (Julia is not installed; the reference's only test,
test/runtests.jl:7-44, checks the Booth gradient and never calls the solver).  What CAN be pinned
is checked here: (1) every known-answer anchor the reference holds, (2) the committed golden
traces (tests/golden/traces.json, restatement-derived), (3) independent cross-checks of the
pieces that have textbook definitions (finite-difference gradients, scipy CSR products, an
explicit dense BFGS inverse for the L-BFGS two-loop recursion)."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as O

from helpers import LINESEARCHES, make_pair

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "traces.json")


# ------------------------------------------------------------------ reference-held anchors
def test_booth_gradient_zero_at_minimiser():
    """test/runtests.jl:17-21: x_oracle = [1, 3], ‖∇f‖ < 1e-12 (and f* = 0)."""
    f, g = O.Objective.booth().fdf(np.array([1.0, 3.0]))
    assert f == 0.0 and np.linalg.norm(g) < 1e-12


def _central_fd(fun, x, h=1e-3):
    """8th-order central difference (test/runtests.jl:27-37 uses central_fdm(8, 1))."""
    c = np.array([4 / 5, -1 / 5, 4 / 105, -1 / 280])
    g = np.zeros_like(x)
    for i in range(x.size):
        for k, ck in enumerate(c, start=1):
            e = np.zeros_like(x)
            e[i] = k * h
            g[i] += ck * (fun(x + e) - fun(x - e)) / h
    return g


@pytest.mark.parametrize("maker,n", [(lambda: O.Objective.booth(), 2), (lambda: O.Objective.rosenbrock(10), 10),
                                     (lambda: O.Objective.rosenbrock_chained(7), 7),
                                     (lambda: O.Objective.sparse_ls(40, 4, 6, 24, 0), 40),
                                     (lambda: O.Objective.logreg(60, 12, 4, 24, 1e-3), 12)])
def test_gradients_match_finite_differences(maker, n):
    """test/runtests.jl:23-42 (N_tests = 10 random points, tolerance 1e-5), extended from Booth
    to every objective the oracle defines."""
    obj = maker()
    rng = np.random.default_rng(24)                                 # the author's seed, examples/min.jl:7
    for _ in range(10):
        x = rng.standard_normal(n) * 0.5
        _, g = obj.fdf(x)
        g_nd = _central_fd(lambda z: obj.fdf(z)[0], x)
        assert np.linalg.norm(g - g_nd) < 1e-5 * max(1.0, np.linalg.norm(g))


def test_canonical_example_run():
    """examples/min.jl:16-43: HagerZhang, StrongWolfeBisection(1e-5, 0.8; growth 2, 1000, 100),
    ϵ = 1e-5, max_iters = 1000, x0 = [0.43, 1.23] → :success at the Booth minimiser [1, 3]."""
    cfg = O.make_config("HagerZhang", "StrongWolfeBisection", eps=1e-5, max_iters=1000, c1=1e-5, c2=0.8,
                        growth=2.0, ls_max_iters=1000, zoom_max_iters=100)
    r = O.minimize(O.Objective.booth(), np.array([0.43, 1.23]), cfg)
    assert r.status == "success"
    assert np.allclose(r.minimizer, [1.0, 3.0], atol=1e-5) and r.objective < 1e-10
    assert np.linalg.norm(r.gradient) < 1e-5
    assert r.iters_ran == len(r.trace_objective) < 50               # trace truncated, types.jl:148


@pytest.mark.parametrize("n", [2, 10, 1000])
def test_rosenbrock_minimiser_is_ones(n):
    """examples/helpers/test_funcs.jl:48: the minimiser of Rosenbrock is ones(d)."""
    for obj in (O.Objective.rosenbrock(n), O.Objective.rosenbrock_chained(n)):
        f, g = obj.fdf(np.ones(n))
        assert f == 0.0 and np.all(g == 0.0)
    cfg = O.make_config(max_iters=2000)
    r = O.minimize(O.Objective.rosenbrock(n), O.rosenbrock_x0(n, 24, 0.0), cfg)
    assert r.status == "success" and np.allclose(r.minimizer, 1.0, atol=1e-4)


# ------------------------------------------------------------------ golden traces
def _cases():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


def _unhex(a):
    return np.array([float.fromhex(v) for v in a])


def _make_objective(name):
    if name == "booth":
        return O.Objective.booth(), np.array([0.43, 1.23])
    if name.startswith("rosenbrock_n"):
        n = int(name.split("_n")[1])
        return O.Objective.rosenbrock(n), O.rosenbrock_x0(n, 24, 0.0)
    if name == "sparse_ls_n2000":
        return O.Objective.sparse_ls(2000, 10, 64, 24, 0), np.zeros(2000)
    if name == "logreg_3000x500":
        return O.Objective.logreg(3000, 500, 20, 24, 1e-4), np.zeros(500)
    raise KeyError(name)


@pytest.mark.parametrize("idx", range(54))
def test_oracle_reproduces_golden_trace(idx):
    c = _cases()[idx]
    obj, x0 = _make_objective(c["name"])
    ocfg, _, _ = make_pair(c["flavour"], c["linesearch"], max_iters=c["max_iters"], sum_mode=c["sum_mode"],
                           beta_form=c["beta_form"])
    r = O.minimize(obj, x0, ocfg)
    k = len(c["trace_objective"])
    assert r.status == c["status"] and r.iters_ran == c["iters_ran"]
    assert [int(v) for v in r.trace_objective_evals[:k]] == c["trace_objective_evals"]
    if c["name"].startswith("logreg"):          # libm exp / log1p may differ in the last ulp between hosts
        np.testing.assert_allclose(r.trace_objective[:k], _unhex(c["trace_objective"]), rtol=1e-9)
        np.testing.assert_allclose(r.trace_step_size[:k], _unhex(c["trace_step_size"]), rtol=1e-9)
    else:                                       # IEEE +,−,×,÷,√ only: bit-exact
        assert np.array_equal(r.trace_objective[:k], _unhex(c["trace_objective"]))
        assert np.array_equal(r.trace_grad_norm[:k], _unhex(c["trace_grad_norm"]))
        assert np.array_equal(r.trace_step_size[:k], _unhex(c["trace_step_size"]))
        assert r.objective == float.fromhex(c["objective"])
        assert np.array_equal(r.minimizer[:8], _unhex(c["minimizer_head"]))


def test_golden_case_count():
    assert len(_cases()) == 54


# ------------------------------------------------------------------ independent cross-checks
def test_csr_products_match_scipy():
    ora = O.Objective.sparse_ls(500, 7, 40, 3, 0)
    rp, ci, va = ora.csr(False)
    A = sp.csr_matrix((va, ci, rp), shape=(500, 500))
    rpT, ciT, vaT = ora.csr(True)
    AT = sp.csr_matrix((vaT, ciT, rpT), shape=(500, 500))
    assert abs(A.T - AT).max() == 0.0
    # rows of the transpose are sorted by source row (the sequential scatter order)
    for j in range(500):
        seg = ciT[rpT[j]:rpT[j + 1]]
        assert np.all(np.diff(seg) > 0)
    x = np.random.default_rng(0).standard_normal(500)
    np.testing.assert_allclose(ora.spmv(x), A @ x, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(ora.spmv(x, transposed=True), A.T @ x, rtol=1e-13, atol=1e-13)
    f, g = ora.fdf(x)
    r = A @ x - ora.rhs()
    assert abs(f - 0.5 * r @ r) <= 1e-12 * f
    np.testing.assert_allclose(g, A.T @ r, rtol=1e-12, atol=1e-12)
    xt = O.sparse_ls_xtrue(500, 3)
    np.testing.assert_allclose(A @ xt, ora.rhs(), rtol=1e-13)


def test_generator_structure():
    """Banded-random generator (spec text in oracle/cgo_oracle.c): one entry per stratum, distinct
    columns, |col − row| ≤ W cyclically, diagonally dominant; coh_log2 = 30 gives true diagonals."""
    n, K, W = 4000, 10, 512
    for coh in (0, 4, 30):
        rp, ci, va = O.Objective.sparse_ls(n, K, W, 24, coh).csr(False)
        assert np.array_equal(rp, np.arange(n + 1) * K)
        C_ = ci.reshape(n, K).astype(np.int64)
        rows = np.arange(n)[:, None]
        d = (C_ - rows + n // 2) % n - n // 2
        assert np.all(d[:, 0] == 0) and np.all(np.abs(d) <= W)
        assert all(len(set(r)) == K for r in C_[:50])
        V = va.reshape(n, K)
        assert np.all(V[:, 0] >= 4.0) and np.all(np.abs(V[:, 1:]) <= 0.3)
        if coh == 30:
            assert np.all(d == d[0])
        if coh == 4:
            assert np.all(d[0:16] == d[0]) and not np.all(d == d[0])


def test_sum_modes_agree():
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal(100_003), rng.standard_normal(100_003)
    ref = O.dot(a, b, "comp")
    for mode in ("seq", "pairwise", "cgo"):
        assert abs(O.dot(a, b, mode) - ref) <= 1e-12 * np.sqrt(a.size) * 10
    # the canonical order does not depend on the thread count and is shard-additive
    O.set_cgo_order(7, 1)
    one = O.dot(a[:100_000], b[:100_000], "cgo")
    O.set_cgo_order(7, 4)
    four = O.dot(a[:100_000], b[:100_000], "cgo")
    O.set_cgo_order(296, 1)
    assert abs(one - four) <= 1e-11 * abs(one) + 1e-9


def test_lbfgs_direction_matches_dense_bfgs_inverse():
    """The new LBFGS flavour has no reference counterpart: pin its two-loop
    recursion (N&W Alg. 7.4) against the explicit product form of the inverse BFGS update
    H⁺ = (I − ρ s yᵀ) H (I − ρ y sᵀ) + ρ s sᵀ, H₀ = γ I, on a small strictly convex problem."""
    n, m = 12, 4
    ora = O.Objective.sparse_ls(n, 3, 2, 5, 0)
    x0 = np.linspace(-1, 1, n)
    # run k iterations with max_iters = k and k+1; the (k+1)-th step is x_{k+1} − x_k = a·u_k
    runs = {}
    for k in range(1, 7):
        cfg = O.make_config("LBFGS", "StrongWolfeBisection", eps=1e-14, max_iters=k, lbfgs_m=m, c1=1e-4, c2=0.9)
        runs[k] = O.minimize(ora, x0, cfg)
    xs = [x0] + [runs[k].minimizer for k in range(1, 7)]
    gs = [ora.fdf(x)[1] for x in xs]
    for k in range(1, 6):
        S = [xs[j + 1] - xs[j] for j in range(max(0, k - m), k)]
        Y = [gs[j + 1] - gs[j] for j in range(max(0, k - m), k)]
        gamma = (S[-1] @ Y[-1]) / (Y[-1] @ Y[-1])
        H = gamma * np.eye(n)
        for s, y in zip(S, Y):
            rho = 1.0 / (y @ s)
            V = np.eye(n) - rho * np.outer(y, s)
            H = V.T @ H @ V + rho * np.outer(s, s)
        u = -H @ gs[k]
        step = xs[k + 1] - xs[k]
        a = runs[k + 1].trace_step_size[k]
        np.testing.assert_allclose(step, a * u, rtol=1e-7, atol=1e-10 * np.linalg.norm(u))


@pytest.mark.parametrize("linesearch", LINESEARCHES)
def test_statuses_reachable(linesearch):
    """Failure statuses are returned, never raised (optim.jl:93-121): the barrier objective is
    non-finite outside |x| < 1 and drives the feasibility back-off paths."""
    ocfg, _, _ = make_pair("HagerZhang", linesearch, max_iters=3)
    r = O.minimize(O.Objective.booth(), np.array([0.43, 1.23]), ocfg)
    assert r.status in ("max_iters_reached", "success")
    ocfg, _, _ = make_pair("HagerZhang", linesearch, max_iters=200)
    r = O.minimize(O.Objective.barrier(6), np.linspace(-0.5, 0.5, 6), ocfg)
    assert r.status in O.STATUS
    if linesearch != "Backtracking":
        # (Backtracking adopts the REJECTED trial with the previous ϕ, geometric.jl:141-144 →
        # optim.jl:136-139, SURVEY.md §8a LS-3: it can step outside the feasible box; replicated)
        assert np.all(np.abs(r.minimizer) < 1.0)


# ------------------------------------------------------------------ golden traces of the callers
CALLERS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "traces_callers.json")


def _caller_cases(kind):
    with open(CALLERS) as f:
        return [c for c in json.load(f)["cases"] if c["kind"] == kind]


@pytest.mark.parametrize("idx", range(16))
def test_oracle_reproduces_golden_solvesystem(idx):
    """solvesystem (src/engine/solve_system.jl) on dev/solve_sys.jl's settings, as written and with
    fix_stale_iterate (tests/golden/traces_callers.json, restatement-derived)."""
    c = _caller_cases("solvesystem")[idx]
    cfg = O.make_config(c["flavour"], max_iters=c["max_iters"], sum_mode=c["sum_mode"], beta_form=c["beta_form"], mu=0.1)
    r = O.solvesystem(O.Objective.booth(), np.array([0.43, 1.23]), cfg,
                      O.solvesys_ls(1.0, fix_stale_iterate=c["fix_stale_iterate"]))
    k = len(c["trace_objective"])
    assert r.status == c["status"] and r.iters_ran == c["iters_ran"] and r.fdf_evals_total == c["fdf_evals_total"]
    assert [int(v) for v in r.trace_objective_evals[:k]] == c["trace_objective_evals"]
    assert np.array_equal(r.trace_objective[:k], _unhex(c["trace_objective"]))
    assert np.array_equal(r.trace_grad_norm[:k], _unhex(c["trace_grad_norm"]))
    assert np.array_equal(r.trace_step_size[:k], _unhex(c["trace_step_size"]))
    assert np.array_equal(r.minimizer, _unhex(c["minimizer"]))


@pytest.mark.parametrize("idx", range(2))
def test_oracle_reproduces_golden_primal_barrier(idx):
    """primalbarriermethod! on examples/constrained.jl's problem (log terms: libm, so the scalar
    outcomes are pinned exactly and the objectives to 1e-9)."""
    c = _caller_cases("primalbarrier")[idx]
    cfgs = [O.make_config("HagerZhang", "Wolfe", c1=1e-3, c2=0.9, ls_max_iters=100, sum_mode="cgo", beta_form="fused"),
            O.make_config("LiuStorrey", "Backtracking", c1=1e-3, c2=0.9, ls_max_iters=300, sum_mode="cgo", beta_form="fused")]
    b = O.primalbarrier(O.Objective.booth(), [-10.0, -10.0], [10.0, 10.0], [0.43, 1.23], cfgs, 1e-8, 10.0, 100,
                        update_iterate=c["update_iterate"])
    assert b.status == c["status"] and b.iters_ran == c["iters_ran"] and b.t_final == float.fromhex(c["t_final"])
    assert [len(s) for s in b.centering_results] == c["attempts"]
    assert [s[-1].status for s in b.centering_results] == c["final_statuses"]
    np.testing.assert_allclose([s[-1].objective for s in b.centering_results], _unhex(c["final_objectives"]), rtol=1e-9)


@pytest.mark.parametrize("idx", range(4))
def test_oracle_reproduces_golden_batched_order(idx):
    """the batched solver's reduction order (32 lanes, one tile: oracle.set_cgo_batched)"""
    c = _caller_cases("batched_order")[idx]
    ocfg, _, _ = make_pair("HagerZhang", c["linesearch"], max_iters=200)
    O.set_cgo_batched(c["lanes"])
    try:
        r = O.minimize(O.Objective.rosenbrock(512), O.rosenbrock_x0(512, 24, 0.1), ocfg)
    finally:
        O.set_cgo_lanes(256)
    assert r.status == c["status"] and r.iters_ran == c["iters_ran"] and r.fdf_evals_total == c["fdf_evals_total"]
    assert r.objective == float.fromhex(c["objective"])
    assert np.array_equal(r.trace_objective[:len(c["trace_objective"])], _unhex(c["trace_objective"]))
