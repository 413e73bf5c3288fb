"""GPU parity tests (through the C ABI) of the CSR least-squares objective ½‖Ax − b‖²
(BASELINE.json configs[2]): generator, explicit transpose, SpMV / SpMVᵀ kernels, the fused trial
pack and whole CG runs, all BIT-EXACT against the oracle in canonical-order mode; plus
size-independent properties at the full n = 2e8."""
import os

import numpy as np
import pytest

import cgoptim_b200 as cg
from oracle import oracle as O

from helpers import gate_numbers, record_gate, FLAVOURS, assert_same_run, make_pair
from numpy_workspace import NumpyObjective

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cg.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n,K,W,coh", [(3000, 10, 64, 0), (3000, 10, 64, 3), (20_000, 10, 4096, 0),
                                       (1000, 1, 0, 0), (5002, 4, 9, 0), (70_000, 10, 30_000, 0)])
def test_generator_and_transpose_match_oracle(ctx, n, K, W, coh):
    """Device generator == oracle generator; explicit transpose == stable counting-sort transpose."""
    obj = cg.SparseLSGPU(n, K, W, 24, coh, ctx)
    ora = O.Objective.sparse_ls(n, K, W, 24, coh)
    rp, ci, va, b = obj.csr(False)
    orp, oci, ova = ora.csr(False)
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(va, ova)
    assert np.array_equal(b, ora.rhs())
    rpT, ciT, vaT, _ = obj.csr(True)
    orpT, ociT, ovaT = ora.csr(True)
    assert np.array_equal(rpT, orpT) and np.array_equal(ciT, ociT) and np.array_equal(vaT, ovaT)
    obj.close()


@pytest.mark.parametrize("n", [3000, 100_000])
def test_spmv_bit_exact_and_adjoint(ctx, n):
    obj = cg.SparseLSGPU(n, 10, None if n > 5000 else 64, 24, 0, ctx)
    ora = O.Objective.sparse_ls(n, 10, None if n > 5000 else 64, 24, 0)
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    Ax, ATy = obj.spmv(x), obj.spmv(y, transposed=True)
    assert np.array_equal(Ax, ora.spmv(x)) and np.array_equal(ATy, ora.spmv(y, transposed=True))
    lhs, rhs = float(Ax @ y), float(x @ ATy)                      # ⟨Ax, y⟩ = ⟨x, Aᵀy⟩
    assert abs(lhs - rhs) <= 1e-12 * (np.linalg.norm(Ax) * np.linalg.norm(y))
    obj.close()


def _ragged_csr(nrows, ncols, seed, long_row=None):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 9, nrows)
    lens[rng.integers(0, nrows, nrows // 10)] = 0                  # empty rows
    if long_row is not None:
        lens[long_row] = 7000                                      # spans several 2560-entry chunks
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    nnz = int(rowptr[-1])
    col = rng.integers(0, ncols, nnz).astype(np.int32)             # duplicates allowed
    val = rng.standard_normal(nnz)
    b = rng.standard_normal(nrows)
    return rowptr, col, val, b


@pytest.mark.parametrize("nrows,ncols,long_row", [(1, 1, None), (257, 300, None), (5000, 3001, 17),
                                                  (300, 4000, 299), (4096, 4096, 0)])
def test_from_csr_ragged_bit_exact(ctx, nrows, ncols, long_row):
    """User-supplied CSR with empty rows, duplicate columns, rows longer than a chunk and a
    rectangular shape: SpMV, SpMVᵀ, f and g identical to the oracle."""
    rowptr, col, val, b = _ragged_csr(nrows, ncols, 5, long_row)
    obj = cg.SparseLSGPU_from_csr(nrows, ncols, rowptr, col, val, b, ctx)
    ora = O.Objective.sparse_ls_csr(nrows, ncols, rowptr, col, val, b)
    ora.set_trial_site(*obj.trial_site)
    ora.set_sum_mode("cgo")
    rpT, ciT, vaT, _ = obj.csr(True)
    orpT, ociT, ovaT = ora.csr(True)
    assert np.array_equal(rpT, orpT) and np.array_equal(ciT, ociT) and np.array_equal(vaT, ovaT)
    rng = np.random.default_rng(2)
    x = rng.standard_normal(ncols)
    assert np.array_equal(obj.spmv(x), ora.spmv(x))
    ws = obj.make_workspace(x, fuse_direction=False)
    f, g = ora.fdf(x)
    assert ws.f_x0 == f
    assert np.array_equal(ws.download()[1], g)
    ws.close()
    obj.close()


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("nrows,ncols,long_row", [(1, 1, None), (33, 40, 5), (257, 300, None), (5000, 3001, 17),
                                                  (4100, 4100, 4099)])
def test_from_csr_ragged_both_kernel_families(mode, nrows, ncols, long_row):
    """The same ragged inputs with the kernel family forced (cgo_ctx_set_csr_mode: 1 = k_csr_rows, TMA ring, fused
    dots; 2 = sliced layout + k_spmv_direct, dots in BLAS-1 passes): empty rows, a last slice of fewer than 32 rows,
    a 7000-entry row (many batches of either width), slices of unequal lengths.  SpMV, SpMVᵀ, the downloaded
    matrices, f, g and a short CG run must equal the oracle at the objective's own reduction site."""
    c = cg.Context(0)
    c.set_csr_mode(mode)
    try:
        rowptr, col, val, b = _ragged_csr(nrows, ncols, 5, long_row)
        obj = cg.SparseLSGPU_from_csr(nrows, ncols, rowptr, col, val, b, c)
        assert obj.trial_site == ((2, 4) if mode == 2 else (1, 1))
        ora = O.Objective.sparse_ls_csr(nrows, ncols, rowptr, col, val, b)
        ora.set_trial_site(*obj.trial_site)
        ora.set_sum_mode("cgo")
        rp, ci, va, bb = obj.csr(False)                   # a sliced matrix comes back row-major
        assert np.array_equal(rp, rowptr) and np.array_equal(ci, col) and np.array_equal(va, val) and np.array_equal(bb, b)
        rpT, ciT, vaT, _ = obj.csr(True)
        orpT, ociT, ovaT = ora.csr(True)
        assert np.array_equal(rpT, orpT) and np.array_equal(ciT, ociT) and np.array_equal(vaT, ovaT)
        rng = np.random.default_rng(7)
        x, y = rng.standard_normal(ncols), rng.standard_normal(nrows)
        assert np.array_equal(obj.spmv(x), ora.spmv(x))
        assert np.array_equal(obj.spmv(y, transposed=True), ora.spmv(y, transposed=True))
        ws = obj.make_workspace(x, fuse_direction=False)
        f, g = ora.fdf(x)
        assert ws.f_x0 == f and np.array_equal(ws.download()[1], g)
        ws.close()
        if nrows > 1:
            ocfg, cfg, ls = make_pair("HagerZhang", max_iters=8)
            assert_same_run(cg.minimizeobjective(obj, np.zeros(ncols), cfg, ls), O.minimize(ora, np.zeros(ncols), ocfg))
        obj.close()
    finally:
        c.close()


@pytest.mark.parametrize("mode,coh", [(1, 0), (2, 30), (2, 3)])
def test_synthetic_matrix_on_the_other_kernel_family(mode, coh):
    """The generator picks the kernel family from the coherence of the offsets; forcing the other one must change
    nothing but the site of the Σr² reduction (which the oracle follows): whole runs stay bit-exact."""
    c = cg.Context(0)
    c.set_csr_mode(mode)
    try:
        n = 50_000
        obj = cg.SparseLSGPU(n, 10, 4096, 24, coh, c)
        assert obj.trial_site == ((2, 4) if mode == 2 else (1, 1))
        ora = O.Objective.sparse_ls(n, 10, 4096, 24, coh)
        ora.set_trial_site(*obj.trial_site)
        for flavour in ("HagerZhang", "LBFGS"):
            ocfg, cfg, ls = make_pair(flavour, max_iters=12)
            assert_same_run(cg.minimizeobjective(obj, np.zeros(n), cfg, ls), O.minimize(ora, np.zeros(n), ocfg))
        obj.close()
    finally:
        c.close()


def test_from_csr_rejects_bad_input(ctx):
    rowptr, col, val, b = _ragged_csr(10, 10, 1)
    col2 = col.copy()
    if col2.size:
        col2[0] = 10
        with pytest.raises(cg.CgoError):
            cg.SparseLSGPU_from_csr(10, 10, rowptr, col2, val, b, ctx)
    rp2 = rowptr.copy()
    rp2[3], rp2[4] = rp2[4] + 1, rp2[3]
    with pytest.raises(cg.CgoError):
        cg.SparseLSGPU_from_csr(10, 10, rp2, col, val, b, ctx)


@pytest.mark.parametrize("n,G", [(3000, 296), (3000, 3), (100_000, 1184), (400_000, 7)])
def test_trial_pack_bit_exact(ctx, n, G):
    """evalϕdϕ! (cg_utils.jl:3-22) on the CSR objective == oracle fdf + canonical-order dots."""
    ctx.set_reduction_ctas(G)
    O.set_cgo_order(G, 1)
    try:
        W = 64 if n < 5000 else None
        obj = cg.SparseLSGPU(n, 10, W, 24, 0, ctx)
        nws = NumpyObjective(O.Objective.sparse_ls(n, 10, W, 24, 0)).make_workspace(np.zeros(n))
        ws = obj.make_workspace(np.zeros(n), fuse_direction=False)
        assert ws.f_x0 == nws.f_x0 and ws.norm_df_x0 == nws.norm_df_x0
        ws.reset_direction(); nws.reset_direction()
        assert ws.dot_g_u() == nws.dot_g_u() and ws.dot_u_u() == nws.dot_u_u()
        for a in (0.05, 0.02):
            assert ws.eval_trial(a) == nws.eval_trial(a)
            assert np.array_equal(ws.pack[:9], nws.pack[:9])
        assert np.array_equal(ws.download_vector("df_xp"), nws.gp_)
        ws.accept(); nws.accept()
        ws.update_dir(0.3); nws.update_dir(0.3)
        assert ws.dot_g_u() == nws.dot_g_u() and ws.dot_u_u() == nws.dot_u_u()
        assert ws.eval_trial(0.01) == nws.eval_trial(0.01)
        assert np.array_equal(ws.pack[:9], nws.pack[:9])
        ws.close()
        obj.close()
    finally:
        ctx.set_reduction_ctas(296)
        O.set_cgo_order(296, 1)


@pytest.mark.parametrize("flavour", FLAVOURS)
def test_full_run_bit_exact(ctx, flavour):
    n = 20_000
    ocfg, cfg, ls = make_pair(flavour, max_iters=200)
    obj = cg.SparseLSGPU(n, 10, 2048, 24, 0, ctx)
    ora = O.minimize(O.Objective.sparse_ls(n, 10, 2048, 24, 0), np.zeros(n), ocfg)
    ret = cg.minimizeobjective(obj, np.zeros(n), cfg, ls)
    assert ora.status == "success"
    assert_same_run(ret, ora, what=flavour)
    # the problem is consistent (b = A x_true): the minimiser is x_true
    assert np.linalg.norm(ret.minimizer - O.sparse_ls_xtrue(n, 24)) <= 1e-4 * np.sqrt(n)
    obj.close()


def test_north_star_gates_vs_reference_shaped_oracle(ctx):
    """north_star's gates against the oracle in the reference's own shape (sequential sums,
    literal β, unfused passes): f and ‖g‖ within 1e-10 relative over the first 50 iterations,
    identical step sizes / fdf-eval counts, final objective within 1e-8, iterations ±2."""
    n = 50_000
    ocfg, cfg, ls = make_pair("HagerZhang", sum_mode="seq", beta_form="literal", max_iters=400)
    obj = cg.SparseLSGPU(n, 10, 8192, 24, 0, ctx)
    ora = O.minimize(O.Objective.sparse_ls(n, 10, 8192, 24, 0), np.zeros(n), ocfg)
    ret = cg.minimizeobjective(obj, np.zeros(n), cfg, ls)
    # 1e-10 window: as in test_gpu_rosenbrock, the leading iterations on which two legitimate
    # reference summation orders (sequential vs compensated) agree with each other to 2.5e-11
    # (near convergence f = ½‖Ax − b‖² sits on its own rounding floor and is defined to ~1e-9 only)
    ocfg2, _, _ = make_pair("HagerZhang", sum_mode="comp", beta_form="literal", max_iters=400)
    ora2 = O.minimize(O.Objective.sparse_ls(n, 10, 8192, 24, 0), np.zeros(n), ocfg2)
    m = min(50, len(ora.trace_objective), len(ora2.trace_objective))
    w, nums = gate_numbers(ret, ora, ora2, m)
    record_gate("sparse least squares n=5e4 coh 0, HZ + StrongWolfe, device vs reference-shaped oracle", **nums)
    assert w >= 10
    np.testing.assert_allclose(ret.trace.objective[:w], ora.trace_objective[:w], rtol=1e-10)
    np.testing.assert_allclose(ret.trace.grad_norm[:w], ora.trace_grad_norm[:w], rtol=1e-10)
    np.testing.assert_allclose(ret.trace.objective[:m], ora.trace_objective[:m], rtol=1e-8)
    k = m
    assert np.array_equal(ret.trace.step_size[:k], ora.trace_step_size[:k])
    assert np.array_equal(ret.trace.objective_evals[:k], ora.trace_objective_evals[:k])
    assert ret.status == ora.status == "success"
    assert abs(ret.iters_ran - ora.iters_ran) <= 2
    assert abs(ret.objective - ora.objective) <= 1e-8 * max(abs(ora.objective), 1e-300) or ret.objective < 1e-9
    obj.close()


@pytest.mark.parametrize("linesearch", ["Wolfe", "YuanWeiLuWolfe", "Backtracking"])
def test_other_linesearches_bit_exact(ctx, linesearch):
    n = 6000
    ocfg, cfg, ls = make_pair("HagerZhang", linesearch, max_iters=60)
    obj = cg.SparseLSGPU(n, 10, 512, 24, 0, ctx)
    ora = O.minimize(O.Objective.sparse_ls(n, 10, 512, 24, 0), np.zeros(n), ocfg)
    ret = cg.minimizeobjective(obj, np.zeros(n), cfg, ls)
    assert_same_run(ret, ora, what=linesearch)
    obj.close()


def test_fused_direction_equals_unfused(ctx):
    n = 50_000
    _, cfg, ls = make_pair(max_iters=25)
    obj = cg.SparseLSGPU(n, 10, None, 24, 0, ctx)
    a = cg.minimizeobjective(obj, np.zeros(n), cfg, ls)
    b = cg.minimizeobjective(obj, np.zeros(n), cfg, ls, fuse_direction=False)
    assert np.array_equal(a.trace.objective, b.trace.objective)
    assert np.array_equal(a.minimizer, b.minimizer) and np.array_equal(a.gradient, b.gradient)
    obj.close()


@pytest.mark.parametrize("n", [20_000_000])
def test_large_n_properties(ctx, n):
    """Size-independent properties (same code path as n = 2e8, which bench.py runs):
    r(x_true) = 0 exactly (b was built by the same row sums) ⇒ f = 0, g = 0; f(0) = ½‖b‖²;
    ⟨Ax, y⟩ = ⟨x, Aᵀy⟩; CG decreases f monotonically and run-to-run identically."""
    obj = cg.SparseLSGPU(n, 10, None, 24, 0, ctx)
    xt = O.sparse_ls_xtrue(n, 24)
    ws = obj.make_workspace(xt, fuse_direction=False)
    assert ws.f_x0 == 0.0 and ws.norm_df_x0 == 0.0
    ws.close()
    rng = np.random.default_rng(3)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    Ax, ATy = obj.spmv(x), obj.spmv(y, transposed=True)
    assert abs(float(Ax @ y) - float(x @ ATy)) <= 1e-11 * np.linalg.norm(Ax) * np.linalg.norm(y)
    _, _, _, b = None, None, None, None
    ws = obj.make_workspace(np.zeros(n), fuse_direction=False)
    bb = obj.spmv(xt)                                              # = b
    assert abs(ws.f_x0 - 0.5 * float(bb @ bb)) <= 1e-12 * ws.f_x0
    ws.close()
    _, cfg, ls = make_pair(max_iters=6)
    r1 = cg.minimizeobjective(obj, np.zeros(n), cfg, ls)
    r2 = cg.minimizeobjective(obj, np.zeros(n), cfg, ls)
    assert np.all(np.diff(r1.trace.objective) < 0)
    assert np.array_equal(r1.trace.objective, r2.trace.objective) and np.array_equal(r1.minimizer, r2.minimizer)
    obj.close()


@pytest.mark.parametrize("coh", [0, 30])
def test_oracle_compared_run_n2e7(ctx, coh):
    """A run at a tenth of the full size (n = 2e7, 2e8 nonzeros: 264 tiles per persistent CTA,
    the dynamic slice queue, multi-tile reductions) against the oracle in the canonical order, bit for bit: five
    iterations, both matrix variants (coh 0: k_spmv_direct + BLAS-1 dots; coh 30: fused k_csr_rows)."""
    n = 20_000_000
    ocfg, cfg, ls = make_pair("HagerZhang", max_iters=5, sum_mode="cgo")
    obj = cg.SparseLSGPU(n, 10, None, 24, coh, ctx)
    ora_obj = O.Objective.sparse_ls(n, 10, None, 24, coh, threads=os.cpu_count() or 1)
    assert ora_obj.trial_site() == obj.trial_site
    ora = O.minimize(ora_obj, np.zeros(n), ocfg)
    ret = cg.minimizeobjective(obj, np.zeros(n), cfg, ls)
    assert ora.iters_ran == 5
    assert_same_run(ret, ora, what=f"n=2e7 coh={coh}")
    obj.close()


def test_full_size_properties_n2e8(ctx):
    """BASELINE.json configs[2] at its full size (n = 2e8, 2e9 nonzeros, 60 GB of matrices), through
    properties that need no oracle run: r(x_true) = 0 exactly (b was built by the same row sums) ⇒
    f = 0 and g = 0 bit for bit; f(0) > 0; CG decreases f monotonically; two runs are bit-identical."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a 180 GB B200")
    n = 200_000_000
    obj = cg.SparseLSGPU(n, 10, None, 24, 30, ctx)
    xt = O.sparse_ls_xtrue(n, 24)
    ws = obj.make_workspace(xt, fuse_direction=False)
    assert ws.f_x0 == 0.0 and ws.norm_df_x0 == 0.0
    ws.close()
    del xt
    _, cfg, ls = make_pair(max_iters=4)
    x0 = np.zeros(n)
    r1 = cg.minimizeobjective(obj, x0, cfg, ls)
    r2 = cg.minimizeobjective(obj, x0, cfg, ls)
    assert r1.trace.objective[0] > 0 and np.all(np.diff(r1.trace.objective) < 0)
    assert np.array_equal(r1.trace.objective, r2.trace.objective) and np.array_equal(r1.trace.step_size, r2.trace.step_size)
    assert np.array_equal(r1.minimizer, r2.minimizer)
    obj.close()


def test_hessian_vector_product_and_exact_line_minimum(ctx):
    """north_star's Hessian-vector product of the least-squares objective: hv = Aᵀ(A u) through the
    production SpMV kernels, against scipy; and what it is for — on a quadratic,
    ϕ(α) = ϕ₀ + α dϕ₀ + ½ α² u·Hu, so a* = −dϕ₀ / u·Hu zeroes the directional derivative."""
    import scipy.sparse as sp
    n = 30_000
    obj = cg.SparseLSGPU(n, 10, 512, 24, 0, ctx)
    rp, ci, va, b = obj.csr(False)
    A = sp.csr_matrix((va, ci, rp), shape=(n, n))
    x0 = np.random.default_rng(5).standard_normal(n)
    ws = obj.make_workspace(x0, fuse_direction=False)
    ws.reset_direction()                                   # u = −g
    g = ws.download()[1]
    uHu, hv = ws.hessvec_dir()
    ref = A.T @ (A @ (-g))
    np.testing.assert_allclose(hv, ref, rtol=1e-12, atol=1e-12 * np.max(np.abs(ref)))
    assert abs(uHu - (-g) @ ref) <= 1e-12 * abs(uHu)
    dphi0 = ws.dot_g_u()
    a_star = -dphi0 / uHu
    phi, dphi = ws.eval_trial(a_star)
    assert abs(dphi) <= 1e-9 * abs(dphi0)                  # exact minimiser along u
    assert abs(phi - (ws.f_x0 + a_star * dphi0 + 0.5 * a_star**2 * uHu)) <= 1e-11 * abs(ws.f_x0)
    ws.close()
    ro = cg.RosenbrockGPU(64, ctx)                         # objectives without one say so
    w2 = ro.make_workspace(np.zeros(64))
    with pytest.raises(cg.CgoError):
        w2.hessvec_dir()
    w2.close(); ro.close(); obj.close()


@pytest.mark.parametrize("flavour,linesearch", [("HagerZhang", "StrongWolfeBisection"), ("LBFGS", "Wolfe"),
                                                 ("LiuStorrey", "YuanWeiLuWolfe")])
def test_quadratic_aware_linesearch_matches_plain_path(ctx, flavour, linesearch):
    """with v = A u known, every trial of a line search is scalar arithmetic.  The
    decisions (step sizes, evaluation counts) must be those of the plain path, the objective within
    rounding of it, and each iteration must cost exactly one SpMV and one SpMVᵀ however many trials."""
    n = 40_000
    obj = cg.SparseLSGPU(n, 10, 1024, 24, 0, ctx)
    _, cfg, ls = make_pair(flavour, linesearch, max_iters=40, eps=1e-9)
    x0 = np.zeros(n)
    plain = cg.minimizeobjective(obj, x0, cfg, ls)
    ctx.timing(True)
    ctx.timing_read(reset=True)
    quad = cg.minimizeobjective(obj, x0, cfg, ls, quadratic_linesearch=True)
    t = ctx.timing_read(reset=True)
    ctx.timing(False)
    k = min(len(plain.trace.objective), len(quad.trace.objective), 30)
    assert k >= 10
    assert np.array_equal(quad.trace.step_size[:k], plain.trace.step_size[:k])
    assert np.array_equal(quad.trace.objective_evals[:k], plain.trace.objective_evals[:k])
    # r is carried forward as r + a v instead of being recomputed as A xp − b (the recursive residual of
    # every CG code): it drifts by ≈ ε‖r₀‖ per step, i.e. f = ½‖r‖² by ≈ ε √(f f₀), which only shows once f
    # has fallen by twenty orders of magnitude
    fp, fq, f0 = plain.trace.objective[:k], quad.trace.objective[:k], plain.trace.objective[0]
    assert np.all(np.abs(fq - fp) <= 1e-9 * fp + 1e-14 * np.sqrt(fp * f0))
    gp_, gq = plain.trace.grad_norm[:k], quad.trace.grad_norm[:k]
    assert np.all(np.abs(gq - gp_) <= 1e-7 * gp_ + 1e-13 * gp_[0])
    assert quad.status == plain.status and abs(quad.iters_ran - plain.iters_ran) <= 2
    iters, evals = quad.iters_ran, int(quad.trace.objective_evals.sum())
    assert evals > iters                                   # some line searches needed several trials …
    assert t["spmv"][1] <= iters + 2 and t["spmvT"][1] <= iters + 2      # … but each cost one SpMV + one SpMVᵀ
    obj.close()


@pytest.mark.parametrize("coh", [0, 30])
def test_quadratic_linesearch_survives_other_users_of_the_objective(ctx, coh):
    """The residual the quadratic-aware line search starts from belongs to one state; a Hessian-vector product or a
    second workspace on the same objective overwrites it.  The library notices and forms r = A x − b again."""
    n = 30_000
    obj = cg.SparseLSGPU(n, 10, 512, 24, coh, ctx)
    rng = np.random.default_rng(11)
    a = obj.make_workspace(rng.standard_normal(n), quadratic_linesearch=True)
    a.reset_direction()
    ref = a.eval_trial(1e-3)                      # (ϕ, dϕ) from r(x) left by the state's own first evaluation
    b = obj.make_workspace(rng.standard_normal(n), fuse_direction=False)     # another state on the same objective …
    b.reset_direction()
    b.hessvec_dir()                               # … and a Hessian-vector product: r is gone
    a.reset_direction()                           # a new line search of the first state
    again = a.eval_trial(1e-3)
    assert again == ref
    a.norm_df_xp()                                # materialise the step: r += a v, g⁺ = Aᵀ r
    b.close()
    fresh = obj.make_workspace(a.download_vector("xp"), fuse_direction=False)
    assert abs(fresh.f_x0 - a.pack[0]) <= 1e-12 * fresh.f_x0
    a.close(); fresh.close(); obj.close()


def test_quadratic_aware_linesearch_rejects_what_it_cannot_do(ctx):
    obj = cg.SparseLSGPU(2000, 10, 64, 24, 0, ctx)
    _, cfg, ls = make_pair("HagerZhang", "Backtracking")
    with pytest.raises(TypeError):
        cg.minimizeobjective(obj, np.zeros(2000), cfg, ls, quadratic_linesearch=True)
    ro = cg.RosenbrockGPU(64, ctx)
    _, cfg, ls = make_pair("HagerZhang")
    with pytest.raises(cg.CgoError):
        cg.minimizeobjective(ro, np.zeros(64), cfg, ls, quadratic_linesearch=True)
    obj.close(); ro.close()
