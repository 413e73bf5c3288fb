"""Worker of tests/test_gpu_multirank.py: one process per GPU (torchrun), rows / vector slices
sharded over the ranks, compared bit for bit with the oracle run on the WHOLE problem in
canonical order with the same shard count (include/cgoptim.h: every rank reduces its shard, the
shard results are added in rank order)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cgoptim_b200 as cg  # noqa: E402
from oracle import oracle as O  # noqa: E402
from helpers import make_pair  # noqa: E402


def sources_sha():
    """what a log of this worker vouches for: the worker itself and every CUDA source of libcgoptim.so"""
    import glob
    import hashlib
    h = hashlib.sha1()
    for f in [os.path.abspath(__file__)] + sorted(glob.glob(os.path.join(ROOT, "conjugategradientoptim.jl_b200", "csrc", "*.cu*"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = cg.Context(local)
    ctx.comm_init_torch()
    O.set_cgo_order(296, world)
    fails = []
    if rank == 0:
        print(f"PEER_MEMORY={int(ctx.peer_memory)}", flush=True)
        print(f"WORLD={world} SOURCES_SHA={sources_sha()}", flush=True)

    def check(name, obj, x0_full, ora_obj, flavour, max_iters):
        ocfg, cfg, ls = make_pair(flavour, max_iters=max_iters)
        lo, hi = obj.offset, obj.offset + obj.n_local
        assert (lo, hi) == cg.shard_range(obj.n_global, world, rank, 2)
        ret = cg.minimizeobjective(obj, x0_full[lo:hi], cfg, ls)
        ora = O.minimize(ora_obj, x0_full, ocfg)
        ok = (ret.status == ora.status and ret.iters_ran == ora.iters_ran
              and np.array_equal(ret.trace.objective, ora.trace_objective)
              and np.array_equal(ret.trace.grad_norm, ora.trace_grad_norm)
              and np.array_equal(ret.trace.step_size, ora.trace_step_size)
              and np.array_equal(ret.trace.objective_evals, ora.trace_objective_evals)
              and np.array_equal(ret.minimizer, ora.minimizer[lo:hi])
              and np.array_equal(ret.gradient, ora.gradient[lo:hi]))
        if not ok:
            fails.append(f"{name}/{flavour}: rank {rank} status {ret.status}/{ora.status} iters {ret.iters_ran}/{ora.iters_ran} "
                         f"f {ret.objective!r}/{ora.objective!r}")

    n = 40_000
    for flavour in ("HagerZhang", "LBFGS"):
        obj = cg.RosenbrockGPU(n, ctx)
        check("rosenbrock", obj, O.rosenbrock_x0(n, 24, 0.1), O.Objective.rosenbrock(n), flavour, 40)
        obj.close()
    # the reference's own chained Rosenbrock (examples/helpers/test_funcs.jl:50-57): ±1 halo of the trial point
    for flavour in ("HagerZhang", "LBFGS"):
        obj = cg.RosenbrockChainedGPU(n, ctx)
        check("rosenbrock_chained", obj, O.rosenbrock_x0(n, 24, 0.1), O.Objective.rosenbrock_chained(n), flavour, 40)
        obj.close()
    for coh in (0, 30):
        for flavour in ("HagerZhang", "LBFGS"):
            obj = cg.SparseLSGPU(n, 10, 2048, 24, coh, ctx)
            check(f"sparse_ls coh={coh}", obj, np.zeros(n), O.Objective.sparse_ls(n, 10, 2048, 24, coh), flavour, 60)
            obj.close()
    # SURVEY.md §8f N1 and north_star's Hessian-vector product, sharded: hv = Aᵀ(A u) and u·Hu against scipy on the
    # oracle's global CSR; the quadratic-aware line search must take the plain path's decisions on the same ranks
    import scipy.sparse as sp
    for coh in (0, 30):
        obj = cg.SparseLSGPU(n, 10, 2048, 24, coh, ctx)
        lo, hi = obj.offset, obj.offset + obj.n_local
        ora_obj = O.Objective.sparse_ls(n, 10, 2048, 24, coh)
        rp, ci, va = ora_obj.csr(False)
        A = sp.csr_matrix((va, ci, rp), shape=(n, n))
        xr = np.random.default_rng(5).standard_normal(n)
        ws = obj.make_workspace(xr[lo:hi], fuse_direction=False)
        ws.reset_direction()                               # u = −g
        g_loc = ws.download()[1]
        uHu, hv = ws.hessvec_dir()
        g_full = A.T @ (A @ xr - ora_obj.rhs())
        ref = A.T @ (A @ (-g_full))
        ok = (np.allclose(g_loc, g_full[lo:hi], rtol=1e-12, atol=1e-12 * np.max(np.abs(g_full)))
              and np.allclose(hv, ref[lo:hi], rtol=1e-12, atol=1e-12 * np.max(np.abs(ref)))
              and abs(uHu - (-g_full) @ ref) <= 1e-12 * abs(uHu))
        ws.close()
        if not ok:
            fails.append(f"hessvec coh={coh}: rank {rank} uHu {uHu!r} vs {(-g_full) @ ref!r}")
        _, cfg, ls = make_pair("HagerZhang", max_iters=40, eps=1e-9)
        plain = cg.minimizeobjective(obj, np.zeros(hi - lo), cfg, ls)
        quad = cg.minimizeobjective(obj, np.zeros(hi - lo), cfg, ls, quadratic_linesearch=True)
        k = min(len(plain.trace.objective), len(quad.trace.objective), 30)
        fp, fq, f0 = plain.trace.objective[:k], quad.trace.objective[:k], plain.trace.objective[0]
        ok = (k >= 10 and np.array_equal(quad.trace.step_size[:k], plain.trace.step_size[:k])
              and np.array_equal(quad.trace.objective_evals[:k], plain.trace.objective_evals[:k])
              and bool(np.all(np.abs(fq - fp) <= 1e-9 * fp + 1e-14 * np.sqrt(fp * f0)))
              and quad.status == plain.status and abs(quad.iters_ran - plain.iters_ran) <= 2)
        if not ok:
            fails.append(f"quadratic line search coh={coh}: rank {rank} {quad.status}/{plain.status} {quad.iters_ran}/{plain.iters_ran} k={k}")
        obj.close()
    # the callers next to the hot path, sharded: solvesystem (src/engine/solve_system.jl) bit for bit,
    # the box log barrier (src/engine/primal_barrier.jl) at the libm tolerance
    n = 40_000
    for fix in (False, True):
        ocfg, cfg, _ = make_pair("YuanWangSheng", max_iters=8 if not fix else 20)
        obj = cg.SparseLSGPU(n, 10, 2048, 24, 0, ctx)
        lo, hi = obj.offset, obj.offset + obj.n_local
        ret = cg.solvesystem(obj, np.zeros(hi - lo), cfg, cg.setupLinesearchSolveSys(1.0), fix_stale_iterate=fix)
        ora = O.solvesystem(O.Objective.sparse_ls(n, 10, 2048, 24, 0), np.zeros(n), ocfg,
                            O.solvesys_ls(1.0, fix_stale_iterate=fix))
        ok = (ret.status == ora.status and ret.iters_ran == ora.iters_ran
              and np.array_equal(ret.trace.objective, ora.trace_objective)
              and np.array_equal(ret.trace.grad_norm, ora.trace_grad_norm)
              and np.array_equal(ret.trace.objective_evals, ora.trace_objective_evals)
              and np.array_equal(ret.minimizer, ora.minimizer[lo:hi]))
        if not ok:
            fails.append(f"solvesystem fix={fix}: rank {rank} {ret.status}/{ora.status} {ret.iters_ran}/{ora.iters_ran}")
        obj.close()
    lbs, ubs = -2.0 * np.ones(n), 0.8 * np.ones(n)
    ocfg, cfg, ls = make_pair("HagerZhang", max_iters=30, eps=1e-6)
    f0 = cg.RosenbrockGPU(n, ctx)
    lo, hi = f0.offset, f0.offset + f0.n_local
    bar = cg.BoxBarrierGPU(f0, lbs[lo:hi], ubs[lo:hi], 5.0)
    assert bar.infeasible_count(np.zeros(hi - lo)) == 0 and bar.infeasible_count(np.ones(hi - lo)) == n
    ret = cg.minimizeobjective(bar, np.zeros(hi - lo), cfg, ls)
    ora = O.minimize(O.Objective.box_barrier(O.Objective.rosenbrock(n), lbs, ubs, 5.0), np.zeros(n), ocfg)
    k = min(25, len(ora.trace_objective), len(ret.trace.objective))
    ok = (k >= 5 and np.array_equal(ret.trace.step_size[:k], ora.trace_step_size[:k])
          and np.allclose(ret.trace.objective[:k], ora.trace_objective[:k], rtol=1e-10, atol=0)
          and np.allclose(ret.minimizer, ora.minimizer[lo:hi], rtol=1e-7, atol=1e-9))
    if not ok:
        fails.append(f"barrier: rank {rank} {ret.status}/{ora.status} {ret.iters_ran}/{ora.iters_ran} k={k}")
    bar.close(); f0.close()

    # cfg 4 (BASELINE.json configs[3]): sample-sharded logistic regression + L-BFGS.  exp / log1p
    # differ from glibc in the last ulp and the gradient partials are added shard by shard, so the
    # comparison is at north_star's tolerances with identical line-search decisions.
    N, d, lam = 20_000, 2002, 1e-4
    for flavour in ("LBFGS", "HagerZhang"):
        ocfg, cfg, ls = make_pair(flavour, eps=1e-6, max_iters=60, c1=1e-4, c2=0.9, lbfgs_m=10)
        ctx.set_gather_block_bytes(8 * 600 if flavour == "HagerZhang" else 40 << 20)   # also the column-blocked passes
        obj = cg.LogRegGPU(N, d, 20, 24, lam, ctx)
        assert (obj.csr_blocks(False) > 1) == (flavour == "HagerZhang")
        lo, hi = obj.offset, obj.offset + obj.n_local
        assert (lo, hi) == cg.shard_range(d, world, rank, 2) and obj.n_global == d
        ret = cg.minimizeobjective(obj, np.zeros(hi - lo), cfg, ls)
        ora_obj = O.Objective.logreg(N, d, 20, 24, lam)
        ora_obj.set_trial_site(*obj.trial_site)
        ora = O.minimize(ora_obj, np.zeros(d), ocfg)
        k = min(50, len(ora.trace_objective), len(ret.trace.objective))
        ok = (k >= 10 and ret.status == ora.status and abs(ret.iters_ran - ora.iters_ran) <= 2
              and np.array_equal(ret.trace.step_size[:k], ora.trace_step_size[:k])
              and np.array_equal(ret.trace.objective_evals[:k], ora.trace_objective_evals[:k])
              and np.allclose(ret.trace.objective[:k], ora.trace_objective[:k], rtol=1e-10, atol=0)
              and np.allclose(ret.trace.grad_norm[:k], ora.trace_grad_norm[:k], rtol=1e-8, atol=0)
              and abs(ret.objective - ora.objective) <= 1e-8 * abs(ora.objective)
              and np.allclose(ret.minimizer, ora.minimizer[lo:hi], rtol=1e-6, atol=1e-9))
        if not ok:
            fails.append(f"logreg/{flavour}: rank {rank} status {ret.status}/{ora.status} iters {ret.iters_ran}/{ora.iters_ran} "
                         f"f {ret.objective!r}/{ora.objective!r} k={k}")
        obj.close()
    ctx.barrier()
    t = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(t)
    for f in fails:
        print("MISMATCH", f, flush=True)
    if rank == 0:
        print("MULTIRANK_OK" if int(t.item()) == 0 else f"MULTIRANK_FAILED {int(t.item())}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
