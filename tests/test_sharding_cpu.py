"""world_size-2 (and 3) `gloo` tests of the N>1 path's host logic, on CPU: contiguous shards
(cgo_shard_range through the C ABI, host-only), SPMD lock-step of the engine / line-search state
machines on replicated scalars, and the rank-ordered combination of per-shard canonical sums
(include/cgoptim.h).  The vector work is done by the test-only numpy workspace on each rank's
slice; the result must equal the oracle run on the whole problem with the same shard count."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cgoptim_b200 as cg
    from oracle import oracle as O
    from helpers import make_pair
    from numpy_workspace import NumpyObjective, NumpyWorkspace

    class ShardedWorkspace(NumpyWorkspace):
        """Per-rank slice; every reduction = canonical sum of the slice, all-gathered, added in
        rank order on every rank (what cgo_finish_pack does with NCCL)."""

        def _combine(self, local):
            buf = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(buf, torch.tensor([local], dtype=torch.float64))
            s = buf[0].item()
            for r in range(1, world):
                s = s + buf[r].item()
            return s

        def _bdot(self, a, b):
            return self._combine(super()._bdot(a, b))

        def _tdot(self, a, b):
            return self._combine(super()._tdot(a, b))

    class ShardedObjective(NumpyObjective):
        def make_workspace(self, x_initial, lbfgs_m=0, fuse_direction=True, beta_form="fused"):
            return ShardedWorkspace(self, x_initial, lbfgs_m, fuse_direction, beta_form)

    class LocalRosenbrock:
        """fdf of the slice; f is combined across ranks in rank order"""

        def __init__(self, n_local):
            self.o = O.Objective.rosenbrock(n_local)
            self.n = n_local

        def set_sum_mode(self, m, threads=0):
            self.o.set_sum_mode(m, threads)

        def trial_site(self):
            return self.o.trial_site()

        def fdf(self, x):
            f, g = self.o.fdf(x)
            buf = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(buf, torch.tensor([f], dtype=torch.float64))
            s = buf[0].item()
            for r in range(1, world):
                s = s + buf[r].item()
            return s, g

    ok, msgs = True, []
    n = 20_000 + 2 * rank * 0           # same n on every rank
    for flavour, linesearch in (("HagerZhang", "StrongWolfeBisection"), ("LBFGS", "StrongWolfeBisection"),
                                ("YuanWangSheng", "Wolfe"), ("LiuStorrey", "Backtracking")):
        lo, hi = cg.shard_range(n, world, rank, 2)
        ocfg, cfg, ls = make_pair(flavour, linesearch, max_iters=40)
        x0 = O.rosenbrock_x0(n, 24, 0.1)
        O.set_cgo_order(296, 1)          # each rank reduces its own slice as ONE shard
        ret = cg.minimizeobjective(ShardedObjective(LocalRosenbrock(hi - lo)), x0[lo:hi], cfg, ls)
        O.set_cgo_order(296, world)      # the oracle sees the whole problem, `world` shards
        ora = O.minimize(O.Objective.rosenbrock(n), x0, ocfg)
        same = (ret.status == ora.status and ret.iters_ran == ora.iters_ran
                and np.array_equal(ret.trace.objective, ora.trace_objective)
                and np.array_equal(ret.trace.grad_norm, ora.trace_grad_norm)
                and np.array_equal(ret.trace.step_size, ora.trace_step_size)
                and np.array_equal(ret.minimizer, ora.minimizer[lo:hi]))
        if not same:
            ok = False
            msgs.append(f"{flavour}/{linesearch} rank {rank}: {ret.status}/{ora.status} {ret.iters_ran}/{ora.iters_ran}")
    # the callers next to the hot path, sharded the same way: solvesystem (as written and as published)
    n = 3000
    for fix in (False, True):
        lo, hi = cg.shard_range(n, world, rank, 2)
        ocfg, cfg, _ = make_pair("YuanWangSheng", max_iters=4)
        x0 = O.rosenbrock_x0(n, 24, 0.1)
        O.set_cgo_order(296, 1)
        ret = cg.solvesystem(ShardedObjective(LocalRosenbrock(hi - lo)), x0[lo:hi], cfg,
                             cg.setupLinesearchSolveSys(1.0), fix_stale_iterate=fix)
        O.set_cgo_order(296, world)
        ora = O.solvesystem(O.Objective.rosenbrock(n), x0, ocfg, O.solvesys_ls(1.0, fix_stale_iterate=fix))
        same = (ret.status == ora.status and ret.iters_ran == ora.iters_ran
                and np.array_equal(ret.trace.objective, ora.trace_objective)
                and np.array_equal(ret.trace.objective_evals, ora.trace_objective_evals)
                and np.array_equal(ret.minimizer, ora.minimizer[lo:hi]))
        if not same:
            ok = False
            msgs.append(f"solvesystem fix={fix} rank {rank}: {ret.status}/{ora.status} {ret.iters_ran}/{ora.iters_ran}")
    # shard ranges tile [0, n) with even boundaries
    bounds = [cg.shard_range(1_000_006, world, r, 2) for r in range(world)]
    ok = ok and bounds[0][0] == 0 and bounds[-1][1] == 1_000_006
    ok = ok and all(bounds[r][1] == bounds[r + 1][0] and bounds[r][1] % 2 == 0 for r in range(world - 1))
    q.put((rank, ok, msgs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_host_logic_matches_whole_problem_oracle(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msgs in res:
        assert ok, msgs
