#!/usr/bin/env python
"""Generates tests/golden/traces.json from the CPU oracle (oracle/cgo_oracle.c).

RESTATEMENT-DERIVED, NOT REFERENCE-EXECUTED: Julia is not installed in the build image, and the
reference's own test (test/runtests.jl:7-44) holds no solver vectors, so these
fixtures pin the oracle against silent drift between rounds; they cannot pin it against the
reference.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402

FLAVOURS = ["HagerZhang", "YuanWangSheng", "SallehAlhawarat", "LiuStorrey", "LBFGS"]
LINESEARCHES = ["StrongWolfeBisection", "Wolfe", "YuanWeiLuWolfe", "Backtracking"]
LS_DEFAULTS = {  # tests/helpers.make_pair
    "StrongWolfeBisection": dict(c1=1e-5, c2=0.8, ls_max_iters=1000),
    "Wolfe": dict(c1=1e-3, c2=0.9, ls_max_iters=100),
    "YuanWeiLuWolfe": dict(c1=1e-3, c2=0.9, ls_max_iters=100),
    "Backtracking": dict(c1=1e-3, c2=0.9, ls_max_iters=300),
}


def hexf(a):
    return [float(v).hex() for v in np.atleast_1d(a)]


def case(name, make_obj, x0, flavour, linesearch, sum_mode, beta_form, max_iters, keep=60):
    cfg = O.make_config(flavour, linesearch, max_iters=max_iters, sum_mode=sum_mode, beta_form=beta_form,
                        **LS_DEFAULTS[linesearch])
    r = O.minimize(make_obj(), x0, cfg)
    return {
        "name": name, "flavour": flavour, "linesearch": linesearch, "sum_mode": sum_mode,
        "beta_form": beta_form, "max_iters": max_iters, "x0": hexf(x0) if len(x0) <= 16 else None,
        "status": r.status, "iters_ran": int(r.iters_ran), "objective": float(r.objective).hex(),
        "fdf_evals_total": int(r.fdf_evals_total),
        "trace_objective": hexf(r.trace_objective[:keep]), "trace_grad_norm": hexf(r.trace_grad_norm[:keep]),
        "trace_step_size": hexf(r.trace_step_size[:keep]),
        "trace_objective_evals": [int(v) for v in r.trace_objective_evals[:keep]],
        "minimizer_head": hexf(r.minimizer[:8]),
    }


def main():
    out = []
    booth_x0 = np.array([0.43, 1.23])                                  # examples/min.jl:38
    for fl in FLAVOURS:
        for ls in LINESEARCHES:
            for sm, bf in (("seq", "literal"), ("cgo", "fused")):
                out.append(case("booth", O.Objective.booth, booth_x0, fl, ls, sm, bf, 1000))
    for n in (2, 10, 10_000):
        x0 = O.rosenbrock_x0(n, 24, 0.0)                               # SURVEY.md §8d cfg 1: (−1.2, 1, …)
        for sm, bf in (("seq", "literal"), ("cgo", "fused")):
            out.append(case(f"rosenbrock_n{n}", lambda n=n: O.Objective.rosenbrock(n), x0,
                            "HagerZhang", "StrongWolfeBisection", sm, bf, 1000))
    n = 2000
    for fl in ("HagerZhang", "LBFGS"):
        for sm, bf in (("seq", "literal"), ("cgo", "fused")):
            out.append(case("sparse_ls_n2000", lambda: O.Objective.sparse_ls(n, 10, 64, 24, 0), np.zeros(n),
                            fl, "StrongWolfeBisection", sm, bf, 300))
            out.append(case("logreg_3000x500", lambda: O.Objective.logreg(3000, 500, 20, 24, 1e-4),
                            np.zeros(500), fl, "StrongWolfeBisection", sm, bf, 80))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traces.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "derived_from": "oracle/cgo_oracle.c (restatement-derived, not reference-executed)",
                   "cases": out}, f, indent=0)
    print(f"wrote {len(out)} cases to {path} ({os.path.getsize(path) // 1024} KiB)")


def callers():
    """traces_callers.json: solvesystem (src/engine/solve_system.jl), primalbarriermethod!
    (src/engine/primal_barrier.jl) and the batched solver's reduction order."""
    out = []
    booth_x0 = np.array([0.43, 1.23])                                  # dev/solve_sys.jl:56
    for fl in ("HagerZhang", "YuanWangSheng", "SallehAlhawarat", "LiuStorrey"):
        for fix in (False, True):
            for sm, bf in (("seq", "literal"), ("cgo", "fused")):
                cfg = O.make_config(fl, max_iters=120, sum_mode=sm, beta_form=bf, mu=0.1)
                r = O.solvesystem(O.Objective.booth(), booth_x0, cfg, O.solvesys_ls(1.0, fix_stale_iterate=fix))
                out.append({"kind": "solvesystem", "name": "booth", "flavour": fl, "fix_stale_iterate": fix,
                            "sum_mode": sm, "beta_form": bf, "max_iters": 120, "status": r.status,
                            "iters_ran": int(r.iters_ran), "objective": float(r.objective).hex(),
                            "fdf_evals_total": int(r.fdf_evals_total),
                            "trace_objective": hexf(r.trace_objective[:60]), "trace_grad_norm": hexf(r.trace_grad_norm[:60]),
                            "trace_step_size": hexf(r.trace_step_size[:60]),
                            "trace_objective_evals": [int(v) for v in r.trace_objective_evals[:60]],
                            "minimizer": hexf(r.minimizer)})
    for update in (False, True):                                       # examples/constrained.jl:10-199
        cfgs = [O.make_config("HagerZhang", "Wolfe", c1=1e-3, c2=0.9, ls_max_iters=100, sum_mode="cgo", beta_form="fused"),
                O.make_config("LiuStorrey", "Backtracking", c1=1e-3, c2=0.9, ls_max_iters=300, sum_mode="cgo", beta_form="fused")]
        b = O.primalbarrier(O.Objective.booth(), [-10.0, -10.0], [10.0, 10.0], [0.43, 1.23], cfgs, 1e-8, 10.0, 100,
                            update_iterate=update)
        out.append({"kind": "primalbarrier", "name": "booth_box10", "update_iterate": update, "status": b.status,
                    "iters_ran": int(b.iters_ran), "t_final": float(b.t_final).hex(),
                    "total_objective_evals": int(b.total_objective_evals),
                    "attempts": [len(s) for s in b.centering_results],
                    "final_statuses": [s[-1].status for s in b.centering_results],
                    "final_iters": [int(s[-1].iters_ran) for s in b.centering_results],
                    "final_objectives": [float(s[-1].objective).hex() for s in b.centering_results]})
    n = 512                                                            # batched solver: one warp per problem
    x0 = O.rosenbrock_x0(n, 24, 0.1)
    O.set_cgo_batched(32)
    for ls in LINESEARCHES:
        cfg = O.make_config("HagerZhang", ls, max_iters=200, sum_mode="cgo", beta_form="fused", **LS_DEFAULTS[ls])
        r = O.minimize(O.Objective.rosenbrock(n), x0, cfg)
        out.append({"kind": "batched_order", "name": "rosenbrock_n512", "linesearch": ls, "lanes": 32, "status": r.status,
                    "iters_ran": int(r.iters_ran), "objective": float(r.objective).hex(),
                    "fdf_evals_total": int(r.fdf_evals_total), "trace_objective": hexf(r.trace_objective[:40])})
    O.set_cgo_lanes(256)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traces_callers.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py callers()",
                   "derived_from": "oracle/cgo_oracle.c, oracle/oracle.py (restatement-derived, not reference-executed)",
                   "cases": out}, f, indent=0)
    print(f"wrote {len(out)} cases to {path} ({os.path.getsize(path) // 1024} KiB)")


if __name__ == "__main__":
    main()
    callers()
