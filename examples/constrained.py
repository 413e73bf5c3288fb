"""The reference's examples/constrained.jl on the device: Booth under the box −10 < x < 10 by the primal barrier
method (src/engine/primal_barrier.jl), centering steps by Hager–Zhang CG with the (weak) Wolfe bisection, backed up
by the rerun chain of the example: (BroydenFamily(θ = 1) + Armijo backtracking), (Liu–Storey + Wolfe).

The reference's `hdh!` fills 2n dense constraint gradients (examples/constrained.jl:17-47); for the box they are
±e_d, so the device path takes the box itself (`BoxConstraint(lbs, ubs)`) and evaluates the barrier in one kernel.

Outcome on a B200: three centering steps succeed, each ≈10× closer to Booth's minimiser [1, 3] (the box is inactive
there); from t ≈ 2.5e5 on the example's ϵ = 1e-5 is below what its line searches resolve on t·f0 + ψ and the method
returns `centering_step_issue` at ‖x − [1, 3]‖ ≈ 6e-8 — the oracle's restatement of the reference stops the same way
a few steps later.  The reference's README.md:10 says as much about its own barrier method ("successively solves
harder and harder unconstrained optimization problems, to the point that linesearch failures are common due to finite
numerical precision"); what the Julia package prints for this script is not recorded anywhere in the reference.

    python examples/constrained.py      (needs a B200; there is no CPU fallback)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgoptim_b200 as cg  # noqa: E402
from min import boothfdf  # noqa: E402


def configs(N_vars=2):
    # ## Line search configurations.                                 examples/constrained.jl:64-104
    max_iters_ls, max_step_size, feasibility_max_iters = 100, 1e12, 50
    c1, c2 = 1e-3, 0.9
    linesearch_config_W = cg.WolfeBisection(cg.Wolfe(c1, c2), max_iters_ls, max_step_size, feasibility_max_iters)
    linesearch_config_A = cg.Backtracking(cg.Armijo(c1), 0.9, 300, feasibility_max_iters)
    # ## Conjugate gradient configurations.                          :106-147
    ϵ = 1e-5
    config_HZ = cg.setupCGConfig(ϵ, cg.HagerZhang(), cg.EnableTrace(), max_iters=1000)
    config_LS = cg.setupCGConfig(ϵ, cg.LiuStorrey(), cg.EnableTrace(), max_iters=1000)
    config_Broyden_DFP = cg.setupCGConfig(ϵ, cg.setupBroydenFamily(1.0, N_vars), cg.EnableTrace(), max_iters=1000)
    # Select line search and CG configurations.                      :151-158
    return (config_HZ, linesearch_config_W), [(config_Broyden_DFP, linesearch_config_A), (config_LS, linesearch_config_W)]


def main(verbose=True):
    ctx = cg.default_context()
    N_vars = 2
    fdf_ = cg.UserObjectiveGPU(N_vars, boothfdf, ctx)
    lbs, ubs = -10.0 * np.ones(N_vars), 10.0 * np.ones(N_vars)       # :51-52
    (config, linesearch_config), reruns = configs(N_vars)
    x0 = np.array([0.43, 1.23])                                      # :161
    constraints = cg.setupCvxInequalityConstraint(2 * N_vars, N_vars)            # :171
    barrier_config = cg.setupPrimalBarrierConfig(1e-8, 10.0, 100)                # :177-185
    b_ret = cg.primalbarriermethod_(constraints, fdf_, cg.BoxConstraint(lbs, ubs), x0, config, linesearch_config,
                                    barrier_config, *reruns)                     # :188-198
    ret = b_ret.centering_results[-1][-1]
    if verbose:                                                      # :201-216
        print("status", b_ret.status, "centering steps", b_ret.iters_ran, "t_final", b_ret.t_final,
              "objective evaluations", b_ret.total_objective_evals)
        print("last centering step:", ret.status, "minimizer", ret.minimizer, "objective", ret.objective,
              "norm(gradient)", np.linalg.norm(ret.gradient), "iterations", ret.iters_ran)
    fdf_.close()
    return b_ret


if __name__ == "__main__":
    main()
