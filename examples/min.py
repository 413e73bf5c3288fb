"""The reference's examples/min.jl on the device: the Booth function (examples/helpers/test_funcs.jl:3-12) written
as a user `fdf!(g, x) -> f` on CUDA tensors, minimised with the configuration of examples/min.jl:16-43
(Hager–Zhang + StrongWolfeBisection(1e-5, 0.8), ϵ = 1e-5, x0 = [0.43, 1.23]); then the same call on a built-in
device objective of a size the GPU is for (extended Rosenbrock, n = 10⁷).

    python examples/min.py          (needs a B200; there is no CPU fallback)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgoptim_b200 as cg  # noqa: E402


def boothfdf(g, x):
    """boothfdf! — g and x are torch CUDA tensors over the solver's own device vectors"""
    a, b = x[0] + 2 * x[1] - 7, 2 * x[0] + x[1] - 5
    g[0] = 2 * a + 4 * b
    g[1] = 4 * a + 2 * b
    return a * a + b * b


def configs(flavour=None):
    # line search configuration.                                     examples/min.jl:15-21
    linesearch_config = cg.setupStrongWolfeBisection(1e-5, 0.8, a_max_growth_factor=2.0, max_iters=1000,
                                                     zoom_max_iters=100)
    # conjugate gradient configuration.                              :24-35
    config = cg.setupCGConfig(1e-5, flavour or cg.HagerZhang(), cg.EnableTrace(), max_iters=1000)
    return config, linesearch_config


def main(verbose=True, n_large=10_000_000):
    ctx = cg.default_context()
    config, linesearch_config = configs()
    fdf_ = cg.UserObjectiveGPU(2, boothfdf, ctx)
    x0 = np.array([0.43, 1.23])                                      # :38
    ret = cg.minimizeobjective(fdf_, x0, config, linesearch_config)  # :41-43
    if verbose:
        print("Results:")                                            # :46-49
        print("  minimizer", ret.minimizer, "objective", ret.objective, "norm(gradient)", np.linalg.norm(ret.gradient),
              "status", ret.status)
        print("  objective evaluations", int(ret.trace.objective_evals.sum()), "iterations", ret.iters_ran)
    fdf_.close()
    big = None
    if n_large:
        rosen = cg.RosenbrockGPU(n_large, ctx)
        big = cg.minimizeobjective(rosen, rosen.default_x0(24, 0.0), config, linesearch_config)
        if verbose:
            print(f"extended Rosenbrock n = {n_large}: status {big.status}, {big.iters_ran} iterations, "
                  f"f = {big.objective:.3e}, |x - 1|_inf = {np.abs(big.minimizer - 1.0).max():.2e}")
        rosen.close()
    return ret, big


if __name__ == "__main__":
    main()
