# Needs Julia, a B200 and libcgoptim.so (LIBCGOPTIM env var).  Mirrors examples/min.jl:16-43 and
# tests/test_gpu_rosenbrock.py: the reference on the CPU vs this host on the GPU, same objective.
using Test, LinearAlgebra
import ConjugateGradientOptim
using B200CGOptim

const CGO = ConjugateGradientOptim

# extended Rosenbrock as a host closure for the reference (expression order of the device kernel)
function rosenfdf!(g::Vector{Float64}, x::Vector{Float64})
    f = 0.0
    for p in 1:2:length(x)
        t = x[p+1] - x[p] * x[p]; om = 1.0 - x[p]
        f += (100.0 * t) * t + om * om
        g[p] = (-400.0 * x[p]) * t - 2.0 * om
        g[p+1] = 200.0 * t
    end
    return f
end

@testset "drop-in: reference (CPU) vs B200 host, extended Rosenbrock n = 10_000" begin
    n = 10_000
    x0 = repeat([-1.2, 1.0], n ÷ 2)
    ls = CGO.setupStrongWolfeBisection(1e-5, 0.8; a_max_growth_factor = 2.0, max_iters = 1000, zoom_max_iters = 100)
    cfg = CGO.setupCGConfig(1e-5, CGO.HagerZhang(), CGO.EnableTrace(); max_iters = 1000)
    ref = CGO.minimizeobjective(rosenfdf!, x0, cfg, ls)
    ctx = Context(0)
    ret = CGO.minimizeobjective(RosenbrockGPU(n, ctx), x0, cfg, ls)
    k = min(50, length(ref.trace.objective), length(ret.trace.objective))
    @test ret.status == ref.status == :success
    @test ret.trace.step_size[1:k] == ref.trace.step_size[1:k]                  # identical accept/reject decisions
    @test ret.trace.objective_evals[1:k] == ref.trace.objective_evals[1:k]
    @test all(isapprox.(ret.trace.objective[1:k], ref.trace.objective[1:k]; rtol = 1e-10))
    @test all(isapprox.(ret.trace.grad_norm[1:k], ref.trace.grad_norm[1:k]; rtol = 1e-10))
    @test isapprox(ret.objective, ref.objective; rtol = 1e-8, atol = 1e-20)
    @test abs(ret.iters_ran - ref.iters_ran) <= 2
end
