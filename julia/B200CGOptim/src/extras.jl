# Callers next to the hot path and the batched solver (SURVEY.md §8f N2-N4) over the C ABI.
# Unexecuted here (Julia is not in the build image); the tested twin is
# conjugategradientoptim.jl_b200/engine/{solve_system,primal_barrier}.py and device.py.

# ---------------------------------------------------------------- solvesystem (src/engine/solve_system.jl)
solvesys_begin!(ws::DeviceWorkspace) = @cgo(cgo_solvesys_begin, (Ptr{Cvoid},), ws.h)
"updateiteratesolvesys! (:237-253) + f_x_next = fdf!(info.df_xp, x_next) (:179) -> (f_x_next, norm(info.df_xp))"
function solvesys_project!(ws::DeviceWorkspace, m::Float64; fix_stale_iterate::Bool = false)
    materializedirection!(ws)
    @cgo(cgo_solvesys_project, (Ptr{Cvoid}, Float64, Int32, Ptr{Float64}), ws.h, m, Int32(fix_stale_iterate), ws.buf)
    ws.pack = copy(ws.buf)
    return ws.pack[P_PHI], sqrt(ws.pack[P_GPGP])
end
solvesys_accept!(ws::DeviceWorkspace; fix_stale_iterate::Bool = false) =
    @cgo(cgo_solvesys_accept, (Ptr{Cvoid}, Int32), ws.h, Int32(fix_stale_iterate))

# linesearch! (solve_system.jl:29-55) on the device container: same loop, evalϕdϕ! is one launch
function CGO.linesearch!(info::DeviceWorkspace, config::CGO.LinesearchSolveSys{Float64}, fdf!::DeviceObjective)
    info.hint = config.s                                                  # the first trial is always a = s·ρ⁰
    norm_u_sq = dot_u_u(info)                                             # :39
    for i = 0:config.max_iters-1                                          # :41
        a = config.s * config.ρ^i                                         # :42
        f_xp, dϕ_xp = evaltrial!(info, a)                                 # :44
        norm_df_xp = sqrt(info.pack[P_GPGP])                              # :47
        if !(-dϕ_xp < config.σ * a * norm_df_xp * norm_u_sq)              # :48
            return f_xp, norm_df_xp, a, i, true                           # :50
        end
    end
    return NaN, NaN, NaN, config.max_iters - 1, false                     # (:54 is an UndefVarError in the reference)
end

"info.xp, info.df_xp on the host (updateresult! with the container, types.jl:116-131)"
function download_trial(ws::DeviceWorkspace)
    x, g = Vector{Float64}(undef, ws.n), Vector{Float64}(undef, ws.n)
    check(ccall((:cgo_download_vector, LIBCGOPTIM[]), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), ws.h, 3, x))
    check(ccall((:cgo_download_vector, LIBCGOPTIM[]), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), ws.h, 4, g))
    return x, g
end

"""
solvesystem (src/engine/solve_system.jl:64-239) for a device objective: the reference's loop with
its vector lines replaced by the three calls above; `fix_stale_iterate = true` projects from the
current iterate (Alg. 3.1 as published) instead of from `x_next` (:171-177 as written).
"""
function CGO.solvesystem(fdf!::DeviceObjective, x_initial::Vector{Float64}, config::CGConfig{Float64,BT,ET},
                         linesearch_config::CGO.LinesearchSolveSys{Float64};
                         fix_stale_iterate::Bool = false) where {BT<:CGβConfig,ET}
    max_iters, β_config = config.max_iters, config.β_config               # :74-75
    info = DeviceWorkspace(fdf!, x_initial)                               # :80-87
    solvesys_begin!(info)                                                 # :82
    x, df_x = vec(info, :x), vec(info, :df_x)
    f_x, norm_df_x = info.f_x0, info.norm_df_x0
    ret = Results(f_x, Float64[], Float64[], 0, :incomplete, setuptrace(Float64, config.trace_status))   # :93-101
    resizetrace!(ret.trace, max_iters)
    initializeLineSearchContainer!(info, β_config, df_x, x)               # :104-105

    finish!(f, i, status; from_trial = false) = begin                     # updateresult!, types.jl:116-151
        ret.objective = f
        ret.minimizer, ret.gradient = from_trial ? download_trial(info) : download(info)
        ret.iters_ran, ret.status = i, status
        from_trial || resizetrace!(ret.trace, i)
        close!(info)
        ret
    end

    for n = 1:max_iters                                                   # :109
        norm_df_x < config.ϵ && return finish!(f_x, n - 1, :success)      # :112-123
        f_xp, norm_df_xp, a_star, evals, ok = linesearch!(info, linesearch_config, fdf!)             # :126-130
        ok || return finish!(f_x, n - 1, :linesearch_failed)              # :131-142
        if norm_df_xp < config.ϵ                                          # :146-168
            resizetrace!(ret.trace, n)
            updatetrace!(ret.trace, f_xp, norm_df_xp, a_star, evals, n)
            return finish!(f_xp, n, :success; from_trial = true)
        end
        m = a_star * info.pack[P_DPHI] / norm_df_xp^2                     # :246
        f_x_next, norm_next = solvesys_project!(info, m; fix_stale_iterate = fix_stale_iterate)      # :171-179
        (isfinite(f_x_next) && isfinite(norm_next)) ||
            return finish!(f_x, n - 1, :non_finite_objective_or_gradient_proposed)                   # :180-194
        β = getβ(β_config, vec(info, :df_xp), df_x, vec(info, :u))        # :201-206
        solvesys_accept!(info; fix_stale_iterate = fix_stale_iterate)     # :196, :207-208
        f_x, norm_df_x = f_x_next, norm_next                              # :197, :209
        updatedir!(vec(info, :u), df_x, β)                                # :212
        updatetrace!(ret.trace, f_x, norm_df_x, a_star, evals, n)         # :215-222
    end
    return finish!(f_x, max_iters, :max_iters_reached)                    # :225-233
end

# ---------------------------------------------------------------- primal barrier (src/engine/primal_barrier.jl)
"The `hdh!` of examples/constrained.jl:17-47 as data: fi = [x − ubs; lbs − x] (this rank's shard)."
struct BoxConstraint
    lbs::Vector{Float64}
    ubs::Vector{Float64}
end
"t·f0(x) − Σ log(ubs − x) − Σ log(x − lbs) around a device objective (evalbarrier!, :112-133)."
function BoxBarrierGPU(inner::DeviceObjective, box::BoxConstraint, t::Float64 = 1.0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:cgo_obj_box_barrier_create, LIBCGOPTIM[]), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Ref{Ptr{Cvoid}}), inner.ctx.h, inner.h, box.lbs, box.ubs, t, h))
    return _wrap(inner.ctx, h[])          # keep `inner` reachable for as long as the barrier objective lives
end
set_t!(bar::DeviceObjective, t::Float64) = @cgo(cgo_obj_barrier_set_t, (Ptr{Cvoid}, Float64), bar.h, t)
function infeasible_count(bar::DeviceObjective, x::Vector{Float64})
    c = Ref{Int64}(0)
    check(ccall((:cgo_obj_barrier_infeasible, LIBCGOPTIM[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}), bar.h, x, c))
    return c[]
end

"primalbarriermethod! (:158-255) with a device objective and box constraints; outer loop as written."
function CGO.primalbarriermethod!(constraints::CGO.CvxInequalityConstraint{Float64}, f0df0!::DeviceObjective,
                                  hdh!::BoxConstraint, x_initial::Vector{Float64}, centering_config::CGO.CGConfig,
                                  linesearch_config::CGO.LineSearchConfig, barrier_config::CGO.PrimalBarrierConfig{Float64},
                                  rerun_config_tuples...; update_iterate::Bool = false)
    x = copy(x_initial)                                                   # :179
    rets = Vector{Any}(undef, barrier_config.max_iters)                   # :184
    fdf! = BoxBarrierGPU(f0df0!, hdh!, 1.0)                               # :205-213
    infeasible_count(fdf!, x) > 0 &&
        return CGO.assembleresults!(rets, :infeasible_start, 0, barrier_config.t_initial)           # :187-198
    t = barrier_config.t_initial                                          # verifyt0 :259-277
    if !isfinite(t) || t < 0
        ws = DeviceWorkspace(f0df0!, x_initial); t = (ws.f_x0 - barrier_config.inf_f0_lb) * barrier_config.barrier_growth_factor
    end
    for i = 1:barrier_config.max_iters                                    # :215
        set_t!(fdf!, t)
        rets[i] = CGO.minimizeobjectivererun(fdf!, x, centering_config, linesearch_config, rerun_config_tuples...)   # :217-223
        rets[i][end].status != :success && return CGO.assembleresults!(rets, :centering_step_issue, i, t)           # :224-232
        CGO.getNconstraints(constraints) / t < barrier_config.barrier_tol &&
            return CGO.assembleresults!(rets, :success, i, t)                                                       # :235-243
        update_iterate && (x = copy(rets[i][end].minimizer))
        t = barrier_config.barrier_growth_factor * t                      # :246
    end
    return CGO.assembleresults!(rets, :max_iters_reached, barrier_config.max_iters, t)              # :249-254
end

# ---------------------------------------------------------------- batched solver (SURVEY.md §8 cfg 5)
struct BatchedConfig                       # cgo_batched_config, include/cgoptim.h
    eps::Float64; max_iters::Int64; flavour::Int32; linesearch::Int32
    mu::Float64; c1::Float64; c2::Float64; growth::Float64
    ls_max_iters::Int64; zoom_max_iters::Int64
    delta1::Float64; max_step_size::Float64; discount::Float64; feas_max_iters::Int64
end
"minimizeobjective for many independent extended-Rosenbrock problems: X0 is n × nprob (column per problem)."
function minimizeobjective_batched(ctx::Context, X0::Matrix{Float64}, bc::BatchedConfig)
    n, nprob = size(X0)
    obj, gn = zeros(nprob), zeros(nprob)
    it, ev, st = zeros(Int64, nprob), zeros(Int64, nprob), zeros(Int32, nprob)
    xm = similar(X0)
    check(ccall((:cgo_batched_minimize_rosenbrock, LIBCGOPTIM[]), Cint,
        (Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Ref{BatchedConfig}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}, Ptr{Int64},
         Ptr{Float64}, Ptr{Float64}), ctx.h, nprob, n, X0, bc, obj, it, st, ev, xm, gn))
    return (objective = obj, minimizer = xm, grad_norm = gn, iters_ran = it, status = st, fdf_evals = ev)
end
