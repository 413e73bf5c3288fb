# minimizeobjective / minimizeobjectivererun for a device objective: the loop of
# src/engine/optim.jl:6-208 with every vector line replaced by one ccall (SURVEY.md §3.1).

# `x_initial` may be a `DeviceStart` (the iterate of a live workspace: no upload); `keep_workspace = ws -> …` is
# handed the run's workspace instead of it being closed, so that a restart can begin from it on the device.
function CGO.minimizeobjective(fdf!::DeviceObjective, x_initial::Union{Vector{Float64},DeviceStart},
                               config::CGConfig{Float64,BT,ET}, linesearch_config::LineSearchConfig;
                               β_literal::Bool = false, keep_workspace::Union{Nothing,Function} = nothing) where {BT,ET}
    max_iters, β_config = config.max_iters, config.β_config                   # :14-17
    lbfgs_m = β_config isa LBFGS ? β_config.m : 0
    info = DeviceWorkspace(fdf!, x_initial; lbfgs_m = lbfgs_m, β_literal = β_literal)   # :20-26, :45
    x, df_x = vec(info, :x), vec(info, :df_x)
    f_x, norm_df_x = info.f_x0, info.norm_df_x0
    norm_df_xp = NaN
    β = initializeβ(Float64, β_config)                                        # :29
    β isa Vector && (β = 0.0)
    fdf_evals_ran = -1
    f_x0 = f_x                                                                # :31
    ret = Results(f_x, Float64[], Float64[], 0, :incomplete, setuptrace(Float64, config.trace_status))  # :34-41
    resizetrace!(ret.trace, max_iters)                                        # :42
    initializeLineSearchContainer!(info, β_config, df_x, x)                   # :46
    a_initial = NaN                                                           # :47

    finish!(i, status) = begin                                                # updateresult!, types.jl:134-151
        ret.objective = f_x
        ret.minimizer, ret.gradient = download(info)
        ret.iters_ran, ret.status = i, status
        resizetrace!(ret.trace, i)
        keep_workspace === nothing ? close!(info) : keep_workspace(info)
        ret
    end

    for n = 1:max_iters                                                       # :50
        if isfinite(f_x) && isfinite(norm_df_x) && norm_df_x < config.ϵ       # :53-80
            return finish!(n - 1, f_x <= f_x0 ? :success : :increasing_objective)
        end
        f_xp, a_star, fdf_evals_ran, status = linesearch!(info, linesearch_config, fdf!, f_x, df_x, a_initial)  # :83
        a_initial = a_star                                                    # :92
        status == :success || return finish!(n - 1, status)                   # :93-104
        norm_df_xp = sqrt(info.pack[P_GPGP])                                  # :107 norm(info.df_xp)
        if !isfinite(f_xp) || !isfinite(norm_df_xp)                           # :108-121
            return finish!(n - 1, :non_finite_objective_or_gradient_proposed)
        end
        β = getβ(β_config, vec(info, :df_xp), df_x, vec(info, :u))            # :130-135
        accept!(info)                                  # :136-140  x ← xp, df_x ← df_xp: pointer swaps
        f_x, norm_df_x = f_xp, norm_df_xp                                     # :138, :141
        updatedir!(vec(info, :u), df_x, β)                                    # :145
        updatetrace!(ret.trace, f_x, norm_df_x, a_star, fdf_evals_ran, n)     # :152-159
    end
    return finish!(max_iters, :max_iters_reached)                             # :162-170
end

function CGO.minimizeobjectivererun(fdf!::DeviceObjective, x_initial::Vector{Float64},
                                    config::CGConfig{Float64,BT,ET}, linesearch_config::LineSearchConfig,
                                    rerun_config_tuples...) where {BT,ET}
    # the next attempt starts from rets[end].minimizer (:197): that vector is still on the device, in the previous
    # attempt's workspace — start from it there (one D2D copy) instead of uploading the host copy
    live = Any[nothing]
    keep = ws -> (live[1] === nothing || close!(live[1]); live[1] = ws)
    try
        rets = [minimizeobjective(fdf!, x_initial, config, linesearch_config; keep_workspace = keep)]   # :183-188
        for (rerun_config, backup_linesearch_config) in rerun_config_tuples   # :191
            rets[end].status == :success && break                             # :203
            previous = live[1]
            live[1] = nothing
            try
                push!(rets, minimizeobjective(fdf!, DeviceStart(previous), rerun_config, backup_linesearch_config;
                                              keep_workspace = keep))        # :195-200
            finally
                close!(previous)
            end
        end
        return rets
    finally
        live[1] === nothing || close!(live[1])
    end
end
