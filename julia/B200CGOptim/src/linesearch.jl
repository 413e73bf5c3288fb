# Line-search state machines on the device workspace: methods added to the reference's
# `linesearch!` / `evalϕdϕ!`.  Scalar logic restated from src/linesearch/{nocedal,wolfe,geometric}.jl
# (including the quirks listed in SURVEY.md §8a); each evalϕdϕ! is ONE fused kernel launch.

# evalϕdϕ! (src/cg_utils.jl:3-22)
CGO.evalϕdϕ!(xp::DeviceVector, df_xp::DeviceVector, fdf!::DeviceObjective, a::Float64, x::DeviceVector, u::DeviceVector) =
    evaltrial!(xp.ws, a)

_vecs(info::DeviceWorkspace) = (vec(info, :xp), vec(info, :df_xp), vec(info, :x), vec(info, :u))

# ---------------- StrongWolfeBisection (nocedal.jl:33-158)
function CGO.linesearch!(info::DeviceWorkspace, config::StrongWolfeBisection{Float64}, fdf!::DeviceObjective,
                         f_x::Float64, df_x::DeviceVector, a_initial::Float64)
    c1, c2, growth = config.c1, config.c2, config.a_max_growth_factor
    xp, df_xp, x, u = _vecs(info)
    (0.0 < a_initial && isfinite(a_initial)) || (a_initial = 1.0)            # :49-52
    ϕ_0 = f_x
    info.hint = a_initial                       # lets the deferred updatedir! ride on the first trial
    dϕ_0 = devdot(df_x, u)                                                    # :56
    dϕ_0 > 0.0 && return ϕ_0, 0.0, 0, :non_descent_search_direction           # :57-63
    a_prev, ϕ_a_prev = 0.0, ϕ_0
    a, ϕ_a, dϕ_a = a_initial, ϕ_0, dϕ_0
    a_max = a * growth
    evals = 0
    non_initial = false
    for _ = 1:config.max_iters                                                # :76
        ϕ_a, dϕ_a = evalϕdϕ!(xp, df_xp, fdf!, a, x, u)                        # :78
        evals += 1
        chk1 = ϕ_a > ϕ_0 + c1 * a * dϕ_0                                      # :81
        chk2 = ϕ_a >= ϕ_a_prev                                                # :82
        if chk1 || (chk2 && non_initial)                                      # :83-105
            return zoom!(info, fdf!, a_prev, a, ϕ_a_prev, ϕ_0, dϕ_0, c1, c2, evals, config.zoom_max_iters)
        end
        abs(dϕ_a) <= -c2 * dϕ_0 && return ϕ_a, a, evals, :success            # :107-110
        if dϕ_a >= 0                                                          # :112-134
            return zoom!(info, fdf!, a, a_prev, ϕ_a, ϕ_0, dϕ_0, c1, c2, evals, config.zoom_max_iters)
        end
        a_prev, ϕ_a_prev, non_initial = a, ϕ_a, true                          # :137-139
        a_max = a * growth                                                    # :143
        a > a_max && return ϕ_a, a, evals, :linesearch_a_max_overflow         # :144-149
        a = (a_max + a) / 2                                                   # :150
    end
    return ϕ_a, a, evals, :linesearch_max_iters_reached                       # :157
end
# zoom! (nocedal.jl:162-209)
function zoom!(info::DeviceWorkspace, fdf!, a_lb, a_ub, ϕ_a_lb, ϕ_0, dϕ_0, c1, c2, evals, max_iters)
    xp, df_xp, x, u = _vecs(info)
    a, ϕ_a, dϕ_a = 0.0, 0.0, 0.0
    for _ = 1:max_iters
        a = (a_lb + a_ub) / 2                                                 # :187
        ϕ_a, dϕ_a = evalϕdϕ!(xp, df_xp, fdf!, a, x, u)                        # :190
        evals += 1
        if (ϕ_a > ϕ_0 + c1 * a * dϕ_0) || (ϕ_a >= ϕ_a_lb)                     # :193
            a_ub = a
        else
            abs(dϕ_a) <= -c2 * dϕ_0 && return ϕ_a, a, evals, :success        # :196
            dϕ_a * (a_ub - a_lb) >= 0 && (a_ub = a_lb)                        # :200
            a_lb, ϕ_a_lb = a, ϕ_a
        end
    end
    return ϕ_a, a, evals, :zoom_max_iters_reached                             # :208
end

# ---------------- findfeasiblestepsize! (wolfe.jl:171-207)
function findfeasible!(info::DeviceWorkspace, fdf!, evals::Int, a::Float64, reduction::Float64, lb::Float64, max_iters::Int)
    xp, df_xp, x, u = _vecs(info)
    lb > a && return 0.0, 0.0, a, evals, :lb_larger                           # :186-188
    ϕ_a, dϕ_a = evalϕdϕ!(xp, df_xp, fdf!, a, x, u)                            # :191
    evals += 1
    iter = 1
    while a > lb && iter < max_iters                                          # :195
        (isfinite(ϕ_a) && isfinite(dϕ_a)) && return ϕ_a, dϕ_a, a, evals, :success
        a = a * reduction                                                     # :200
        ϕ_a, dϕ_a = evalϕdϕ!(xp, df_xp, fdf!, a, x, u)
        evals += 1
        iter += 1
    end
    return ϕ_a, dϕ_a, a, evals, :infeasible                                   # :206
end

# evalwolfeconditions with u on the device (wolfe.jl:219-251, :264-294)
function CGO.evalwolfeconditions(c::YuanWeiLuWolfe{Float64}, ϕ_a, dϕ_a, a, u::DeviceVector, ϕ_0, dϕ_0)
    @assert 0.0 < c.δ1 < c.c1 < c.c2 < 1.0                                    # :233
    norm_u_sq = devdot(u, u)                                                  # :240
    chk1 = ϕ_a <= ϕ_0 + c.c1 * a * dϕ_0 + a * min(-c.δ1 * dϕ_0, c.c1 * a * norm_u_sq / 2)   # :243-244
    chk2 = dϕ_a >= c.c2 * dϕ_0 + min(-c.δ1 * dϕ_0, c.c1 * a * norm_u_sq)                    # :247-248
    return chk1, chk2
end
function CGO.evalwolfeconditions(c::Wolfe{Float64}, ϕ_a, dϕ_a, a, u::DeviceVector, ϕ_0, dϕ_0)
    @assert 0.0 < c.c1 < c.c2 < 1.0                                           # :278
    return ϕ_a <= ϕ_0 + c.c1 * a * dϕ_0, dϕ_a >= c.c2 * dϕ_0                  # :285-290
end

# ---------------- WolfeBisection (wolfe.jl:13-165)
function CGO.linesearch!(info::DeviceWorkspace, config::WolfeBisection{Float64,CT}, fdf!::DeviceObjective,
                         f_x::Float64, df_x::DeviceVector, a_initial::Float64) where CT
    reduction, growth = 0.5, 2.0                                              # :23-24
    xp, df_xp, x, u = _vecs(info)
    max_step = config.max_step_size
    (max_step > a_initial > 0.0) || (a_initial = min(1.0, max_step / 2))      # :30-32
    ϕ_0 = f_x
    isfinite(ϕ_0) || return ϕ_0, 0.0, 0, :accepted_non_finite_iterate         # :36-38
    info.hint = a_initial
    dϕ_0 = devdot(df_x, u)                                                    # :40
    dϕ_0 > 0.0 && return ϕ_0, 0.0, 0, :non_descent_search_direction           # :41-43
    a, evals = a_initial, 0
    lb, ub = 0.0, Inf
    ϕ_a, dϕ_a, a, evals, st = findfeasible!(info, fdf!, evals, a, reduction, 0.0, config.feasibility_max_iters)
    st == :success || return ϕ_0, 0.0, 0, :cannot_find_initial_feasible_step  # :64
    for _ = 1:config.max_iters                                                # :67
        valid_large, valid_small = evalwolfeconditions(config.condition, ϕ_a, dϕ_a, a, u, ϕ_0, dϕ_0)
        if !valid_large || !valid_small
            if !valid_large
                ub = a                                                        # :86
                a = (lb + ub) / 2                                             # :95
            else
                lb = a                                                        # :98
                if !isfinite(ub)
                    a = growth * a                                            # :102
                    a > max_step && return ϕ_0, 0.0, 0, :max_step_length_reached   # :104-112
                else
                    a = (lb + ub) / 2                                         # :114
                end
            end
            if !(lb < a < ub)                                                 # :122
                if !isapprox(norm_u_plus_g(info), 0.0)                        # :123
                    lb, ub = 0.0, Inf
                    a = a_initial
                    resetdirection!(info)                                     # :129  u[:] = −df_x (dϕ_0 NOT recomputed)
                end                                                           # :131 builds a tuple but does not return it
            end
            ϕ_a, dϕ_a, a, evals, st = findfeasible!(info, fdf!, evals, a, reduction, lb, config.feasibility_max_iters)
            st == :success || return ϕ_0, 0.0, 0, :cannot_find_feasible_step  # :153-158
        else
            return ϕ_a, a, evals, :success                                    # :160
        end
    end
    return ϕ_a, a, evals, :linesearch_max_iters_reached                       # :164
end

# ---------------- Backtracking (geometric.jl:22-152)
function CGO.linesearch!(info::DeviceWorkspace, config::Backtracking{Float64,CT}, fdf!::DeviceObjective,
                         f_x::Float64, df_x::DeviceVector, a_initial::Float64) where CT
    xp, df_xp, x, u = _vecs(info)
    ϕ_0 = f_x
    isfinite(ϕ_0) || return ϕ_0, 0.0, 0, :accepted_non_finite_iterate         # :39-41
    isfinite(a_initial) && (info.hint = a_initial)
    dϕ_0 = devdot(df_x, u)                                                    # :43
    dϕ_0 > 0.0 && return ϕ_0, 0.0, 0, :non_descent_search_direction           # :44-46
    evals = 0
    a = a_initial
    isfinite(a) || (a = abs(ϕ_0) / devdot(u, u))                              # :50-53
    isfinite(a) || (a = 1.0)                                                  # :54-57
    ϕ_a, dϕ_a, a, evals, st = findfeasible!(info, fdf!, evals, a, 0.5, 0.0, config.feasibility_max_iters)
    st == :success || return ϕ_0, 0.0, 0, :cannot_find_initial_feasible_step  # :74
    ϕ_a, dϕ_a = evalϕdϕ!(xp, df_xp, fdf!, a, x, u)                            # :78 (redundant second eval, kept)
    evals += 1
    valid = evalbacktrackcondition(config.condition, ϕ_a, a, ϕ_0, dϕ_0)       # :81
    # geometricsearch! (:102-152): grow (a/ρ) while valid, else shrink (a·ρ)
    ρ = config.discount_factor
    a_prev, ϕ_prev = a, ϕ_a
    for _ = 1:config.max_iters
        a = valid ? a / ρ : a * ρ                                             # :127
        isfinite(a) || return ϕ_prev, a_prev, evals, :non_finite_step_proposed          # :128-130
        a == a_prev && return ϕ_prev, a_prev, evals, :proposed_step_same_as_current_step # :132-134
        ϕ_a, _ = evalϕdϕ!(xp, df_xp, fdf!, a, x, u)                           # :137
        evals += 1
        if !evalbacktrackcondition(config.condition, ϕ_a, a, ϕ_0, dϕ_0)       # :140-144
            return ϕ_prev, a_prev, evals, :success   # previous (ϕ, a); xp/df_xp hold the rejected trial (quirk kept)
        end
        a_prev, ϕ_prev = a, ϕ_a
    end
    return ϕ_a, a, evals, :linesearch_max_iters_reached                       # :151
end
