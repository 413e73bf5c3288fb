# β plugins on the device workspace — methods added to the reference's plugin functions
# (src/cg_flavours.jl:2-35: updatedir!, initializeβ, initializeLineSearchContainer!, getβ).
# Every vector pass of the reference's getβ (3 temporaries, 3–6 dots) was already done by the
# trial kernel; getβ combines the scalar pack here so that Julia's NaN / `max` semantics apply.

function CGO.updatedir!(u::DeviceVector, df_x::DeviceVector, β::Float64)     # cg_flavours.jl:2-15
    @assert u.ws === df_x.ws
    u.ws.cached = nothing
    u.ws.pending_β = β                    # deferred: fused into the first trial of the next line search
    return nothing
end

function CGO.initializeLineSearchContainer!(info::DeviceWorkspace, ::CGβConfig, df_x::DeviceVector, x::DeviceVector)
    resetdirection!(info)                 # cg_flavours.jl:22-35: u = −df_x (x, df_x already on the device)
    return nothing
end

function hzfamily(ws::DeviceWorkspace, R::Float64)
    P = ws.pack
    m = 2 * P[P_YY] / R                                              # cg_flavours.jl:73 / :102
    if ws.β_literal                       # tmp2 = g_next ./ R; tmp1 = y − m .* u; dot(tmp1, tmp2)   :71-76
        v = Ref{Float64}(0.0)
        check(ccall((:cgo_beta_literal, LIBCGOPTIM[]), Cint, (Ptr{Cvoid}, Float64, Float64, Ref{Float64}), ws.h, R, m, v))
        return v[]
    end
    return (P[P_YGP] - m * P[P_DPHI]) / R # Σ (y_i − m u_i)(g⁺_i / R) = (y·g⁺ − m u·g⁺)/R
end

function CGO.getβ(::HagerZhang, g_next::DeviceVector, g::DeviceVector, u::DeviceVector)   # :87-108
    ws = g_next.ws
    return hzfamily(ws, ws.pack[P_UY])                               # R = dot(u, y)   :98
end
function CGO.getβ(c::YuanWangSheng{Float64}, g_next::DeviceVector, g::DeviceVector, u::DeviceVector)  # :51-79
    P = g_next.ws.pack
    R1 = c.μ * sqrt(P[P_UU]) * sqrt(P[P_YY])                         # :65
    R2 = P[P_UY]                                                     # :66
    R3 = 2 * P[P_YY] * P[P_DPHI] / P[P_YGP]                          # :67
    return hzfamily(g_next.ws, max(R1, R2, R3))                      # :68 (NaN-propagating max)
end
function CGO.getβ(::SallehAlhawarat, g_next::DeviceVector, g::DeviceVector, u::DeviceVector)   # :133-151
    P = g_next.ws.pack
    norm_sq = sqrt(P[P_GPGP])^2                                      # :140 norm(g_next)^2
    tmp = P[P_GPG]                                                   # :141
    norm_sq > tmp || return 0.0
    return (norm_sq - tmp) / (P[P_DPHI] - P[P_UG])                   # :145
end
function CGO.getβ(::LiuStorrey, g_next::DeviceVector, g::DeviceVector, u::DeviceVector)        # :157-170
    P = g_next.ws.pack
    return P[P_YGP] / (-P[P_UY])                                     # :166-167
end

# ---- LBFGS(m): the quasi-Newton slot (src/qn_flavours.jl is a dense no-op update, SURVEY.md §0)
struct LBFGS <: QNβConfig
    m::Int
    LBFGS(m::Integer = 10) = (@assert 1 <= m <= 64; new(m))
end
struct LBFGSHistoryToken end
CGO.initializeβ(::Type{T}, ::LBFGS) where T = LBFGSHistoryToken()
CGO.initializeLineSearchContainer!(info::DeviceWorkspace, ::LBFGS, df_x::DeviceVector, x::DeviceVector) = resetdirection!(info)
function CGO.getβ(::LBFGS, g_next::DeviceVector, g::DeviceVector, u::DeviceVector)
    ws = g_next.ws                        # called at optim.jl:130, before x ← xp: s = xp − x, y = g⁺ − g
    check(ccall((:cgo_lbfgs_stage_pair, LIBCGOPTIM[]), Cint, (Ptr{Cvoid}, Ptr{Float64}), ws.h, ws.buf))
    sy, yy = ws.buf[1], ws.buf[2]
    if sy > 0.0
        check(ccall((:cgo_lbfgs_commit_pair, LIBCGOPTIM[]), Cint, (Ptr{Cvoid}, Int32, Float64, Float64), ws.h, 1, 1.0 / sy, sy / yy))
    else
        check(ccall((:cgo_lbfgs_commit_pair, LIBCGOPTIM[]), Cint, (Ptr{Cvoid}, Int32, Float64, Float64), ws.h, 0, 0.0, 1.0))
    end
    return LBFGSHistoryToken()
end
function CGO.updatedir!(u::DeviceVector, df_x::DeviceVector, ::LBFGSHistoryToken)
    ws = u.ws
    ws.cached = nothing; ws.pending_β = nothing
    check(ccall((:cgo_lbfgs_update_dir, LIBCGOPTIM[]), Cint, (Ptr{Cvoid}, Ptr{Float64}), ws.h, ws.buf))
    ws.dpack .= ws.buf[1:2]
    return nothing
end
