# ccall bindings of libcgoptim.so — one Julia function per entry point of include/cgoptim.h.
import Libdl

const LIBCGOPTIM = Ref{String}(get(ENV, "LIBCGOPTIM", joinpath(@__DIR__, "..", "..", "..",
    "conjugategradientoptim.jl_b200", "libcgoptim.so")))
const PACK_LEN = 16
# pack indices (+1: Julia arrays are 1-based), include/cgoptim.h
const P_PHI, P_DPHI, P_GPGP, P_YY, P_UY, P_YGP, P_GPG, P_UG, P_UU, P_DIR_GU, P_DIR_UU, P_XPXP = 1:12
const D_GU, D_UU = 1, 2

struct CgoError <: Exception
    code::Cint
    msg::String
end

lasterror() = unsafe_string(ccall((:cgo_last_error, LIBCGOPTIM[]), Cstring, ()))
# CUDA / NCCL / argument errors throw; numerical trouble never does (it is data: statuses)
check(rc::Cint) = rc == 0 ? nothing : throw(CgoError(rc, lasterror()))

macro cgo(name, argtypes, args...)
    quote
        check(ccall(($(QuoteNode(name)), LIBCGOPTIM[]), Cint, $(esc(argtypes)), $(map(esc, args)...)))
    end
end
