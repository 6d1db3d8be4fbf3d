# Julia host shim over libbnr.so (C ABI: include/bnr.h).  WRITTEN BUT NOT EXECUTED in the build image: no julia
# binary is installed there.  The identical ABI is exercised from Python/ctypes (bayesiannetworkregression.jl_b200/).
#
# Intended use inside BayesianNetworkRegression.jl:  replace the body of generate_samples! (src/gibbs.jl:897-1020)
# by a call to `fit_b200`, keeping Fit!'s kwargs / parameters.log / Summary untouched.
module BNRB200

using TypedTables

const LIBBNR = get(ENV, "LIBBNR", joinpath(@__DIR__, "..", "bayesiannetworkregression.jl_b200", "libbnr.so"))

# struct bnr_params (include/bnr.h) -- field order and types must match
struct BnrParams
    n::Int32; V::Int32; R::Int32; num_chains::Int32
    chain_offset::Int32; device::Int32; trace_full_chains::Int32; trace_gamma_xi_all::Int32
    trace_rows::Int64; seed::UInt64
    eta::Float64; zeta::Float64; iota::Float64; a_delta::Float64; b_delta::Float64; nu::Float64
    gig_inject_len::Int32; gamma_mode::Int32; chain_groups::Int32; trace_gamma_xi_chains::Int32
end

function check(code::Cint)
    code == 0 && return
    msg = unsafe_string(ccall((:bnr_last_error, LIBBNR), Cstring, ()))
    error("libbnr error $code: $msg")
end

const VARS = (:τ², :u, :ξ, :γ, :S, :θ, :Δ, :M, :μ, :λ, :πᵥ)      # BNR_VAR_* order

"""
    fit_b200(X_new, y, R; η, ζ, ι, aΔ, bΔ, ν, nburn, nsamp, num_chains, seed) -> (state::Table, rhatξ, rhatγ)

Drop-in for the pmap section of generate_samples! (src/gibbs.jl:938-957): `X_new` is the n×q Matrix{Float64}
produced by setup_X!, chains are batched on GPU 0.  The returned Table wraps buffers filled by
bnr_get_trace in the reference's (iteration, d1, d2) layout, so no copy or permutation is needed.
"""
function fit_b200(X_new::Matrix{Float64}, y::Vector{Float64}, R::Integer; η=1.01, ζ=1.0, ι=1.0, aΔ=1.0, bΔ=1.0,
                  ν=10, nburn=30000, nsamp=20000, num_chains=2, seed=1)
    n, q = size(X_new)
    V = Int((-1 + sqrt(1 + 8q)) / 2)
    total = nburn + nsamp
    p = Ref(BnrParams(n, V, R, num_chains, 0, 0, 1, 1, total, UInt64(seed), η, ζ, ι, aΔ, bΔ, Float64(ν), 64, 0, 0, 0))   # gamma_mode = auto, chain_groups = default
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:bnr_create, LIBBNR), Cint, (Ref{BnrParams}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}), p, X_new, y, h))
    try
        check(ccall((:bnr_init_state, LIBBNR), Cint, (Ptr{Cvoid},), h[]))                # row 1 (initialize_variables!)
        check(ccall((:bnr_run, LIBBNR), Cint, (Ptr{Cvoid}, Int64), h[], total - 1))      # rows 2:total (run!)
        check(ccall((:bnr_sync, LIBBNR), Cint, (Ptr{Cvoid},), h[]))
        # return_psrf_VOI: R-hat over rows nburn+1:total of every chain (0-based first row = nburn)
        check(ccall((:bnr_moments_from_trace, LIBBNR), Cint, (Ptr{Cvoid}, Int64, Int64), h[], nburn, nsamp))
        rξ = Vector{Float64}(undef, V); rγ = Vector{Float64}(undef, q)
        check(ccall((:bnr_rhat, LIBBNR), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h[], rξ, rγ))
        dims = Dict(:τ² => (1, 1), :u => (R, V), :ξ => (V, 1), :γ => (q, 1), :S => (q, 1), :θ => (1, 1), :Δ => (1, 1),
                    :M => (R, R), :μ => (1, 1), :λ => (R, 1), :πᵥ => (R, 3))
        cols = map(enumerate(VARS)) do (k, name)
            a = Array{Float64,3}(undef, total, dims[name]...)
            check(ccall((:bnr_get_trace, LIBBNR), Cint, (Ptr{Cvoid}, Int32, Int32, Int64, Int64, Ptr{Float64}),
                        h[], 0, k - 1, 0, total, a))
            name => a
        end
        return Table(; cols...), Table(ξ = rξ), Table(γ = rγ)
    finally
        ccall((:bnr_destroy, LIBBNR), Cint, (Ptr{Cvoid},), h[])
    end
end

end # module
