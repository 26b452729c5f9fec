# WaterLilyB200.jl — the `ccall` binding a WaterLily.jl maintainer would load to make
# `Simulation(...; mem=B200Array)` run `mom_step!` in libwaterlily_b200.so.
#
# STATUS: written against the reference's API (marinlauber/WaterLily v1.3.0) and the C ABI in
# include/waterlily_b200.h.  There is no Julia in the build image or on the GPU box, so this file
# has NOT been executed; the tested drivers of the same ABI are the ctypes binding in
# waterlily_b200/_cabi.py and tests/.  It is the reference-side stub INTEGRATION.md refers to.
#
# Design: the reference picks its backend by array type (src/WaterLily.jl:66-78, src/Flow.jl:114-118,
# src/util.jl:137).  `B200Array` is an opaque view of one field behind a `wl_sim` handle; methods of
# the reference's functions are overloaded on it.  Construction is two-phase because the reference
# allocates each field separately (`zeros(T,Ng) |> mem`) while the library owns all fields of a
# simulation behind one handle: the first conversion of a Flow's field creates the handle.
module WaterLilyB200

using WaterLily
import WaterLily: mom_step!, BC!, exitBC!, perBC!, CFL, measure!, update!, solver!, Flow, MultiLevelPoisson,
                  AbstractPoisson, BCTuple, time

const lib = get(ENV, "WATERLILY_B200_LIB", "libwaterlily_b200.so")

# enum wl_field
const WL_U, WL_U0, WL_F, WL_P, WL_SIGMA, WL_V, WL_MU0, WL_MU1 = Cint.(0:7)

mutable struct Handle
    ptr::Ptr{Cvoid}
    D::Int
    Ng::NTuple
    function Handle(N::NTuple{D,Int}; ν, Δt, perdir, exitBC, device=0) where D
        out = Ref{Ptr{Cvoid}}(C_NULL)
        per = Cint[(j in perdir) ? 1 : 0 for j in 1:D]
        rc = ccall((:wl_create, lib), Cint,
                   (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}, Cdouble, Cdouble, Ptr{Cint}, Cint, Cint, Cint),
                   out, D, Cint[N...], ν, Δt, per, exitBC, 0, device)
        check(rc)
        h = new(out[], D, N .+ 2)
        finalizer(x -> ccall((:wl_destroy, lib), Cint, (Ptr{Cvoid},), x.ptr), h)
    end
end

lasterror() = unsafe_string(ccall((:wl_last_error, lib), Cstring, ()))
# WL_ERR_ARG reproduces the reference's assertion texts (e.g. "MultiLevelPoisson requires size=a2ⁿ, where n>2")
check(rc) = rc == 0 ? nothing : (rc == -1 ? throw(AssertionError(lasterror())) : error(lasterror()))

"""A field that lives behind a handle.  `Array(a)` downloads, `copyto!(a, host)` uploads."""
struct B200Array{T,N} <: AbstractArray{T,N}
    h::Handle
    field::Cint
    dims::NTuple{N,Int}
end
Base.size(a::B200Array) = a.dims
Base.Array(a::B200Array{T,N}) where {T,N} = (host = Array{Float32,N}(undef, a.dims);
    check(ccall((:wl_download, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Int64), a.h.ptr, a.field, host, length(host))); host)
Base.copyto!(a::B200Array, host::Array{Float32}) =
    (check(ccall((:wl_upload, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Int64), a.h.ptr, a.field, host, length(host))); a)
Base.getindex(a::B200Array, I...) = Array(a)[I...]   # scalar access = download (tests wrap it in @allowscalar)

# --- the hot path -------------------------------------------------------------------------------
# mom_step!(a::Flow,b::AbstractPoisson)  — src/Flow.jl:153-169
function mom_step!(a::Flow{D,T,<:B200Array}, b::AbstractPoisson) where {D,T}
    h = a.u.h
    U = Cdouble[BCTuple(a.U, a.Δt, D)...]                       # :154, evaluated on the host as in the reference
    t0, t1 = sum(@view a.Δt[1:end-1]), sum(a.Δt)
    acc(t) = (a.g === nothing && !(a.U isa Function)) ? C_NULL :
             Cdouble[(a.g === nothing ? 0.0 : a.g(i, t)) +
                     (a.U isa Function ? WaterLily.ForwardDiff.derivative(τ -> a.U(i, τ), t) : 0.0) for i in 1:D]   # :158,165
    dt, n1, n2 = Ref{Cdouble}(0), Ref{Cint}(0), Ref{Cint}(0)
    check(ccall((:wl_mom_step, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ref{Cint}, Ref{Cint}),
                h.ptr, U, acc(t0), acc(t1), dt, n1, n2))
    push!(a.Δt, T(dt[]))                                        # :168
    push!(b.n, Int16(n1[]), Int16(n2[]))                        # MultiLevelPoisson.jl:98, twice per step
    return nothing
end

# BC!(a,A,saveexit,perdir) — src/util.jl:192-210
BC!(a::B200Array, A, saveexit=false, perdir=()) = check(ccall((:wl_bc, lib), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Ptr{Cint}), a.h.ptr, a.field, Cdouble[A...], saveexit,
    Cint[(j in perdir) ? 1 : 0 for j in 1:a.h.D]))
# exitBC!(u,u⁰,U,Δt) — src/util.jl:216-222
exitBC!(u::B200Array, u⁰::B200Array, U, Δt) = check(ccall((:wl_exit_bc, lib), Cint,
    (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble), u.h.ptr, Cdouble[U...], Δt))
# perBC!(a,perdir) — src/util.jl:227-231
perBC!(a::B200Array, perdir) = check(ccall((:wl_per_bc, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}),
    a.h.ptr, a.field, Cint[(j in perdir) ? 1 : 0 for j in 1:a.h.D]))
# CFL(a) — src/Flow.jl:172-175
function CFL(a::Flow{D,T,<:B200Array}; Δt_max=10) where {D,T}
    dt = Ref{Cdouble}(0)
    check(ccall((:wl_cfl, lib), Cint, (Ptr{Cvoid}, Ref{Cdouble}), a.u.h.ptr, dt))
    T(dt[])
end
# update!(ml) — src/MultiLevelPoisson.jl:62-68
update!(ml::MultiLevelPoisson{T,<:B200Array}) where T =
    check(ccall((:wl_pois_update, lib), Cint, (Ptr{Cvoid},), ml.x.h.ptr))
# solver!(ml;tol,itmx) — src/MultiLevelPoisson.jl:87-99
function solver!(ml::MultiLevelPoisson{T,<:B200Array}; tol=1e-4, itmx=32) where T
    n = Ref{Cint}(0)
    check(ccall((:wl_pois_solve, lib), Cint, (Ptr{Cvoid}, Cdouble, Cint, Ref{Cint}), ml.x.h.ptr, tol, itmx, n))
    push!(ml.n, Int16(n[]))
end
# measure!(a::Flow,body;t,ϵ) — src/Body.jl:31-53: the geometry stays in Julia on the host (north star):
# run the reference's own loop on host arrays, then upload μ₀, μ₁, V and apply the two BC! calls on the device.
function measure!(a::Flow{D,T,<:B200Array}, body::WaterLily.AbstractBody; t=zero(T), ϵ=1) where {D,T}
    host = Flow(size(a.p) .- 2, a.U; T, perdir=a.perdir, exitBC=a.exitBC)   # scratch host Flow (mem=Array)
    measure!(host, body; t, ϵ)
    copyto!(a.μ₀, host.μ₀); copyto!(a.μ₁, host.μ₁); copyto!(a.V, host.V); copyto!(a.σ, host.σ)
end
measure!(::Flow{D,T,<:B200Array}, ::WaterLily.NoBody; t=0, ϵ=1) where {D,T} = nothing

# Metrics — src/Metrics.jl:94-141.  nds(body, loc(0,I,Float64), t) is evaluated here, on the host, for every cell
# of the padded grid (the geometry stays in Julia); products and Float64 sums run on the device; df is flow.f.
function nds_field(p::B200Array{T,D}, body, t) where {T,D}
    a = zeros(Float64, size(p)..., D)
    for I in CartesianIndices(size(p)); a[I, :] .= WaterLily.nds(body, loc(0, I, Float64), t); end
    a
end
function pressure_force(p::B200Array{T,D}, df::B200Array, body, t=0, _=Float64) where {T,D}
    nds = nds_field(p, body, t); out = zeros(Cdouble, D)
    check(ccall((:wl_pressure_force, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ptr{Cdouble}), p.h.ptr, nds, length(nds), out)); out
end
function viscous_force(u::B200Array{T}, ν, df::B200Array, body, t=0, _=Float64) where T
    D = ndims(u) - 1; nds = nds_field(df, body, t)[ntuple(_ -> :, D)..., 1, :]; out = zeros(Cdouble, D)   # ν is the handle's own
    check(ccall((:wl_viscous_force, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ptr{Cdouble}), u.h.ptr, nds, length(nds), out)); out
end
function pressure_moment(x₀, p::B200Array{T,D}, df::B200Array, body, t=0, _=Float64) where {T,D}
    nds = nds_field(p, body, t); out = zeros(Cdouble, D)
    check(ccall((:wl_pressure_moment, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}),
                p.h.ptr, collect(Float64, x₀), nds, length(nds), out)); out
end

# derived output fields — `@inside flow.σ[I] = WaterLily.curl(3,I,flow.u)` and friends (src/Metrics.jl:14-77) on the device
const DERIVED = Dict(:ke => 0, :curl => 1, :ω_mag => 2, :λ₂ => 3, :ω_θ => 4)
function derived!(a::Flow{D,T,<:B200Array}, kind::Symbol; i=0, U=nothing, z=nothing, center=nothing) where {D,T}
    dbl(v) = v === nothing ? C_NULL : collect(Float64, v)
    check(ccall((:wl_derived, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                a.σ.h.ptr, DERIVED[kind], i, dbl(U), dbl(z), dbl(center)))
    a.σ
end

end # module
