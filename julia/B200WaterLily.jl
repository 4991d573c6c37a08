# B200WaterLily.jl — Julia binding of libwl_b200.so (include/wl_b200.h) for WaterLily.jl's plug-in surface.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image.  The file is a thin `ccall` layer on
# purpose; everything it calls is exercised through the identical C ABI by the Python mirror (waterlily.jl_b200/) and tests/.
#
#   using WaterLily, B200WaterLily
#   sim = Simulation((512,512,512), (0,0,0), 512; ν=…, perdir=(1,2,3), u0=…, T=Float32,
#                    flow_ctor=B200WaterLily.flow_ctor, pois_ctor=B200WaterLily.pois_ctor)
#   sim_step!(sim, 10; remeasure=false)        # dispatches to mom_step!(::B200Flow, ::B200Poisson) below
module B200WaterLily

using WaterLily
import WaterLily: AbstractFlow, AbstractPoisson, mom_step!, measure!, update!, time, CFL

const lib = get(ENV, "WL_B200_LIB", joinpath(@__DIR__, "..", "waterlily.jl_b200", "csrc", "libwl_b200.so"))

struct WLConfig            # struct wl_config
    D::Int32; n::NTuple{3,Int32}; uBC::NTuple{3,Float32}; perdir::NTuple{3,Int32}
    exitBC::Int32; lambda::Int32; nu::Float32; dt0::Float32
    pois_kind::Int32; smoother::Int32; tol::Float32; itmx::Int32; device::Int32; flags::Int32
end

check(rc) = rc == 0 || error(unsafe_string(ccall((:wl_last_error, lib), Cstring, ())))

const FIELD = (u=0, u⁰=1, f=2, p=3, σ=4, V=5, μ₀=6, μ₁=7)
lam_id(λ) = λ === WaterLily.quick ? 0 : λ === WaterLily.cds ? 1 : λ === WaterLily.vanLeer ? 2 :
            error("B200WaterLily: only quick, cds and vanLeer are compiled into the flux kernel")

"Flow replacement (src/Flow.jl:114-148): all fields live on the B200 inside the handle."
mutable struct B200Flow{D} <: AbstractFlow{D,Float32}
    h::Ptr{Cvoid}
    N::NTuple{D,Int}
    uBC::NTuple{D,Float32}; ν::Float32; exitBC::Bool; perdir::NTuple; λ
    Δt::Vector{Float32}                      # mirrored after every step (flow.Δt is read by sim_time, JLD2/VTK exts)
end

function flow_ctor(dims::NTuple{D}, uBC; u0=nothing, Δt=0.25, ν=0., g=nothing, λ=WaterLily.quick, T=Float32, mem=Array,
                   perdir=(), exitBC=false, kw...) where D
    uBC isa Function && error("B200WaterLily: function-valued uBC is not supported (host closure)")
    isnothing(g) || error("B200WaterLily: g(i,x,t) is not supported (host closure)")
    T === Float32 || error("B200WaterLily: only T=Float32")
    pad3(t, z) = ntuple(i -> i <= D ? t[i] : z, 3)
    cfg = WLConfig(D, Int32.(pad3(dims, 1)), Float32.(pad3(uBC, 0)), Int32.(ntuple(i -> i in perdir, 3)), exitBC, lam_id(λ),
                   ν, Δt, 0, 0, 1f-4, 0, 0, 0)
    h = Ref{Ptr{Cvoid}}()
    check(ccall((:wl_create, lib), Cint, (Ref{WLConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    a = B200Flow{D}(h[], dims .+ 2, Float32.(uBC), ν, exitBC, perdir, λ, Float32[Δt])
    if !isnothing(u0)                        # apply!(u0,u) on the host, upload, BC!/exitBC!/u⁰ on the device (src/Flow.jl:140-142)
        u = Array{Float32}(undef, a.N..., D)
        u0 isa Function ? WaterLily.apply!(u0, u) : WaterLily.apply!((i, x) -> u0[i], u)
        upload!(a, :u, u); check(ccall((:wl_apply_bc, lib), Cint, (Ptr{Cvoid},), a.h))
    end
    finalizer(x -> ccall((:wl_destroy, lib), Cint, (Ptr{Cvoid},), x.h), a)
end

upload!(a::B200Flow, f::Symbol, A::Array{Float32}) =
    check(ccall((:wl_upload, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}, Cint), a.h, FIELD[f], A, 0))
# one component (0-based `i`) of a vector field held separately on the host: no assembled copy
upload_component!(a::B200Flow, f::Symbol, i::Integer, A::Array{Float32}) =
    check(ccall((:wl_upload_component, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float32}, Cint), a.h, FIELD[f], i, A, 0))
function download(a::B200Flow{D}, f::Symbol) where D
    nc = f in (:p, :σ) ? () : f === :μ₁ ? (D, D) : (D,)
    A = Array{Float32}(undef, a.N..., nc...)
    check(ccall((:wl_download, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}, Cint), a.h, FIELD[f], A, 0)); A
end
# flow.u, flow.p, … materialise lazily from the device (Metrics, MeanFlow, JLD2/VTK, Makie read them)
Base.getproperty(a::B200Flow, s::Symbol) = haskey(FIELD, s) ? download(a, s) : getfield(a, s)

"MultiLevelPoisson replacement: a view on the hierarchy owned by the flow's handle (x≡p, L≡μ₀, z≡σ alias as in the reference)."
struct B200Poisson <: AbstractPoisson{Float32,Array{Float32},Array{Float32}}
    flow::B200Flow
    n::Vector{Int16}
end
# Simulation's constructor calls measure!(flow,body) and only then pois_ctor(flow) (src/WaterLily.jl:104-105): the reference
# builds D, iD and the coarse levels from the measured μ₀, so this constructor rebuilds the handle's hierarchy (the one wl_create
# made belongs to the body-free μ₀≡1).  The library would also do it by itself before the next step (any upload of μ₀/μ₁/V marks
# the hierarchy stale), but the reference order is kept explicit here.
function pois_ctor(flow::B200Flow)
    check(ccall((:wl_update, lib), Cint, (Ptr{Cvoid},), flow.h))
    B200Poisson(flow, Int16[])
end
update!(b::B200Poisson) = check(ccall((:wl_update, lib), Cint, (Ptr{Cvoid},), b.flow.h))

function sync_histories!(a::B200Flow, b::B200Poisson)
    len = Ref{Cint}(0)
    check(ccall((:wl_get_dt, lib), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ref{Cint}), a.h, C_NULL, len))
    resize!(a.Δt, len[]); check(ccall((:wl_get_dt, lib), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ref{Cint}), a.h, a.Δt, len))
    check(ccall((:wl_get_iters, lib), Cint, (Ptr{Cvoid}, Ptr{Int16}, Ref{Cint}), a.h, C_NULL, len))
    resize!(b.n, len[]); check(ccall((:wl_get_iters, lib), Cint, (Ptr{Cvoid}, Ptr{Int16}, Ref{Cint}), a.h, b.n, len))
end

"""mom_step!(a,b;udf,kwargs...) (src/Flow.jl:156-167).  The one udf the library has built in is WaterLily's own LES model,
`sim_step!(sim; udf=sgs!, νₜ=smagorinsky, S, Cs, Δ)` (src/util.jl:46-76): `wl_set_sgs` switches the device implementation on for the
step (νₜ must be the docstring's Smagorinsky–Lilly function — the library evaluates exactly that; `S` is scratch the library owns).
Any other udf is a host closure and is rejected."""
function mom_step!(a::B200Flow, b::B200Poisson; udf=nothing, νₜ=nothing, S=nothing, Cs=0, Δ=0, kwargs...)
    if udf === WaterLily.sgs!
        check(ccall((:wl_set_sgs, lib), Cint, (Ptr{Cvoid}, Cfloat, Cfloat), a.h, Cs, Δ))
    elseif isnothing(udf)
        check(ccall((:wl_set_sgs, lib), Cint, (Ptr{Cvoid}, Cfloat, Cfloat), a.h, 0, 0))
    else
        error("B200WaterLily: udf is a host closure (built in: udf=sgs! with νₜ=smagorinsky)")
    end
    check(ccall((:wl_mom_step, lib), Cint, (Ptr{Cvoid},), a.h)); sync_histories!(a, b)
end

"measure!(flow,body) for a static body: measured once on the host by WaterLily itself, then uploaded (src/Body.jl:28-51)."
function measure!(a::B200Flow{D}, body::WaterLily.AbstractBody; t=0f0, ϵ=1) where D
    body isa WaterLily.NoBody && return
    host = WaterLily.Flow(a.N .- 2, a.uBC; perdir=a.perdir, exitBC=a.exitBC, T=Float32)   # scratch CPU flow
    WaterLily.measure!(host, body; t, ϵ)          # includes the trailing BC!(μ₀,0), BC!(V,0,exitBC) (src/Body.jl:49-50)
    upload!(a, :μ₀, host.μ₀); upload!(a, :μ₁, host.μ₁); upload!(a, :V, host.V); upload!(a, :σ, host.σ)
    # every one of these uploads marks the Poisson hierarchy stale inside the library: it is rebuilt by pois_ctor / update!,
    # or at the latest by the next wl_mom_step — sim_step!(sim; remeasure=true) therefore works with host-measured bodies
end
time(a::B200Flow) = sum(@view(a.Δt[1:end-1]))
CFL(a::B200Flow) = (v = Ref{Float32}(); check(ccall((:wl_cfl, lib), Cint, (Ptr{Cvoid}, Ref{Float32}), a.h, v)); v[])

# ---- device-side versions of the user-facing helpers around the step (SURVEY.md §8f); each is one ccall ------------------------
struct WLBodyPrim          # struct wl_body_prim: AutoBody(sphere | x-axis torus, (x,t) -> x .- vel.*t), combined by ∪ ∩ −
    kind::Int32; op::Int32; center::NTuple{3,Float32}; R::Float32; r::Float32; vel::NTuple{3,Float32}
end
"Registers a parametrised body; measure!(sim) and sim_step!(sim; remeasure=true) then run on the device (src/Body.jl:28-51)."
set_body!(a::B200Flow, prims::Vector{WLBodyPrim}; ϵ=1) =
    check(ccall((:wl_set_body, lib), Cint, (Ptr{Cvoid}, Ptr{WLBodyPrim}, Cint, Cfloat), a.h, prims, length(prims), ϵ))
measure_device!(a::B200Flow; t) = check(ccall((:wl_measure, lib), Cint, (Ptr{Cvoid}, Cfloat), a.h, t))
remeasure!(a::B200Flow, on::Bool) = check(ccall((:wl_set_remeasure, lib), Cint, (Ptr{Cvoid}, Cint), a.h, on))
"pressure_force, viscous_force, pressure_moment(x₀), viscous_moment(x₀) of the registered body (src/Metrics.jl:111-190): 4×3 Float64."
function body_forces(a::B200Flow; x₀=(0f0, 0f0, 0f0))
    out = zeros(Float64, 12); x = Float32[x₀...]
    check(ccall((:wl_body_forces, lib), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float64}), a.h, x, out)); reshape(out, 3, 4)
end
"MeanFlow on the device (src/Metrics.jl:205-257): init / update! / reset! / copy!(flow, meanflow)."
meanflow_init!(a::B200Flow; uu_stats=false) = check(ccall((:wl_meanflow_init, lib), Cint, (Ptr{Cvoid}, Cint), a.h, uu_stats))
meanflow_update!(a::B200Flow) = check(ccall((:wl_meanflow_update, lib), Cint, (Ptr{Cvoid},), a.h))
meanflow_reset!(a::B200Flow; t_init=0f0) = check(ccall((:wl_meanflow_reset, lib), Cint, (Ptr{Cvoid}, Cfloat), a.h, t_init))
meanflow_copy!(a::B200Flow) = check(ccall((:wl_meanflow_copy_to_flow, lib), Cint, (Ptr{Cvoid},), a.h))
"Enumerated forcings in place of the closures g(i,x,t), uBC(i,x,t) (src/Flow.jl:64-73): g = g0 + g1·t, U = uBC + U1·t + ½U2·t²."
set_forcing!(a::B200Flow; g0=zeros(Float32, 3), g1=zeros(Float32, 3), U1=zeros(Float32, 3), U2=zeros(Float32, 3)) =
    check(ccall((:wl_set_forcing, lib), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}), a.h,
                Float32.(g0), Float32.(g1), Float32.(U1), Float32.(U2)))

end # module
