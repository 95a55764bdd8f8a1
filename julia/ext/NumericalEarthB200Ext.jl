# NumericalEarthB200Ext — package extension that re-hosts the atmosphere–surface interface
# computation of NumericalEarth.jl on libne_b200.so (hand-written sm_100a CUDA kernels).
#
# STATUS: written against include/ne_b200.h; Julia is not installed in the build container, so this
# file is NOT executed by the test-suite.  It is kept thin: every `ccall` below has a 1:1 Python
# ctypes twin in numericalearth.jl_b200/{abi,interface,formulations}.py which IS tested against the
# oracle on a B200.  See INTEGRATION.md for the binding rules.
#
# Activation (mirrors ext/NumericalEarthReactantExt.jl:18-31, which overrides behaviour by dispatch on
# the architecture type parameter): the user builds the exchange grid on `B200(GPU())`; the methods
# below then win dispatch for `EarthSystemModel`s whose exchange grid lives on that architecture.
#
#   Project.toml additions:
#     [weakdeps]   CUDA = "052768ef-5323-5732-b1bb-66c8b64840ba"
#     [extensions] NumericalEarthB200Ext = "CUDA"
#   ENV["NE_B200_LIB"] = "/path/to/libne_b200.so"
module NumericalEarthB200Ext

using NumericalEarth
using NumericalEarth.EarthSystemModels: EarthSystemModel
using NumericalEarth.EarthSystemModels.InterfaceComputations:
    SimilarityTheoryFluxes, CoefficientBasedFluxes, ConvergenceStopCriteria, FixedIterations,
    MomentumRoughnessLength, ScalarRoughnessLength, ReynoldsScalingFunction, WindDependentWaveFormulation,
    TemperatureDependentAirViscosity, ConvectiveGustiness, SubgridVelocityCorrection, SimilarityScales,
    EdsonMomentumStabilityFunction, EdsonScalarStabilityFunction, ShebaMomentumStabilityFunction,
    ShebaScalarStabilityFunction, PaulsonMomentumStabilityFunction, PaulsonScalarStabilityFunction,
    LinearStableStabilityFunction, SplitStabilityFunction, LogarithmicSimilarityProfile,
    COARELogarithmicSimilarityProfile, PolynomialNeutralDragCoefficient, LargeYeagerTransferCoefficients,
    ImpureSaturationSpecificHumidity, WaterMoleFraction, BulkTemperature, SkinTemperature, DiffusiveFlux,
    InteriorDiffusivity, RelativeVelocity, WindVelocity, interface_kernel_parameters
using Oceananigans
using Oceananigans.Architectures: AbstractArchitecture, GPU, architecture
using CUDA

const libne = get(ENV, "NE_B200_LIB", "libne_b200.so")

"`B200(GPU())`: architecture wrapper that selects the libne_b200 kernels (cf. ReactantState)."
struct B200{A} <: AbstractArchitecture
    child :: A
end
Oceananigans.Architectures.child_architecture(a::B200) = a.child
Oceananigans.Architectures.device(a::B200) = Oceananigans.Architectures.device(a.child)
Oceananigans.Architectures.array_type(a::B200) = Oceananigans.Architectures.array_type(a.child)

const B200Model = EarthSystemModel{<:Any, <:Any, <:Any, <:Any, <:Any, <:Any, <:B200}

#####
##### POD mirrors of include/ne_b200.h (isbits, C layout).  Field order == header order.
#####

struct NeSlot;           ptr::Ptr{Cvoid}; value::Cdouble; end
struct NeExchangeGrid;   nx::Int64; ny::Int64; hx::Int64; hy::Int64; i_lo::Int64; i_hi::Int64; j_lo::Int64; j_hi::Int64; end
struct NeStabilityFn;    kind::Int32; pad::Int32; p::NTuple{12, Cdouble}; end
struct NeStabilityProfile; split::Int32; pad::Int32; a::NeStabilityFn; b::NeStabilityFn; end
struct NeStopCriteria;   kind::Int32; maxiter::Int32; tolerance::Cdouble; end
# … NeRoughnessLength, NeSubgridVelocity, NeFluxFormulation, NeInterfaceProperties, NeMediumProperties,
#   NeSurfaceRadiation, NeThermoParams, NeInterpDesc, NeAtmosOceanDesc, NeAtmosSeaIceDesc, NeSeaIceOceanDesc,
#   NeAssembleOceanDesc, NeAssembleSeaIceDesc, NeApplyRadiationDesc follow the header field by field
#   (see numericalearth.jl_b200/abi.py for the complete, tested list).  `ne_struct_size(name)` is checked
#   against `sizeof(T)` for every mirror in `__init__`.

function __init__()
    for (name, T) in (("NeSlot", NeSlot), ("NeExchangeGrid", NeExchangeGrid), ("NeStabilityFn", NeStabilityFn),
                      ("NeStabilityProfile", NeStabilityProfile), ("NeStopCriteria", NeStopCriteria))
        n = ccall((:ne_struct_size, libne), Int64, (Cstring,), name)
        n == sizeof(T) || error("libne_b200 ABI mismatch for $name: C $n vs Julia $(sizeof(T))")
    end
end

check(rc) = rc == 0 || begin
    msg = unsafe_string(ccall((:ne_last_error, libne), Cstring, ()))
    rc == -2 ? throw(ArgumentError("no sm_100a kernel variant: $msg (no CPU fallback)")) : error("libne_b200: $msg")
end

#####
##### plugin types -> POD variants.  Anything else is an ArgumentError (no KernelAbstractions fallback).
#####

pad12(v...) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 12)

stability_fn(ψ::EdsonMomentumStabilityFunction) = NeStabilityFn(1, 0, pad12(ψ.ζmax, ψ.A⁺, ψ.B⁺, ψ.C⁺, ψ.D⁺, ψ.A⁻, ψ.B⁻, ψ.C⁻, ψ.D⁻, ψ.E⁻, ψ.F⁻))
stability_fn(ψ::EdsonScalarStabilityFunction)   = NeStabilityFn(2, 0, pad12(ψ.ζmax, ψ.A⁺, ψ.B⁺, ψ.C⁺, ψ.D⁺, ψ.E⁺, ψ.A⁻, ψ.B⁻, ψ.C⁻, ψ.D⁻, ψ.E⁻, ψ.F⁻))
stability_fn(ψ::ShebaMomentumStabilityFunction) = NeStabilityFn(3, 0, pad12(ψ.a, ψ.b))
stability_fn(ψ::ShebaScalarStabilityFunction)   = NeStabilityFn(4, 0, pad12(ψ.a, ψ.b, ψ.c))
stability_fn(ψ::PaulsonMomentumStabilityFunction) = NeStabilityFn(5, 0, pad12(ψ.a, ψ.b))
stability_fn(ψ::PaulsonScalarStabilityFunction)   = NeStabilityFn(6, 0, pad12(ψ.a))
stability_fn(ψ::LinearStableStabilityFunction)    = NeStabilityFn(7, 0, pad12(ψ.coefficient, ψ.maximum_stability_parameter))
stability_fn(::Returns)                           = NeStabilityFn(0, 0, pad12())   # Returns(zero(FT))
stability_fn(ψ) = throw(ArgumentError("stability function $(typeof(ψ)) has no sm_100a kernel variant"))

stability_profile(ψ::SplitStabilityFunction) = NeStabilityProfile(1, 0, stability_fn(ψ.stable), stability_fn(ψ.unstable))
stability_profile(ψ) = NeStabilityProfile(0, 0, stability_fn(ψ), NeStabilityFn(0, 0, pad12()))

stop_criteria(s::ConvergenceStopCriteria) = NeStopCriteria(0, s.maxiter, s.tolerance)
stop_criteria(s::FixedIterations)         = NeStopCriteria(1, s.iterations, 0.0)
stop_criteria(s) = throw(ArgumentError("solver_stop_criteria $(typeof(s)) has no kernel variant"))

#####
##### array unwrapping: parent(field.data) device pointers + (size, halo)
#####

devptr(f::Oceananigans.Fields.Field) = Ptr{Cvoid}(UInt(pointer(parent(f))))
devptr(a::CuArray)                   = Ptr{Cvoid}(UInt(pointer(a)))
slot(f::Oceananigans.Fields.Field)     = NeSlot(devptr(f), 0.0)
slot(::Oceananigans.Fields.ZeroField)  = NeSlot(C_NULL, 0.0)
slot(c::Oceananigans.Fields.ConstantField) = NeSlot(C_NULL, Float64(c.constant))
slot(x::Number)                        = NeSlot(C_NULL, Float64(x))
"pointer to the k = Nz plane of a 3-D field's parent (ocean surface T, S, u, v)"
function surface_plane(f)
    p = parent(f); Nx, Ny, Nz = size(f.grid); Hz = f.grid.Hz
    return NeSlot(Ptr{Cvoid}(UInt(pointer(p, 1 + (Nz + Hz - 1) * size(p, 1) * size(p, 2)))), 0.0)
end

function exchange_grid(grid; halo_ring = true)
    Nx, Ny, _ = size(grid); Hx, Hy, _ = Oceananigans.Grids.halo_size(grid)
    halo_ring ? NeExchangeGrid(Nx, Ny, Hx, Hy, 0, Nx + 1, 0, Ny + 1) : NeExchangeGrid(Nx, Ny, Hx, Hy, 1, Nx, 1, Ny)
end

stream() = Ptr{Cvoid}(UInt(CUDA.stream().handle))
suffix(grid) = eltype(grid) === Float64 ? "f64" : "f32"

#####
##### the overridden entry points (same generic functions update_state! calls,
##### src/EarthSystemModels/time_step_earth_system_model.jl:38-83)
#####

function NumericalEarth.EarthSystemModels.InterfaceComputations.compute_atmosphere_ocean_fluxes!(model::B200Model)
    desc = atmosphere_ocean_descriptor(model)   # fills NeAtmosOceanDesc as numericalearth.jl_b200/interface.py:atmosphere_ocean_desc
    f = eltype(model.interfaces.exchanger.grid) === Float64 ? :ne_atmosphere_ocean_fluxes_f64 : :ne_atmosphere_ocean_fluxes_f32
    GC.@preserve model check(ccall((f, libne), Cint, (Ref{NeAtmosOceanDesc}, Ptr{Cvoid}), desc, stream()))
    return nothing
end

# compute_atmosphere_sea_ice_fluxes!, compute_sea_ice_ocean_fluxes!, interpolate_state!(exchanger, grid,
# ::PrescribedAtmosphere / ::PrescribedRadiation, model), update_net_fluxes!(model, ocean / sea_ice),
# apply_air_sea_radiative_fluxes!, apply_air_sea_ice_radiative_fluxes! and
# initialize!(exchanger::ComponentExchanger, grid, ::PrescribedAtmosphere) follow the same three lines:
# build the descriptor (INTEGRATION.md §3 lists the field-by-field mapping), pick the _f64/_f32 symbol,
# ccall on CUDA.stream().  The host-side TimeInterpolator is computed exactly as the reference does
# (cpu_interpolating_time_indices, src/Atmospheres/interpolate_atmospheric_state.jl:57-60) and passed in
# NeTimeInterp, and fractional indices may be left to the reference's own initialize! so that they are
# bit-exact by construction.
#
# Also behind ne_interp_state_*: interpolate_state!(exchanger, grid, ::PrescribedLand, model)
# (src/Lands/interpolate_land_state.jl:6-61; one output field, n_summands = number of runoff series, fractional indices
# of the land grid from ne_frac_indices_* once), and — on TripolarGrid / rotated exchange grids — the intrinsic_vector
# rotation (interpolate_atmospheric_state.jl:123-126): `NeInterpDesc.rotation_cos/sin` are filled once in initialize!
# from Oceananigans' rotation metrics of the exchange grid, `rotate_u/v = 0/1`.
# Host-resident ocean state: ne_host_pipeline_create once, ne_host_pipelined_step_* per coupled step (INTEGRATION.md §5).
#
# Partly-in-memory series: Oceananigans.TimeSteppers.update_state!(atmos::PrescribedAtmosphere) for a B200 model does not
# call update_field_time_series! (whole-window set!(fts), src/Atmospheres/prescribed_atmosphere.jl:154-162); the parent of
# every series is a ring of Nt_mem slices owned by ne_series_ring_create, and the interpolate_state! override does
# load (first step / clock jump) -> acquire -> launch -> release -> prefetch loads, as INTEGRATION.md §6 spells out;
# the raw slices come from the same NCDatasets reads set!(fts) performs (JRA55_field_time_series.jl:60-76), one time
# index at a time, into pinned host buffers.

end # module
