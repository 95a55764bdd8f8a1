# NumericalEarthB200Ext — package extension that re-hosts the atmosphere–surface interface computation of
# NumericalEarth.jl on libne_b200.so (hand-written sm_100a CUDA kernels behind the C ABI of include/ne_b200.h).
#
# What it does.  `update_state!(model)` (src/EarthSystemModels/time_step_earth_system_model.jl:38-83) calls nine generic
# functions with the coupled model; for an `EarthSystemModel` whose architecture is a CUDA `GPU` this extension adds more
# specific methods of exactly those functions.  Each method unwraps device pointers of the Oceananigans fields, translates
# the plugin TYPE TREE (SimilarityTheoryFluxes{…}, InterfaceProperties{…}, roughness lengths, stability functions, albedos …)
# into the POD variants of the header, and `ccall`s one entry point on the task-local CUDA stream.  Nothing else of the
# model changes: `EarthSystemModel`, `run!(simulation)`, `Callback`s, output writers and the `Checkpointer` see the same
# fields with the same meaning (the interface fields are diagnostic: component_interfaces.jl:527-529).
#
#   reference generic function                          method added here                      entry point
#   initialize!(exchanger, grid, ::PrescribedAtmosphere / ::PrescribedRadiation)                 ne_frac_indices_*      (opt-in)
#   interpolate_state!(exchanger, grid, ::PrescribedAtmosphere, model)                           ne_interp_state_*
#   interpolate_state!(exchanger, grid, ::PrescribedRadiation, model)                            ne_interp_state_*
#   correct_state!(::ElevationCorrection, exchanger, grid)                                       ne_correct_atmosphere_elevation_*
#   compute_atmosphere_ocean_fluxes!(model)                                                      ne_atmosphere_ocean_fluxes_*
#   compute_atmosphere_sea_ice_fluxes!(model)                                                    ne_atmosphere_sea_ice_fluxes_*
#   compute_sea_ice_ocean_fluxes!(model)            (+ the FreezingLimitedOceanTemperature clamp) ne_sea_ice_ocean_stress_*, ne_sea_ice_ocean_fluxes_*
#   update_net_fluxes!(model, ocean), update_net_fluxes!(model, sea_ice)                         ne_assemble_net_ocean_fluxes_*, ne_assemble_net_sea_ice_fluxes_*
#   apply_air_sea_radiative_fluxes!(model), apply_air_sea_ice_radiative_fluxes!(model)           ne_apply_radiative_fluxes_*
#
# A plugin object with no kernel variant (a user closure as roughness length or stability function, a Function-valued
# transfer coefficient, a field-valued emissivity …) raises `ArgumentError` at the first step: there is NO CPU fallback and
# no KernelAbstractions dispatch behind these methods (BASELINE.json north_star).
#
# Activation (cf. ext/NumericalEarthReactantExt.jl:18-31, which switches behaviour by dispatch on the architecture parameter
# of EarthSystemModel):
#     [weakdeps]   CUDA = "052768ef-5323-5732-b1bb-66c8b64840ba"
#     [extensions] NumericalEarthB200Ext = "CUDA"
#     ENV["NE_B200_LIB"] = "/path/to/libne_b200.so"        # unset: the extension stays inert and the reference kernels run
#
# STATUS.  Julia is not installed in the build container, so this file has never been executed.  Its ABI half
# (ne_b200_abi.jl: every struct mirror, enum and #define) is GENERATED from the header by tools/gen_julia_abi.py and checked
# field by field (names, order, offsets, sizes) against the tested ctypes mirror by tests/test_julia_abi.py, which also checks
# that this file only uses structs / fields / constants / symbols that exist.  Every descriptor builder below has a line-by-line
# Python twin in numericalearth.jl_b200/{interface,formulations}.py that IS tested against the oracle on a B200.
module NumericalEarthB200Ext

using CUDA
using Oceananigans
using Oceananigans.Architectures: GPU, architecture
using Oceananigans.Fields: Field, ZeroField, ConstantField, AbstractField
using Oceananigans.Grids: halo_size, topology, Flat, λnodes, φnodes, AbstractGrid
using Oceananigans.OutputReaders: cpu_interpolating_time_indices, memory_index, FieldTimeSeries
using Oceananigans.Utils: KernelFunctionOperation
using Oceananigans.Simulations: Simulation
using ClimaSeaIce: SeaIceModel
using ClimaSeaIce.SeaIceThermodynamics: ConductiveFlux, IceSnowConductiveFlux, LinearLiquidus

using NumericalEarth
using NumericalEarth: EarthSystemModel
using NumericalEarth.EarthSystemModels
using NumericalEarth.EarthSystemModels: DegreesCelsius, DegreesKelvin, sea_ice_concentration, intercepted_snowfall,
                                        ocean_surface_temperature, ocean_temperature, ocean_salinity, boundary_layer_height,
                                        thermodynamics_parameters
using NumericalEarth.EarthSystemModels.InterfaceComputations
using NumericalEarth.EarthSystemModels.InterfaceComputations:
    ComponentExchanger, ComponentInterfaces, AtmosphereInterface, SeaIceOceanInterface, computed_fluxes, ZeroFluxes,
    SimilarityTheoryFluxes, CoefficientBasedFluxes, ConvergenceStopCriteria, FixedIterations, SimilarityScales,
    MomentumRoughnessLength, ScalarRoughnessLength, LandRoughnessLength, LandZeroPlaneDisplacement, ReynoldsScalingFunction,
    WindDependentWaveFormulation,
    TemperatureDependentAirViscosity, ConvectiveGustiness, SubgridVelocityCorrection,
    EdsonMomentumStabilityFunction, EdsonScalarStabilityFunction, ShebaMomentumStabilityFunction, ShebaScalarStabilityFunction,
    PaulsonMomentumStabilityFunction, PaulsonScalarStabilityFunction, LinearStableStabilityFunction, SplitStabilityFunction,
    LogarithmicSimilarityProfile, COARELogarithmicSimilarityProfile, PolynomialNeutralDragCoefficient,
    LargeYeagerTransferCoefficients, InterfaceProperties, ImpureSaturationSpecificHumidity, WaterMoleFraction,
    BulkTemperature, SkinTemperature, DiffusiveFlux, InteriorDiffusivity, RelativeVelocity, WindVelocity,
    IceBathHeatFlux, ThreeEquationHeatFlux, MomentumBasedFrictionVelocity, ElevationCorrection, NoSeaIceInterface
using NumericalEarth.Atmospheres: PrescribedAtmosphere, surface_rainfall_flux, surface_snowfall_flux
using NumericalEarth.Radiations: PrescribedRadiation, SurfaceRadiationProperties, LatitudeDependentAlbedo, TabulatedAlbedo, SeaIceAlbedo
using NumericalEarth.Oceans: forcing_barotropic_potential, get_radiative_forcing, TwoColorRadiation
using NumericalEarth.SeaIces: FreezingLimitedOceanTemperature
using Thermodynamics: Thermodynamics as AtmosphericThermodynamics

import NumericalEarth.EarthSystemModels: interpolate_state!, update_net_fluxes!, apply_air_sea_radiative_fluxes!,
                                         apply_air_sea_ice_radiative_fluxes!
import NumericalEarth.EarthSystemModels.InterfaceComputations: initialize!, correct_state!, compute_atmosphere_ocean_fluxes!,
                                                               compute_atmosphere_sea_ice_fluxes!, compute_sea_ice_ocean_fluxes!

include("ne_b200_abi.jl")     # GENERATED: Ne* structs, NE_* constants, NE_STRUCTS

#####
##### library handle, error convention, load-time ABI check
#####

const libne = Ref{String}("")
"true when NE_B200_LIB names a library that was loaded and whose ABI matches: the methods below then take over"
enabled() = !isempty(libne[])

function __init__()
    path = get(ENV, "NE_B200_LIB", "")
    isempty(path) && return nothing
    isfile(path) || error("NE_B200_LIB = $path does not exist")
    version = ccall((:ne_version, path), Cint, ())
    version == NE_ABI_VERSION || error("libne_b200 ABI version $version, this extension was generated for $NE_ABI_VERSION")
    for (name, T) in NE_STRUCTS
        n = ccall((:ne_struct_size, path), Int64, (Cstring,), name)
        n == sizeof(T) || error("libne_b200 ABI mismatch for $name: C sizeof = $n, Julia sizeof = $(sizeof(T))")
    end
    ccall((:ne_device_count, path), Cint, ()) > 0 || error("libne_b200 sees no CUDA device (there is no CPU fallback)")
    libne[] = path
    return nothing
end

function check(rc::Integer)
    rc == NE_OK && return nothing
    msg = unsafe_string(ccall((:ne_last_error, libne[]), Cstring, ()))
    rc == NE_E_NO_VARIANT && throw(ArgumentError("no sm_100a kernel variant: $msg (user closures are not supported; there is no CPU fallback)"))
    error("libne_b200 error $rc: $msg")
end

no_variant(what, x) = throw(ArgumentError("$what of type $(typeof(x)) has no sm_100a kernel variant (there is no CPU fallback)"))

"the caller's stream: CUDA.jl's task-local stream; every entry point only enqueues on it"
stream() = Ptr{Cvoid}(UInt(CUDA.stream().handle))

const GPUGrid = AbstractGrid{<:Any, <:Any, <:Any, <:Any, <:GPU}
const B200Model = EarthSystemModel{R, A, L, I, O, F, C, <:GPU} where {R, A, L, I, O, F, C}

suffix(grid) = eltype(grid) === Float64 ? "f64" : eltype(grid) === Float32 ? "f32" : no_variant("exchange grid element type", eltype(grid))
entry(base::String, grid) = Symbol(base, "_", suffix(grid))
dtype(::Type{Float64}) = NE_F64
dtype(::Type{Float32}) = NE_F32
dtype(T) = no_variant("element type", T)

#####
##### arrays: parent(field) device pointers, {array | constant} slots, exchange-grid layout
#####

devptr(a::CuArray) = Ptr{Cvoid}(UInt(pointer(a)))
devptr(a::AbstractArray) = devptr(parent(a))                       # OffsetArray → its CuArray parent
devptr(f::Field) = devptr(parent(f))
devptr(::Nothing) = Ptr{Cvoid}(C_NULL)

slot(f::Field) = NeSlot(ptr = devptr(f))
slot(a::AbstractArray) = NeSlot(ptr = devptr(a))
slot(::ZeroField) = NeSlot(value = 0.0)
slot(c::ConstantField) = NeSlot(value = Float64(c.constant))
slot(x::Number) = NeSlot(value = Float64(x))
slot(::Nothing) = NeSlot(value = 0.0)
slot(x) = no_variant("field-like object", x)

"pointer to the k = Nz plane of a 3-D field's parent: the ocean surface T, S, u, v (atmosphere_ocean_fluxes.jl:62-71)"
function surface_plane(f::Field)
    p = parent(f)
    Nz, Hz = size(f.grid, 3), halo_size(f.grid)[3]
    size(p, 3) == 1 && return slot(f)
    plane = size(p, 1) * size(p, 2)
    return NeSlot(ptr = Ptr{Cvoid}(UInt(pointer(p, 1 + (Nz + Hz - 1) * plane))))
end
surface_plane(x) = slot(x)

"launch range (0:Nx+1)x(0:Ny+1) of interface_kernel_parameters (InterfaceComputations.jl:100-116), or `:xy`"
function exchange_grid(grid; halo_ring::Bool)
    Nx, Ny, _ = size(grid)
    Hx, Hy, _ = halo_size(grid)
    TX, TY, _ = topology(grid)
    (TX() isa Flat || TY() isa Flat) && no_variant("exchange grid with a Flat horizontal direction", grid)
    return halo_ring ? NeExchangeGrid(nx = Nx, ny = Ny, hx = Hx, hy = Hy, i_lo = 0, i_hi = Nx + 1, j_lo = 0, j_hi = Ny + 1) :
                       NeExchangeGrid(nx = Nx, ny = Ny, hx = Hx, hy = Hy, i_lo = 1, i_hi = Nx, j_lo = 1, j_hi = Ny)
end

"uint8 exchange-layout mask of inactive_node(i, j, Nz) over the parent of a 2-D field, built once per grid (third-party grid logic)"
const inactive_masks = IdDict{Any, Any}()
function inactive_mask(grid)
    get!(inactive_masks, grid) do
        Nx, Ny, Nz = size(grid)
        Hx, Hy, _ = halo_size(grid)
        host = zeros(UInt8, Nx + 2Hx, Ny + 2Hy)
        cpu_grid = Oceananigans.on_architecture(Oceananigans.CPU(), grid)
        for j in 1-Hy:Ny+Hy, i in 1-Hx:Nx+Hx
            host[i + Hx, j + Hy] = Oceananigans.Grids.inactive_node(i, j, Nz, cpu_grid, Center(), Center(), Center())
        end
        CuArray(host)
    end
end
maskptr(grid) = Ptr{Cvoid}(UInt(pointer(inactive_mask(grid))))

#####
##### plugin types → POD variants
#####

pad12(v...) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 12)
pad4(v) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 4)

# stability functions (similarity_theory_turbulent_fluxes.jl:487-752); a `Returns(0)` profile is `nothing`-like :201-204
stability_fn(ψ::EdsonMomentumStabilityFunction) = NeStabilityFn(kind = NE_PSI_EDSON_MOMENTUM, p = pad12(ψ.ζmax, ψ.A⁺, ψ.B⁺, ψ.C⁺, ψ.D⁺, ψ.A⁻, ψ.B⁻, ψ.C⁻, ψ.D⁻, ψ.E⁻, ψ.F⁻))
stability_fn(ψ::EdsonScalarStabilityFunction) = NeStabilityFn(kind = NE_PSI_EDSON_SCALAR, p = pad12(ψ.ζmax, ψ.A⁺, ψ.B⁺, ψ.C⁺, ψ.D⁺, ψ.E⁺, ψ.A⁻, ψ.B⁻, ψ.C⁻, ψ.D⁻, ψ.E⁻, ψ.F⁻))
stability_fn(ψ::ShebaMomentumStabilityFunction) = NeStabilityFn(kind = NE_PSI_SHEBA_MOMENTUM, p = pad12(ψ.a, ψ.b))
stability_fn(ψ::ShebaScalarStabilityFunction) = NeStabilityFn(kind = NE_PSI_SHEBA_SCALAR, p = pad12(ψ.a, ψ.b, ψ.c))
stability_fn(ψ::PaulsonMomentumStabilityFunction) = NeStabilityFn(kind = NE_PSI_PAULSON_MOMENTUM, p = pad12(ψ.a, ψ.b))
stability_fn(ψ::PaulsonScalarStabilityFunction) = NeStabilityFn(kind = NE_PSI_PAULSON_SCALAR, p = pad12(ψ.a))
stability_fn(ψ::LinearStableStabilityFunction) = NeStabilityFn(kind = NE_PSI_LINEAR_STABLE, p = pad12(ψ.coefficient, ψ.maximum_stability_parameter))
stability_fn(ψ::Base.Returns) = iszero(ψ.value) ? NeStabilityFn(kind = NE_PSI_ZERO) : no_variant("constant non-zero stability function", ψ)
stability_fn(ψ) = no_variant("stability function", ψ)

stability_profile(ψ::SplitStabilityFunction) = NeStabilityProfile(split = Int32(1), a = stability_fn(ψ.stable), b = stability_fn(ψ.unstable))
stability_profile(ψ) = NeStabilityProfile(a = stability_fn(ψ))

stop_criteria(s::ConvergenceStopCriteria) = NeStopCriteria(kind = NE_STOP_CONVERGENCE, maxiter = Int32(s.maxiter), tolerance = Float64(s.tolerance))
stop_criteria(s::FixedIterations) = NeStopCriteria(kind = NE_STOP_FIXED_ITERATIONS, maxiter = Int32(s.iterations))
stop_criteria(s) = no_variant("solver_stop_criteria", s)

# roughness lengths (roughness_lengths.jl:1-19, 56-75, 93-139, 149-189, 212-246).  The constant air viscosity keeps ITS OWN
# element type (a Float64 literal by default, :94,126): that is what makes a Float32 model's iteration mixed precision.
function viscosity_fields(ν::Number)
    return (visc_kind = NE_VISC_CONSTANT, visc_dtype = dtype(typeof(ν)), nu = Float64(ν), nu_C = pad4(()))
end
function viscosity_fields(ν::TemperatureDependentAirViscosity)
    return (visc_kind = NE_VISC_TEMPERATURE_DEPENDENT, visc_dtype = dtype(typeof(ν.ℂ₀)), nu = 0.0, nu_C = pad4((ν.ℂ₀, ν.ℂ₁, ν.ℂ₂, ν.ℂ₃)))
end
viscosity_fields(ν) = no_variant("air_kinematic_viscosity", ν)

wave_fields(ℂg::Number) = (wave_kind = NE_WAVE_CONSTANT, wave_constant = Float64(ℂg), wave_Umax = 0.0, wave_C1 = 0.0, wave_C2 = 0.0)
wave_fields(w::WindDependentWaveFormulation) = (wave_kind = NE_WAVE_WIND_DEPENDENT, wave_constant = 0.0, wave_Umax = Float64(w.Umax), wave_C1 = Float64(w.ℂ₁), wave_C2 = Float64(w.ℂ₂))
wave_fields(w) = no_variant("wave_formulation", w)

roughness_length(ℓ::Number; land = false) = NeRoughnessLength(kind = NE_ROUGH_CONSTANT, constant = Float64(ℓ), visc_dtype = NE_F64)
function roughness_length(ℓ::MomentumRoughnessLength; land = false)
    v, w = viscosity_fields(ℓ.air_kinematic_viscosity), wave_fields(ℓ.wave_formulation)
    return NeRoughnessLength(kind = NE_ROUGH_MOMENTUM, wave_kind = w.wave_kind, visc_kind = v.visc_kind, visc_dtype = v.visc_dtype,
                             gravitational_acceleration = Float64(ℓ.gravitational_acceleration), wave_constant = w.wave_constant,
                             smooth_wall_parameter = Float64(ℓ.smooth_wall_parameter), wave_Umax = w.wave_Umax, wave_C1 = w.wave_C1,
                             wave_C2 = w.wave_C2, maximum_roughness_length = Float64(ℓ.maximum_roughness_length), nu = v.nu, nu_C = v.nu_C)
end
function roughness_length(ℓ::ScalarRoughnessLength; land = false)
    v = viscosity_fields(ℓ.air_kinematic_viscosity)
    s = ℓ.reynolds_number_scaling_function
    s isa ReynoldsScalingFunction || no_variant("reynolds_number_scaling_function", s)
    return NeRoughnessLength(kind = NE_ROUGH_SCALAR, visc_kind = v.visc_kind, visc_dtype = v.visc_dtype, nu = v.nu, nu_C = v.nu_C,
                             maximum_roughness_length = Float64(ℓ.maximum_roughness_length), reynolds_A = Float64(s.A), reynolds_b = Float64(s.b))
end
# LandRoughnessLength (roughness_lengths.jl:21-39): a marker the atmosphere–land kernel resolves per cell from the land model's
# roughness field (NeAtmosLandDesc.momentum_roughness_length / scalar_roughness_length).  Over the ocean and sea ice the interior
# properties carry no such field and local_roughness_length returns max(multiplier * minimum, minimum)
# (similarity_theory_turbulent_fluxes.jl:265-278): `land = false` passes that constant.
function roughness_length(ℓ::LandRoughnessLength{T}; land = false) where T
    land && return NeRoughnessLength(kind = NE_ROUGH_LAND, visc_dtype = NE_F64, land_multiplier = Float64(ℓ.multiplier),
                                     land_minimum_roughness_length = Float64(ℓ.minimum_roughness_length))
    return roughness_length(max(ℓ.multiplier * ℓ.minimum_roughness_length, ℓ.minimum_roughness_length))
end
roughness_length(ℓ; land = false) = no_variant("roughness length", ℓ)     # closures ℓ(u★, …) (roughness_lengths.jl:192)

# subgrid velocities (similarity_theory_turbulent_fluxes.jl:45-98)
sgs_slot(::Nothing) = (NE_SGS_NONE, 0.0)
sgs_slot(v::Number) = (NE_SGS_CONSTANT, Float64(v))
sgs_slot(g::ConvectiveGustiness) = (NE_SGS_CONVECTIVE, 0.0)
sgs_slot(x) = no_variant("subgrid velocity", x)
function subgrid_velocity(sv)
    if sv isa SubgridVelocityCorrection
        ck, cc = sgs_slot(sv.convective)
        mk, mc = sgs_slot(sv.mesoscale)
        mk == NE_SGS_CONVECTIVE && no_variant("mesoscale subgrid velocity", sv.mesoscale)
        g = sv.convective isa ConvectiveGustiness ? sv.convective : ConvectiveGustiness{Float64}()
        return NeSubgridVelocity(convective_kind = ck, mesoscale_kind = mk, composite = Int32(1), gustiness_parameter = Float64(g.gustiness_parameter),
                                 minimum_gustiness = Float64(g.minimum_gustiness), convective_constant = cc, mesoscale_constant = mc)
    end
    k, c = sgs_slot(sv)
    g = sv isa ConvectiveGustiness ? sv : ConvectiveGustiness{Float64}()
    return NeSubgridVelocity(convective_kind = k, gustiness_parameter = Float64(g.gustiness_parameter),
                             minimum_gustiness = Float64(g.minimum_gustiness), convective_constant = c)
end

polynomial_drag(p::PolynomialNeutralDragCoefficient) =
    NePolynomialDrag(a = Float64(p.a), b = Float64(p.b), c = Float64(p.c), d = Float64(p.d),
                     high_wind_speed_threshold = Float64(p.high_wind_speed_threshold),
                     high_wind_drag_coefficient = Float64(p.high_wind_drag_coefficient), minimum_wind_speed = Float64(p.minimum_wind_speed))

transfer_coefficient(c::Number) = NeTransferCoefficient(kind = NE_COEFF_CONSTANT, constant = Float64(c))
transfer_coefficient(c::PolynomialNeutralDragCoefficient) = NeTransferCoefficient(kind = NE_COEFF_POLYNOMIAL_DRAG, polynomial = polynomial_drag(c))
transfer_coefficient(c) = no_variant("transfer coefficient", c)      # Function-valued coefficients (coefficient_based_turbulent_fluxes.jl:270)

similarity_form(::LogarithmicSimilarityProfile) = NE_PROFILE_LOGARITHMIC
similarity_form(::COARELogarithmicSimilarityProfile) = NE_PROFILE_COARE
similarity_form(f) = no_variant("similarity_form", f)

# `land = true` (the atmosphere–land descriptor) keeps the per-cell land markers; elsewhere they collapse to constants
function flux_formulation(f::SimilarityTheoryFluxes; land = false)
    ψ, ℓ = f.stability_functions, f.roughness_lengths
    d = f.zero_plane_displacement
    d_kind = NE_DISPLACEMENT_CONSTANT
    if d isa LandZeroPlaneDisplacement        # local_zero_plane_displacement (:296-303): 0 without a land field
        d, d_kind = 0, land ? NE_DISPLACEMENT_LAND : NE_DISPLACEMENT_CONSTANT
    end
    d isa Number || no_variant("zero_plane_displacement", d)
    return NeFluxFormulation(kind = NE_FLUX_SIMILARITY_THEORY, similarity_form = similarity_form(f.similarity_form),
                             von_karman_constant = Float64(f.von_karman_constant), subgrid_velocities = subgrid_velocity(f.subgrid_velocities),
                             psi_momentum = stability_profile(ψ.momentum), psi_temperature = stability_profile(ψ.temperature),
                             psi_water_vapor = stability_profile(ψ.water_vapor),
                             ell_momentum = roughness_length(ℓ.momentum; land), ell_temperature = roughness_length(ℓ.temperature; land),
                             ell_water_vapor = roughness_length(ℓ.water_vapor; land), zero_plane_displacement = Float64(d),
                             zero_plane_displacement_kind = d_kind,
                             stop = stop_criteria(f.solver_stop_criteria))
end

function flux_formulation(f::CoefficientBasedFluxes; land = false)
    c = f.transfer_coefficients
    if c isa LargeYeagerTransferCoefficients       # coefficient_based_turbulent_fluxes.jl:82-108, 288-340
        ψ = c.stability_functions
        ly = NeLargeYeager(von_karman_constant = Float64(c.von_karman_constant), neutral_drag = polynomial_drag(c.neutral_drag_coefficient),
                           psi_momentum = stability_profile(ψ.momentum), psi_temperature = stability_profile(ψ.temperature),
                           reference_height = Float64(c.reference_height), stable_heat = Float64(c.stable_heat_transfer_coefficient),
                           unstable_heat = Float64(c.unstable_heat_transfer_coefficient), moisture = Float64(c.moisture_transfer_coefficient))
        return NeFluxFormulation(kind = NE_FLUX_LARGE_YEAGER, large_yeager = ly, stop = stop_criteria(f.solver_stop_criteria))
    end
    c isa SimilarityScales || no_variant("transfer_coefficients", c)
    coefficients = (transfer_coefficient(c.momentum), transfer_coefficient(c.temperature), transfer_coefficient(c.water_vapor))
    return NeFluxFormulation(kind = NE_FLUX_COEFFICIENT_BASED, coefficients = coefficients, stop = stop_criteria(f.solver_stop_criteria))
end
flux_formulation(f; land = false) = no_variant("flux formulation", f)

# InterfaceProperties (interface_states.jl:8-12, 20-74, 236-277, 284-301, 330-398)
phase_of(::AtmosphericThermodynamics.Liquid) = NE_PHASE_LIQUID
phase_of(::AtmosphericThermodynamics.Ice) = NE_PHASE_ICE
phase_of(p) = no_variant("thermodynamic phase", p)

function interface_properties(ip::InterfaceProperties)
    q = ip.specific_humidity_formulation
    q isa ImpureSaturationSpecificHumidity || no_variant("specific_humidity_formulation", q)
    x = q.water_mole_fraction
    xk, xv, wm = NE_XH2O_ONE, 0.0, 0.0
    mm, mf = pad4(()), pad4(())
    if x isa Number
        xk, xv = NE_XH2O_CONSTANT, Float64(x)
    elseif x isa WaterMoleFraction
        xk, wm = NE_XH2O_SALINITY, Float64(x.water_molar_mass)
        cs = values(x.salinity_constituents)
        length(cs) <= 4 || no_variant("WaterMoleFraction with more than four constituents", x)
        mm, mf = pad4(map(c -> c.molar_mass, cs)), pad4(map(c -> c.mass_fraction, cs))
    elseif !isnothing(x)
        no_variant("water_mole_fraction", x)
    end
    v = ip.velocity_formulation
    vk = v isa RelativeVelocity ? NE_VEL_RELATIVE : v isa WindVelocity ? NE_VEL_WIND : no_variant("velocity_formulation", v)
    t = ip.temperature_formulation
    tk, max_dT, κ, δ, ki, ks = NE_TEMP_BULK, 0.0, 0.0, 0.0, 0.0, 0.0
    if t isa SkinTemperature
        max_dT = Float64(t.max_ΔT)
        F = t.internal_flux
        if F isa DiffusiveFlux
            δ = Float64(F.δ)
            if F.κ isa InteriorDiffusivity
                tk, κ = NE_TEMP_SKIN_DIFFUSIVE_INTERIOR, Float64(F.κ.minimum_diffusivity)
            elseif F.κ isa Number
                tk, κ = NE_TEMP_SKIN_DIFFUSIVE, Float64(F.κ)
            else
                no_variant("DiffusiveFlux diffusivity", F.κ)
            end
        elseif F isa ConductiveFlux
            tk, ki = NE_TEMP_SKIN_CONDUCTIVE, Float64(F.conductivity)
        elseif F isa IceSnowConductiveFlux
            tk, ki, ks = NE_TEMP_SKIN_ICE_SNOW, Float64(F.ice_conductivity), Float64(F.snow_conductivity)
        else
            no_variant("SkinTemperature internal_flux", F)
        end
    elseif !(t isa BulkTemperature)
        no_variant("temperature_formulation", t)
    end
    return NeInterfaceProperties(phase = phase_of(q.phase), x_h2o_kind = xk, velocity_formulation = vk, temperature_formulation = tk,
                                 x_h2o = xv, water_molar_mass = wm, constituent_molar_mass = mm, constituent_mass_fraction = mf,
                                 max_dT = max_dT, kappa = κ, delta = δ, ice_conductivity = ki, snow_conductivity = ks)
end

# AtmosphereThermodynamicsParameters (src/Atmospheres/thermodynamic_parameters.jl:30-258)
function thermo_params(ℂ)
    c, h, p = ℂ.constitutive, ℂ.heat_capacity, ℂ.phase_transitions
    return NeThermoParams(dtype = dtype(eltype(ℂ)), gas_constant = Float64(c.gas_constant), dry_air_molar_mass = Float64(c.dry_air_molar_mass),
                          water_molar_mass = Float64(c.water_molar_mass), kappa_d = Float64(h.dry_air_adiabatic_exponent),
                          cp_v = Float64(h.water_vapor_heat_capacity), cp_l = Float64(h.liquid_water_heat_capacity),
                          cp_i = Float64(h.water_ice_heat_capacity), LH_v0 = Float64(p.reference_vaporization_enthalpy),
                          LH_s0 = Float64(p.reference_sublimation_enthalpy), T_0 = Float64(p.reference_temperature),
                          T_triple = Float64(p.triple_point_temperature), press_triple = Float64(p.triple_point_pressure),
                          T_freeze = Float64(p.water_freezing_temperature), T_icenuc = Float64(p.total_ice_nucleation_temperature))
end

units_of(::DegreesCelsius) = NE_DEGREES_CELSIUS
units_of(::DegreesKelvin) = NE_DEGREES_KELVIN
units_of(u) = no_variant("temperature_units", u)

function medium_properties(p; liquidus = nothing)
    isnothing(p) && return NeMediumProperties()
    L = isnothing(liquidus) ? (hasproperty(p, :liquidus) ? p.liquidus : LinearLiquidus(Float64)) : liquidus
    L isa LinearLiquidus || no_variant("liquidus", L)
    return NeMediumProperties(reference_density = Float64(p.reference_density), heat_capacity = Float64(p.heat_capacity),
                              temperature_units = units_of(p.temperature_units), liquidus_slope = Float64(L.slope),
                              liquidus_freshwater_melting_temperature = Float64(L.freshwater_melting_temperature))
end

#####
##### radiation properties of one surface (src/Radiations/air_sea_interface_radiation_state.jl:4-39)
#####

const node_arrays = IdDict{Any, Any}()
"exchange nodes with halos as device vectors (λ, φ) in the exchange element type, built once per grid"
function exchange_nodes(grid)
    get!(node_arrays, grid) do
        λ = CuArray(collect(eltype(grid), parent(λnodes(grid, Center(); with_halos = true))))
        φ = CuArray(collect(eltype(grid), parent(φnodes(grid, Center(); with_halos = true))))
        ndims(λ) == 1 && ndims(φ) == 1 || no_variant("curvilinear exchange grid without 1-D node axes (pass nodes_2d arrays)", grid)
        (λ, φ)
    end
end

function surface_radiation(model, surface::Symbol)
    radiation = model.radiation
    (isnothing(radiation) || !haskey(radiation.surface_properties, surface)) && return NeSurfaceRadiation()   # enabled = 0: zero radiation state
    grid = model.interfaces.exchanger.grid
    s = radiation.surface_properties[surface]
    state = model.interfaces.exchanger.radiation.state
    s.emissivity isa Number || no_variant("field-valued emissivity", s.emissivity)
    α = s.albedo
    kind, a0, a1 = NE_ALBEDO_CONSTANT, 0.0, 0.0
    field, lat = Ptr{Cvoid}(C_NULL), Ptr{Cvoid}(C_NULL)
    ice, tab = NeSeaIceAlbedo(), NeTabulatedAlbedo()
    if α isa Number
        a0 = Float64(α)
    elseif α isa LatitudeDependentAlbedo                                   # latitude_dependent_albedo.jl:48-53
        kind, a0, a1 = NE_ALBEDO_LATITUDE_DEPENDENT, Float64(α.diffuse), Float64(α.direct)
        lat = devptr(exchange_nodes(grid)[2])
    elseif α isa SeaIceAlbedo                                              # sea_ice_albedo.jl:22-133
        kind = NE_ALBEDO_SEA_ICE
        ice = NeSeaIceAlbedo(ice_albedo = Float64(α.ice_albedo), snow_albedo = Float64(α.snow_albedo),
                             ice_melt_reduction = Float64(α.ice_melt_reduction), snow_melt_reduction = Float64(α.snow_melt_reduction),
                             melting_temperature = Float64(α.melting_temperature), temperature_range = Float64(α.temperature_range),
                             ocean_albedo = Float64(α.ocean_albedo), minimum_ice_thickness = Float64(α.minimum_ice_thickness),
                             minimum_snow_depth = Float64(α.minimum_snow_depth), ice_thickness = devptr(α.ice_thickness),
                             snow_thickness = devptr(α.snow_thickness), surface_temperature = devptr(α.surface_temperature))
    elseif α isa TabulatedAlbedo                                           # tabulated_albedo.jl:39-160: clock scalars on the host
        kind = NE_ALBEDO_TABULATED
        t = Float64(model.clock.time isa Number ? model.clock.time : 0)
        day = trunc(t / 86400)
        δ = deg2rad((23 + 27 / 60) * sind(360 * (day - 80) / 365.25))
        λ, φ = exchange_nodes(grid)
        lat = devptr(φ)
        FT = eltype(grid)
        tab = NeTabulatedAlbedo(table = devptr(α.α_table), n_t = Int32(size(α.α_table, 1)), n_phi = Int32(size(α.α_table, 2)),
                                t_values = (Float64(FT(α.𝓉_values[1])), Float64(FT(α.𝓉_values[2]))),
                                phi_values = (Float64(FT(α.φ_values[1])), Float64(FT(α.φ_values[2]))),
                                solar_constant = Float64(α.S₀), day_to_radians = Float64(FT(α.day_to_radians)),
                                noon_in_seconds = Float64(α.noon_in_seconds), seconds_in_day = t - day * 86400,
                                declination = Float64(FT(δ)), longitude = devptr(λ))
    elseif α isa Field
        kind, field = NE_ALBEDO_FIELD, devptr(α)
    else
        no_variant("albedo", α)
    end
    return NeSurfaceRadiation(enabled = Int32(1), albedo_kind = kind, stefan_boltzmann_constant = Float64(radiation.stefan_boltzmann_constant),
                              albedo = a0, albedo_direct = a1, albedo_field = field, latitude = lat, emissivity = Float64(s.emissivity),
                              downwelling_shortwave = devptr(state.ℐꜜˢʷ), downwelling_longwave = devptr(state.ℐꜜˡʷ),
                              sea_ice_albedo = ice, tabulated_albedo = tab)
end

#####
##### phase 0: fractional indices (prescribed_atmosphere_regridder.jl:41-71, prescribed_radiation_regridder.jl:23-52)
#####
# Off by default: the reference's own `initialize!` then writes the indices and they are bit-exact by construction.
# ENV["NE_B200_FRAC_INDICES"] = "1" computes them with ne_frac_indices_* instead (bit-exact against the oracle's restatement
# of Oceananigans' FractionalIndices; LatitudeLongitudeGrid sources only).

use_library_frac_indices() = enabled() && get(ENV, "NE_B200_FRAC_INDICES", "0") == "1"

function frac_indices!(frac, grid, source_grid)
    source_grid isa Oceananigans.Grids.LatitudeLongitudeGrid || no_variant("source grid of a prescribed component", source_grid)
    λ, φ = exchange_nodes(grid)
    FTa = eltype(source_grid)
    λs = CuArray(collect(FTa, λnodes(source_grid, Center())))
    φs = CuArray(collect(FTa, φnodes(source_grid, Center())))
    regular(x) = x isa Number || x isa AbstractRange
    d = NeFracIndexDesc(grid = exchange_grid(grid; halo_ring = true), nodes_2d = Int32(0), lam = devptr(λ), phi = devptr(φ),
                        src_dtype = dtype(FTa), src_x_regular = Int32(regular(source_grid.Δλᶜᵃᵃ)), src_y_regular = Int32(regular(source_grid.Δφᵃᶜᵃ)),
                        src_nx = size(source_grid, 1), src_ny = size(source_grid, 2), src_lam_nodes = devptr(λs), src_phi_nodes = devptr(φs),
                        frac_i = devptr(frac.i), frac_j = devptr(frac.j))
    GC.@preserve λs φs check(ccall((entry("ne_frac_indices", grid), libne[]), Cint, (Ref{NeFracIndexDesc}, Ptr{Cvoid}), d, stream()))
    CUDA.synchronize()     # λs, φs are temporaries
    return nothing
end

function initialize!(exchanger::ComponentExchanger, grid::GPUGrid, atmosphere::PrescribedAtmosphere)
    use_library_frac_indices() || return invoke(initialize!, Tuple{ComponentExchanger, Any, PrescribedAtmosphere}, exchanger, grid, atmosphere)
    return frac_indices!(exchanger.regridder, grid, atmosphere.grid)
end

function initialize!(exchanger::ComponentExchanger, grid::GPUGrid, radiation::PrescribedRadiation)
    use_library_frac_indices() || return invoke(initialize!, Tuple{ComponentExchanger, Any, PrescribedRadiation}, exchanger, grid, radiation)
    return frac_indices!(exchanger.regridder, grid, radiation.grid)
end

#####
##### phase 1: interpolation (interpolate_atmospheric_state.jl:9-86, interpolate_radiation_state.jl:4-41)
#####

"host-side TimeInterpolator exactly as the reference computes it (:57-60), mapped to in-memory slices"
function time_interp(fts::FieldTimeSeries, arch, t)
    ti = cpu_interpolating_time_indices(arch, fts.times, fts.time_indexing, t)
    n₁, n₂ = ti.first_index, ti.second_index
    ñ = ti.fractional_index
    return NeTimeInterp(frac = Float64(ñ), frac_dtype = dtype(typeof(ñ)), m1 = Int32(memory_index(fts, n₁)), m2 = Int32(memory_index(fts, n₂)),
                        same = Int32(n₁ == n₂))
end

series_tuple(::Nothing) = ()
series_tuple(x::Tuple) = x
series_tuple(x::NamedTuple) = values(x)
series_tuple(x) = (x,)

"one NeInterpDesc for `fields` = vector of (tuple of FTS data arrays summed into one output, output Field)"
function interp_descriptor(grid, frac, source::FieldTimeSeries, time, fields; potential = nothing, ρᵒᶜ = 0.0)
    length(fields) <= 9 || error("at most 9 output fields per interpolation call")
    sg = source.grid
    Hx, Hy, _ = halo_size(sg)
    n_summands = ntuple(f -> f <= length(fields) ? Int32(length(fields[f][1])) : Int32(0), 9)
    all(n -> n <= NE_MAX_SUMMANDS, n_summands) || no_variant("precipitation tuple with more than four summands", fields)
    series = ntuple(9) do f
        ntuple(NE_MAX_SUMMANDS) do k
            (f <= length(fields) && k <= length(fields[f][1])) ? NeTimeSeries(data = devptr(fields[f][1][k])) : NeTimeSeries()
        end
    end
    out = ntuple(f -> f <= length(fields) ? devptr(fields[f][2]) : Ptr{Cvoid}(C_NULL), 9)
    return NeInterpDesc(grid = exchange_grid(grid; halo_ring = true), frac_i = devptr(frac.i), frac_j = devptr(frac.j),
                        src_dtype = dtype(eltype(sg)), n_fields = Int32(length(fields)), src_nx = size(sg, 1), src_ny = size(sg, 2),
                        src_hx = Hx, src_hy = Hy, src_nt = size(parent(source.data), 4), time = time, n_summands = n_summands,
                        series = series, out = out, potential = devptr(potential), potential_from = Int32(4),
                        ocean_reference_density = Float64(ρᵒᶜ))
end

function interpolate_state!(exchanger, grid, atmosphere::PrescribedAtmosphere, model::B200Model)
    enabled() || return invoke(interpolate_state!, Tuple{Any, Any, PrescribedAtmosphere, Any}, exchanger, grid, atmosphere, model)
    grid isa Oceananigans.Grids.LatitudeLongitudeGrid || grid isa Oceananigans.Grids.RectilinearGrid ||
        no_variant("rotated exchange grid (fill NeInterpDesc.rotation_cos/sin from the grid's rotation metrics first)", grid)
    u = atmosphere.velocities.u
    st = exchanger.state
    time = time_interp(u, architecture(grid), model.clock.time)
    fields = [((atmosphere.velocities.u.data,), st.u), ((atmosphere.velocities.v.data,), st.v),
              ((atmosphere.temperature.data,), st.T), ((atmosphere.specific_humidity.data,), st.q),
              ((atmosphere.pressure.data,), st.p),
              (series_tuple(surface_rainfall_flux(atmosphere)), st.Jʳⁿ), (series_tuple(surface_snowfall_flux(atmosphere)), st.Jˢⁿ)]
    potential = forcing_barotropic_potential(model.ocean)           # :80-85, written by the same kernel
    d = interp_descriptor(grid, exchanger.regridder, u, time, fields; potential,
                          ρᵒᶜ = isnothing(potential) ? 0.0 : model.interfaces.ocean_properties.reference_density)
    GC.@preserve atmosphere exchanger check(ccall((entry("ne_interp_state", grid), libne[]), Cint, (Ref{NeInterpDesc}, Ptr{Cvoid}), d, stream()))
    return nothing
end

function interpolate_state!(exchanger, grid, radiation::PrescribedRadiation, model::B200Model)
    enabled() || return invoke(interpolate_state!, Tuple{Any, Any, PrescribedRadiation, Any}, exchanger, grid, radiation, model)
    sw, lw = radiation.downwelling_shortwave, radiation.downwelling_longwave
    time = time_interp(sw, architecture(grid), model.clock.time)
    fields = [((sw.data,), exchanger.state.ℐꜜˢʷ), ((lw.data,), exchanger.state.ℐꜜˡʷ)]
    d = interp_descriptor(grid, exchanger.regridder, sw, time, fields)
    GC.@preserve radiation exchanger check(ccall((entry("ne_interp_state", grid), libne[]), Cint, (Ref{NeInterpDesc}, Ptr{Cvoid}), d, stream()))
    return nothing
end

#####
##### phase 1.5: ElevationCorrection (atmosphere_state_correction.jl:122-146)
#####

function correct_state!(correction::ElevationCorrection, exchanger, grid::GPUGrid)
    enabled() || return invoke(correct_state!, Tuple{ElevationCorrection, Any, Any}, correction, exchanger, grid)
    d = NeElevationCorrectionDesc(grid = exchange_grid(grid; halo_ring = true), T = devptr(exchanger.state.T), p = devptr(exchanger.state.p),
                                  elevation_difference = devptr(correction.elevation_difference), lapse_rate = Float64(correction.lapse_rate),
                                  gravitational_acceleration = Float64(correction.gravitational_acceleration),
                                  dry_air_gas_constant = Float64(correction.dry_air_gas_constant))
    GC.@preserve correction exchanger check(ccall((entry("ne_correct_atmosphere_elevation", grid), libne[]), Cint,
                                                  (Ref{NeElevationCorrectionDesc}, Ptr{Cvoid}), d, stream()))
    return nothing
end

#####
##### phase 2: turbulent fluxes
#####

height_slot(h::Number) = NeSlot(value = Float64(h))
height_slot(h) = slot(h)

"the interior diffusivity plane of InteriorDiffusivity (interface_states.jl:384-391): the ocean's κ operation, computed into a field"
const kappa_fields = IdDict{Any, Any}()
function interior_diffusivity(model, properties)
    properties.temperature_formulation isa SkinTemperature || return Ptr{Cvoid}(C_NULL)
    F = properties.temperature_formulation.internal_flux
    (F isa DiffusiveFlux && F.κ isa InteriorDiffusivity) || return Ptr{Cvoid}(C_NULL)
    κop = model.interfaces.exchanger.ocean.state.κ
    κ = get!(() -> Field(κop), kappa_fields, κop)
    Oceananigans.compute!(κ)                      # Oceananigans' own kernel (closure-dependent, third party)
    return devptr(κ)
end

function compute_atmosphere_ocean_fluxes!(model::B200Model)
    enabled() || return invoke(compute_atmosphere_ocean_fluxes!, Tuple{Any}, model)
    interface = model.interfaces.atmosphere_ocean_interface
    isnothing(interface) && return nothing
    exchanger = model.interfaces.exchanger
    grid = exchanger.grid
    a, o, f = exchanger.atmosphere.state, exchanger.ocean.state, interface.fluxes
    d = NeAtmosOceanDesc(grid = exchange_grid(grid; halo_ring = true),
                         ua = devptr(a.u), va = devptr(a.v), Ta = devptr(a.T), pa = devptr(a.p), qa = devptr(a.q),
                         surface_layer_height = height_slot(model.interfaces.properties.surface_layer_height),
                         boundary_layer_height = height_slot(boundary_layer_height(model.atmosphere)),
                         uo = surface_plane(o.u), vo = surface_plane(o.v), To = surface_plane(o.T), So = surface_plane(o.S),
                         kappa = interior_diffusivity(model, interface.properties), inactive = maskptr(grid),
                         radiation = surface_radiation(model, :ocean), thermo = thermo_params(thermodynamics_parameters(model.atmosphere)),
                         gravitational_acceleration = Float64(model.interfaces.properties.gravitational_acceleration),
                         flux = flux_formulation(interface.flux_formulation), properties = interface_properties(interface.properties),
                         ocean = medium_properties(model.interfaces.ocean_properties),
                         latent_heat = devptr(f.latent_heat), sensible_heat = devptr(f.sensible_heat), water_vapor = devptr(f.water_vapor),
                         x_momentum = devptr(f.x_momentum), y_momentum = devptr(f.y_momentum), interface_temperature = devptr(interface.temperature),
                         friction_velocity = devptr(f.friction_velocity), temperature_scale = devptr(f.temperature_scale),
                         water_vapor_scale = devptr(f.water_vapor_scale))
    GC.@preserve model check(ccall((entry("ne_atmosphere_ocean_fluxes", grid), libne[]), Cint, (Ref{NeAtmosOceanDesc}, Ptr{Cvoid}), d, stream()))
    return nothing
end

function compute_atmosphere_sea_ice_fluxes!(model::B200Model)
    enabled() || return invoke(compute_atmosphere_sea_ice_fluxes!, Tuple{Any}, model)
    interface = model.interfaces.atmosphere_sea_ice_interface
    isnothing(interface) && return nothing
    exchanger = model.interfaces.exchanger
    grid = exchanger.grid
    a, o, s, f = exchanger.atmosphere.state, exchanger.ocean.state, exchanger.sea_ice.state, interface.fluxes
    d = NeAtmosSeaIceDesc(grid = exchange_grid(grid; halo_ring = true),
                          ua = devptr(a.u), va = devptr(a.v), Ta = devptr(a.T), pa = devptr(a.p), qa = devptr(a.q),
                          surface_layer_height = height_slot(model.interfaces.properties.surface_layer_height),
                          boundary_layer_height = height_slot(boundary_layer_height(model.atmosphere)),
                          To = surface_plane(o.T), So = surface_plane(o.S),
                          hi = slot(s.hi), hs = slot(s.hs), hc = slot(s.hc), concentration = slot(s.ℵ), inactive = maskptr(grid),
                          radiation = surface_radiation(model, :sea_ice), thermo = thermo_params(thermodynamics_parameters(model.atmosphere)),
                          gravitational_acceleration = Float64(model.interfaces.properties.gravitational_acceleration),
                          flux = flux_formulation(interface.flux_formulation), properties = interface_properties(interface.properties),
                          ocean = medium_properties(model.interfaces.ocean_properties),
                          sea_ice = medium_properties(model.interfaces.sea_ice_properties),
                          latent_heat = devptr(f.latent_heat), sensible_heat = devptr(f.sensible_heat), water_vapor = devptr(f.water_vapor),
                          x_momentum = devptr(f.x_momentum), y_momentum = devptr(f.y_momentum),
                          interface_temperature = devptr(interface.temperature))      # READ-MODIFY-WRITE: top_surface_temperature
    GC.@preserve model check(ccall((entry("ne_atmosphere_sea_ice_fluxes", grid), libne[]), Cint, (Ref{NeAtmosSeaIceDesc}, Ptr{Cvoid}), d, stream()))
    return nothing
end

# The reference's own no-interface / FreezingLimited specialisations (EarthSystemModels.jl:86-113,
# freezing_limited_ocean_temperature.jl:70-73) and the methods above are specialised on different type parameters of
# EarthSystemModel: their intersections are spelled out so that dispatch stays unambiguous.
const B200NoSeaIceInterfaceModel = EarthSystemModel{R, A, L, I, O, <:NoSeaIceInterface, C, <:GPU} where {R, A, L, I, O, C}
const B200FreezingLimitedModel = EarthSystemModel{R, A, L, <:FreezingLimitedOceanTemperature, O, <:NoSeaIceInterface, C, <:GPU} where {R, A, L, O, C}
const B200NoRadiationModel = EarthSystemModel{<:Nothing, A, L, I, O, F, C, <:GPU} where {A, L, I, O, F, C}
for NoX in (:NoOceanInterface, :NoAtmosInterface, :NoInterface)
    T = :(EarthSystemModel{R, A, L, I, O, <:NumericalEarth.EarthSystemModels.$NoX, C, <:GPU} where {R, A, L, I, O, C})
    @eval compute_atmosphere_ocean_fluxes!(::$T) = nothing
    @eval compute_atmosphere_sea_ice_fluxes!(::$T) = nothing
    @eval compute_sea_ice_ocean_fluxes!(::$T) = nothing
end
compute_atmosphere_sea_ice_fluxes!(::B200NoSeaIceInterfaceModel) = nothing
compute_atmosphere_sea_ice_fluxes!(::B200FreezingLimitedModel) = nothing
compute_sea_ice_ocean_fluxes!(::B200NoSeaIceInterfaceModel) = nothing
apply_air_sea_radiative_fluxes!(::B200NoRadiationModel) = nothing
apply_air_sea_ice_radiative_fluxes!(::B200NoRadiationModel) = nothing

friction_velocity_fields(u★::Number) = (NE_USTAR_CONSTANT, Float64(u★))
friction_velocity_fields(::MomentumBasedFrictionVelocity) = (NE_USTAR_MOMENTUM_BASED, 0.0)
friction_velocity_fields(u★) = no_variant("friction_velocity", u★)

"Δzᶜᶜᶜ(k), k = 1..Nz, as a device vector in the exchange element type (sea_ice_ocean_fluxes.jl:157), built once per grid"
const dz_vectors = IdDict{Any, Any}()
function cell_thicknesses(grid)
    get!(dz_vectors, grid) do
        cpu_grid = Oceananigans.on_architecture(Oceananigans.CPU(), grid)
        Δz = [Oceananigans.Operators.Δzᶜᶜᶜ(1, 1, k, cpu_grid) for k in 1:size(grid, 3)]      # z-star / partial cells: see INTEGRATION.md
        CuArray(convert(Vector{eltype(grid)}, Δz))
    end
end

function compute_sea_ice_ocean_fluxes!(model::B200Model)
    enabled() || return invoke(compute_sea_ice_ocean_fluxes!, Tuple{Any}, model)
    interface = model.interfaces.sea_ice_ocean_interface
    isnothing(interface) && return nothing
    ocean, sea_ice = model.ocean, model.sea_ice
    grid = sea_ice.model.grid
    fluxes, ff = interface.fluxes, interface.flux_formulation
    Tᵒᶜ, Sᵒᶜ = ocean_temperature(ocean), ocean_salinity(ocean)
    uˢⁱ, vˢⁱ = sea_ice.model.velocities
    ρᵒᶜ = model.interfaces.ocean_properties.reference_density
    dynamics = sea_ice.model.dynamics
    if !isnothing(dynamics)      # _compute_sea_ice_ocean_stress! (:79-104): ClimaSeaIce SemiImplicitStress
        τₛ = dynamics.external_momentum_stresses.bottom
        hasproperty(τₛ, :ρₑ) && hasproperty(τₛ, :Cᴰ) || no_variant("sea-ice–ocean stress", τₛ)
        sd = NeSeaIceOceanStressDesc(grid = exchange_grid(grid; halo_ring = true), ui = devptr(uˢⁱ), vi = devptr(vˢⁱ),
                                     uo = surface_plane(τₛ.uₑ).ptr, vo = surface_plane(τₛ.vₑ).ptr,
                                     ocean_density = Float64(τₛ.ρₑ), drag_coefficient = Float64(τₛ.Cᴰ),
                                     x_momentum = devptr(fluxes.x_momentum), y_momentum = devptr(fluxes.y_momentum))
        GC.@preserve model check(ccall((entry("ne_sea_ice_ocean_stress", grid), libne[]), Cint, (Ref{NeSeaIceOceanStressDesc}, Ptr{Cvoid}), sd, stream()))
    end
    kind, αₕ, αₛ, cond, kᵢ, Tᵢ = NE_SIO_ICE_BATH, 0.0, 0.0, Int32(0), 0.0, Ptr{Cvoid}(C_NULL)
    if ff isa IceBathHeatFlux
        αₕ = Float64(ff.heat_transfer_coefficient)
    elseif ff isa ThreeEquationHeatFlux
        kind, αₕ, αₛ = NE_SIO_THREE_EQUATION, Float64(ff.heat_transfer_coefficient), Float64(ff.salt_transfer_coefficient)
        if ff.conductive_flux isa ConductiveFlux
            cond, kᵢ, Tᵢ = Int32(1), Float64(ff.conductive_flux.conductivity), devptr(ff.internal_temperature)
        elseif !isnothing(ff.conductive_flux)
            no_variant("ThreeEquationHeatFlux conductive_flux", ff.conductive_flux)
        end
    else
        no_variant("sea_ice_ocean_heat_flux", ff)
    end
    uk, u★ = friction_velocity_fields(ff.friction_velocity)
    pt = sea_ice.model.phase_transitions
    mass = sea_ice.model.mass_fluxes.thermodynamics
    Hz = halo_size(ocean.model.grid)[3]
    d = NeSeaIceOceanDesc(grid = exchange_grid(grid; halo_ring = false), nz = size(ocean.model.grid, 3), hz = Hz,
                          T = devptr(Tᵒᶜ), S = devptr(Sᵒᶜ), dz = devptr(cell_thicknesses(ocean.model.grid)), dt = Float64(sea_ice.Δt),
                          formulation = kind, friction_velocity_kind = uk, heat_transfer_coefficient = αₕ, salt_transfer_coefficient = αₛ,
                          friction_velocity = u★, has_conductive_flux = cond, conductivity = kᵢ, internal_temperature = Tᵢ,
                          latent_heat = Float64(pt.reference_latent_heat),
                          ocean = medium_properties(model.interfaces.ocean_properties; liquidus = pt.liquidus),
                          hi = slot(sea_ice.model.ice_thickness), hc = slot(sea_ice.model.ice_consolidation_thickness),
                          concentration = slot(sea_ice.model.ice_concentration), ice_salinity = slot(sea_ice.model.tracers.S),
                          ice_mass_flux = slot(mass.ice), snow_mass_flux = slot(mass.snow),
                          x_momentum_in = devptr(fluxes.x_momentum), y_momentum_in = devptr(fluxes.y_momentum),
                          frazil_heat = devptr(fluxes.frazil_heat), interface_heat = devptr(fluxes.interface_heat),
                          salt = devptr(fluxes.salt), freshwater = devptr(fluxes.freshwater),
                          interface_temperature = devptr(interface.temperature), interface_salinity = devptr(interface.salinity))
    GC.@preserve model check(ccall((entry("ne_sea_ice_ocean_fluxes", grid), libne[]), Cint, (Ref{NeSeaIceOceanDesc}, Ptr{Cvoid}), d, stream()))
    return nothing
end

# FreezingLimitedOceanTemperature, the OceanOnlyModel default (freezing_limited_ocean_temperature.jl:73-118): the clamp only
function compute_sea_ice_ocean_fluxes!(model::B200FreezingLimitedModel)
    enabled() || return invoke(compute_sea_ice_ocean_fluxes!, Tuple{NumericalEarth.SeaIces.FreezingLimitedEarthSystemModel}, model)
    ocean, sea_ice = model.ocean, model.sea_ice
    grid = ocean.model.grid
    Δt = ocean.model.clock.iteration == 0 ? Inf : Float64(ocean.Δt)         # :88
    d = NeSeaIceOceanDesc(grid = exchange_grid(grid; halo_ring = false), nz = size(grid, 3), hz = halo_size(grid)[3],
                          T = devptr(ocean.model.tracers.T), S = devptr(ocean.model.tracers.S), dz = devptr(cell_thicknesses(grid)), dt = Δt,
                          formulation = NE_SIO_FREEZE_ONLY,
                          ocean = medium_properties(model.interfaces.ocean_properties; liquidus = sea_ice.liquidus),
                          frazil_heat = devptr(sea_ice.frazil_heat))
    GC.@preserve model check(ccall((entry("ne_sea_ice_ocean_fluxes", grid), libne[]), Cint, (Ref{NeSeaIceOceanDesc}, Ptr{Cvoid}), d, stream()))
    return nothing
end

#####
##### phase 3: net fluxes (Oceans/assemble_net_ocean_fluxes.jl:30-70, SeaIces/assemble_net_sea_ice_fluxes.jl:11-40)
#####

flux_slots(f::ZeroFluxes) = (interface_heat = slot(0), salt = slot(0), freshwater = slot(0), x_momentum = slot(0), y_momentum = slot(0),
                             frazil_heat = slot(0), sensible_heat = slot(0), latent_heat = slot(0), water_vapor = slot(0))
flux_slots(f) = map(slot, NamedTuple{fieldnames(typeof(f))}(ntuple(i -> getfield(f, i), fieldcount(typeof(f)))))

const OceananigansOcean = Simulation{<:Oceananigans.Models.HydrostaticFreeSurfaceModels.HydrostaticFreeSurfaceModel}

function update_net_fluxes!(model::B200Model, ocean::OceananigansOcean)
    enabled() || return invoke(update_net_fluxes!, Tuple{Any, OceananigansOcean}, model, ocean)
    isnothing(model.interfaces.atmosphere_ocean_interface) && isnothing(model.interfaces.sea_ice_ocean_interface) && return nothing
    model.sea_ice isa FreezingLimitedOceanTemperature && isnothing(model.atmosphere) && return nothing
    grid = ocean.model.grid
    net = model.interfaces.net_fluxes.ocean
    ao = flux_slots(computed_fluxes(model.interfaces.atmosphere_ocean_interface))
    io = flux_slots(computed_fluxes(model.interfaces.sea_ice_ocean_interface))
    atmos = model.interfaces.exchanger.atmosphere
    land = model.interfaces.exchanger.land
    d = NeAssembleOceanDesc(grid = exchange_grid(grid; halo_ring = false),
                            sensible_heat = ao.sensible_heat, latent_heat = ao.latent_heat, water_vapor = ao.water_vapor,
                            x_momentum_ao = ao.x_momentum, y_momentum_ao = ao.y_momentum,
                            interface_heat = io.interface_heat, salt_io = io.salt, freshwater_io = io.freshwater,
                            x_momentum_io = io.x_momentum, y_momentum_io = io.y_momentum,
                            ocean_surface_temperature = surface_plane(ocean_surface_temperature(ocean)),
                            concentration = slot(sea_ice_concentration(model.sea_ice)),
                            rainfall = isnothing(atmos) ? slot(0) : slot(atmos.state.Jʳⁿ), snowfall = isnothing(atmos) ? slot(0) : slot(atmos.state.Jˢⁿ),
                            intercepted_snowfall = slot(intercepted_snowfall(model.sea_ice)),
                            land_freshwater = isnothing(land) ? slot(0) : slot(land.state.freshwater_flux),
                            inactive = maskptr(grid), ocean = medium_properties(model.interfaces.ocean_properties),
                            tau_x = devptr(net.u), tau_y = devptr(net.v), JT = devptr(net.T), JS = devptr(net.S), Jw = devptr(net.η),
                            JH = devptr(net.freshwater_heat_content))
    GC.@preserve model check(ccall((entry("ne_assemble_net_ocean_fluxes", grid), libne[]), Cint, (Ref{NeAssembleOceanDesc}, Ptr{Cvoid}), d, stream()))
    return nothing
end

function update_net_fluxes!(model::B200Model, sea_ice::Simulation{<:SeaIceModel})
    enabled() || return invoke(update_net_fluxes!, Tuple{Any, Simulation{<:SeaIceModel}}, model, sea_ice)
    grid = sea_ice.model.grid
    top = model.interfaces.net_fluxes.sea_ice.top
    bottom = model.interfaces.net_fluxes.sea_ice.bottom
    ai = flux_slots(computed_fluxes(model.interfaces.atmosphere_sea_ice_interface))
    io = flux_slots(computed_fluxes(model.interfaces.sea_ice_ocean_interface))
    atmos = model.interfaces.exchanger.atmosphere
    d = NeAssembleSeaIceDesc(grid = exchange_grid(grid; halo_ring = false),
                             sensible_heat = ai.sensible_heat, latent_heat = ai.latent_heat, x_momentum = ai.x_momentum, y_momentum = ai.y_momentum,
                             frazil_heat = io.frazil_heat, interface_heat = io.interface_heat,
                             snowfall = isnothing(atmos) ? slot(0) : slot(atmos.state.Jˢⁿ), concentration = slot(sea_ice_concentration(sea_ice)),
                             inactive = maskptr(grid), top_heat = devptr(top.heat), top_snowfall = devptr(top.snowfall),
                             top_u = devptr(top.u), top_v = devptr(top.v), bottom_heat = devptr(bottom.heat))
    GC.@preserve model check(ccall((entry("ne_assemble_net_sea_ice_fluxes", grid), libne[]), Cint, (Ref{NeAssembleSeaIceDesc}, Ptr{Cvoid}), d, stream()))
    return nothing
end

#####
##### phase 4: radiation (Radiations/apply_air_sea_radiative_fluxes.jl:18-60, apply_air_sea_ice_radiative_fluxes.jl:21-53)
#####

function apply_air_sea_radiative_fluxes!(model::B200Model)
    enabled() || return invoke(apply_air_sea_radiative_fluxes!, Tuple{EarthSystemModel}, model)
    (isnothing(model.ocean) || isnothing(model.interfaces.atmosphere_ocean_interface)) && return nothing
    rf = model.radiation.interface_fluxes
    (isnothing(rf) || !haskey(rf, :ocean)) && return nothing
    grid = model.interfaces.exchanger.grid
    tcr = get_radiative_forcing(model.ocean)
    two_color = tcr isa TwoColorRadiation
    d = NeApplyRadiationDesc(grid = exchange_grid(grid; halo_ring = false), radiation = surface_radiation(model, :ocean),
                             concentration = slot(sea_ice_concentration(model.sea_ice)),
                             surface_temperature = devptr(model.interfaces.atmosphere_ocean_interface.temperature),
                             medium = medium_properties(model.interfaces.ocean_properties), inactive = maskptr(grid),
                             over_sea_ice = Int32(0), two_color = Int32(two_color), heat_flux = devptr(model.interfaces.net_fluxes.ocean.T),
                             two_color_surface_flux = two_color ? devptr(tcr.surface_flux) : Ptr{Cvoid}(C_NULL),
                             upwelling_longwave = devptr(rf.ocean.upwelling_longwave), downwelling_longwave = devptr(rf.ocean.downwelling_longwave),
                             downwelling_shortwave = devptr(rf.ocean.downwelling_shortwave))
    GC.@preserve model check(ccall((entry("ne_apply_radiative_fluxes", grid), libne[]), Cint, (Ref{NeApplyRadiationDesc}, Ptr{Cvoid}), d, stream()))
    return nothing
end

function apply_air_sea_ice_radiative_fluxes!(model::B200Model)
    enabled() || return invoke(apply_air_sea_ice_radiative_fluxes!, Tuple{EarthSystemModel}, model)
    sea_ice = model.sea_ice
    sea_ice isa Simulation{<:SeaIceModel} || return nothing
    rf = model.radiation.interface_fluxes
    (isnothing(rf) || !haskey(rf, :sea_ice)) && return nothing
    grid = sea_ice.model.grid
    d = NeApplyRadiationDesc(grid = exchange_grid(grid; halo_ring = false), radiation = surface_radiation(model, :sea_ice),
                             concentration = slot(sea_ice_concentration(sea_ice)),
                             surface_temperature = devptr(model.interfaces.atmosphere_sea_ice_interface.temperature),
                             medium = medium_properties(model.interfaces.sea_ice_properties), inactive = maskptr(grid),
                             over_sea_ice = Int32(1), heat_flux = devptr(model.interfaces.net_fluxes.sea_ice.top.heat),
                             upwelling_longwave = devptr(rf.sea_ice.upwelling_longwave), downwelling_longwave = devptr(rf.sea_ice.downwelling_longwave),
                             downwelling_shortwave = devptr(rf.sea_ice.downwelling_shortwave))
    GC.@preserve model check(ccall((entry("ne_apply_radiative_fluxes", grid), libne[]), Cint, (Ref{NeApplyRadiationDesc}, Ptr{Cvoid}), d, stream()))
    return nothing
end

#####
##### diagnostics: area-weighted global sums of flux fields (src/Diagnostics/interface_fluxes.jl:90-195) for the
##### conservation checks of a run sharded in latitude bands — the ONE place this path has a collective.
#####
# `ne_diag_reduce_*` leaves n_fields doubles in `result` (device) in a fixed summation order; a run on N GPUs (one rank per
# GPU: MPI.jl / Oceananigans' Distributed) sums the N result vectors with `ne_diag_allreduce_f64(comm, result, n, stream)`:
# NCCL bound at run time by the library, the communicator is the host's own (NCCL.jl: `comm.handle`, created once from the
# MPI communicator of the run); 56 bytes, latency-bound, on the same stream behind the reduction kernels.

struct FluxDiagnostics{R, P, A}
    result :: R      # CuVector{Float64}(n_fields)
    partial :: P     # CuVector{Float64}(n_blocks * n_fields) scratch
    area :: A        # exchange-layout cell areas (Field) or nothing
    n_blocks :: Int
end

FluxDiagnostics(n_fields::Int; area = nothing, n_blocks = 1184) =
    FluxDiagnostics(CUDA.zeros(Float64, n_fields), CUDA.zeros(Float64, n_blocks * n_fields), area, n_blocks)

"sum_i area_i field_i over active cells for every field of `fields` (Fields on the exchange grid); returns diag.result (device)"
"in-place sum of a device vector of Float64 over the ranks of an NCCL communicator (`comm.handle` of NCCL.jl)"
function allreduce_fluxes!(result, nccl_comm::Ptr{Cvoid})
    check(ccall((:ne_diag_allreduce_f64, libne[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
                nccl_comm, devptr(result), Int32(length(result)), stream()))
    return result
end

function reduce_fluxes!(diag::FluxDiagnostics, grid, fields; allreduce! = identity)
    enabled() || error("reduce_fluxes! needs libne_b200 (NE_B200_LIB)")
    length(fields) <= NE_DIAG_MAX_FIELDS || error("at most $NE_DIAG_MAX_FIELDS fields per reduction")
    ptrs = ntuple(k -> k <= length(fields) ? devptr(fields[k]) : Ptr{Cvoid}(C_NULL), NE_DIAG_MAX_FIELDS)
    d = NeDiagDesc(grid = exchange_grid(grid; halo_ring = false), n_fields = Int32(length(fields)), fields = ptrs,
                   area = devptr(diag.area), inactive = maskptr(grid), partial = devptr(diag.partial), n_blocks = diag.n_blocks,
                   result = devptr(diag.result))
    GC.@preserve diag fields check(ccall((entry("ne_diag_reduce", grid), libne[]), Cint, (Ref{NeDiagDesc}, Ptr{Cvoid}), d, stream()))
    allreduce!(diag.result)        # e.g. r -> allreduce_fluxes!(r, comm.handle)
    return diag.result
end

end # module NumericalEarthB200Ext
