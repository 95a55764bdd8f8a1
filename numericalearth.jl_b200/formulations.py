"""Host-side mirror of the reference's flux-formulation plugin types and their translation to the
POD kernel variants of include/ne_b200.h.

Names, keyword arguments and defaults follow the reference (all citations relative to
/root/reference/src/EarthSystemModels/InterfaceComputations/ unless noted):

    SimilarityTheoryFluxes            similarity_theory_turbulent_fluxes.jl:174-214
    ConvectiveGustiness, SubgridVelocityCorrection, mahrt_sun_subgrid_velocity   :45-117
    Edson/Sheba/Paulson/LinearStable/Split stability functions                   :487-789
    MomentumRoughnessLength, ScalarRoughnessLength, ReynoldsScalingFunction,
    WindDependentWaveFormulation, TemperatureDependentAirViscosity              roughness_lengths.jl
    ConvergenceStopCriteria, FixedIterations                                     compute_interface_state.jl:5-26
    CoefficientBasedFluxes, PolynomialNeutralDragCoefficient,
    LargeYeagerTransferCoefficients                                              coefficient_based_turbulent_fluxes.jl
    InterfaceProperties pieces                                                   interface_states.jl
    AtmosphereThermodynamicsParameters                                           ../../Atmospheres/thermodynamic_parameters.jl
    IceBathHeatFlux, ThreeEquationHeatFlux, MomentumBasedFrictionVelocity        sea_ice_ocean_heat_flux_formulations.jl, friction_velocity.jl

Anything that is not one of these types (a Python callable standing in for a Julia closure, an
unknown object) raises NoKernelVariantError — there is no CPU fallback (BASELINE.json north_star).
Values are stored at Float64 and narrowed to the model's float type inside the kernels exactly as
`convert(FT, x)` does in the reference constructors.
"""
import math
from dataclasses import dataclass, field
from typing import Any, Optional

from . import abi as A


class NoKernelVariantError(ValueError):
    """Raised when a plugin type has no sm_100a kernel variant (the reference's ArgumentError)."""


def _f32(x):
    import struct
    return struct.unpack("f", struct.pack("f", float(x)))[0]


def _conv(FT, x):
    """convert(FT, x) for FT in {'f32','f64'} keeping a Python float holding the rounded value."""
    return _f32(x) if FT == "f32" else float(x)


# ---------------------------------------------------------------------------------------------
# stability functions
# ---------------------------------------------------------------------------------------------
@dataclass
class EdsonMomentumStabilityFunction:  # :487-499
    ζmax: float = 50.0
    Ap: float = 0.35
    Bp: float = 0.7
    Cp: float = 0.75
    Dp: float = 5 / 0.35
    Am: float = 15.0
    Bm: float = 2.0
    Cm: float = math.pi / 2
    Dm: float = 10.15
    Em: float = 3.0
    Fm: float = math.pi / math.sqrt(3)

    def pod(self):
        f = A.NeStabilityFn(kind=A.NE_PSI_EDSON_MOMENTUM)
        for k, v in enumerate([self.ζmax, self.Ap, self.Bp, self.Cp, self.Dp, self.Am, self.Bm, self.Cm, self.Dm,
                               self.Em, self.Fm]):
            f.p[k] = v
        return f


@dataclass
class EdsonScalarStabilityFunction:  # :571-584
    ζmax: float = 50.0
    Ap: float = 0.35
    Bp: float = 2 / 3
    Cp: float = 3 / 2
    Dp: float = 14.28
    Ep: float = 8.525
    Am: float = 15.0
    Bm: float = 2.0
    Cm: float = 0.0
    Dm: float = 34.15
    Em: float = 3.0
    Fm: float = math.pi / math.sqrt(3)

    def pod(self):
        f = A.NeStabilityFn(kind=A.NE_PSI_EDSON_SCALAR)
        for k, v in enumerate([self.ζmax, self.Ap, self.Bp, self.Cp, self.Dp, self.Ep, self.Am, self.Bm, self.Cm,
                               self.Dm, self.Em, self.Fm]):
            f.p[k] = v
        return f


@dataclass
class ShebaMomentumStabilityFunction:  # :637-640
    a: float = 6.5
    b: float = 1.3

    def pod(self):
        f = A.NeStabilityFn(kind=A.NE_PSI_SHEBA_MOMENTUM)
        f.p[0], f.p[1] = self.a, self.b
        return f


@dataclass
class ShebaScalarStabilityFunction:  # :659-663
    a: float = 5.0
    b: float = 5.0
    c: float = 3.0

    def pod(self):
        f = A.NeStabilityFn(kind=A.NE_PSI_SHEBA_SCALAR)
        f.p[0], f.p[1], f.p[2] = self.a, self.b, self.c
        return f


@dataclass
class PaulsonMomentumStabilityFunction:  # :683-686
    a: float = 16.0
    b: float = math.pi / 2

    def pod(self):
        f = A.NeStabilityFn(kind=A.NE_PSI_PAULSON_MOMENTUM)
        f.p[0], f.p[1] = self.a, self.b
        return f


@dataclass
class PaulsonScalarStabilityFunction:  # :701-703
    a: float = 16.0

    def pod(self):
        f = A.NeStabilityFn(kind=A.NE_PSI_PAULSON_SCALAR)
        f.p[0] = self.a
        return f


@dataclass
class LinearStableStabilityFunction:  # :742-745
    coefficient: float = 5.0
    maximum_stability_parameter: float = 10.0

    def pod(self):
        f = A.NeStabilityFn(kind=A.NE_PSI_LINEAR_STABLE)
        f.p[0], f.p[1] = self.coefficient, self.maximum_stability_parameter
        return f


@dataclass
class ZeroStabilityFunction:
    """Returns(zero(FT)) — `stability_functions = nothing` (:201-204)."""

    def pod(self):
        return A.NeStabilityFn(kind=A.NE_PSI_ZERO)


@dataclass
class SplitStabilityFunction:  # :712-725
    stable: Any
    unstable: Any


_SIMPLE_PSI = (EdsonMomentumStabilityFunction, EdsonScalarStabilityFunction, ShebaMomentumStabilityFunction,
               ShebaScalarStabilityFunction, PaulsonMomentumStabilityFunction, PaulsonScalarStabilityFunction,
               LinearStableStabilityFunction, ZeroStabilityFunction)


def stability_profile_pod(psi) -> A.NeStabilityProfile:
    p = A.NeStabilityProfile()
    if psi is None:
        psi = ZeroStabilityFunction()
    if isinstance(psi, SplitStabilityFunction):
        for side in (psi.stable, psi.unstable):
            if not isinstance(side, _SIMPLE_PSI):
                raise NoKernelVariantError(f"stability function {side!r} has no kernel variant")
        p.split = 1
        p.a = psi.stable.pod()
        p.b = psi.unstable.pod()
    elif isinstance(psi, _SIMPLE_PSI):
        p.split = 0
        p.a = psi.pod()
    else:
        raise NoKernelVariantError(
            f"stability function {psi!r} has no kernel variant (user-defined callables are not supported; "
            "there is no CPU fallback)")
    return p


@dataclass
class SimilarityScales:  # :432-436
    momentum: Any
    temperature: Any
    water_vapor: Any


def atmosphere_ocean_stability_functions():  # :621-625
    c = EdsonScalarStabilityFunction()
    return SimilarityScales(EdsonMomentumStabilityFunction(), c, c)


def large_yeager_stability_functions():  # :766-771
    stable = LinearStableStabilityFunction()
    return SimilarityScales(SplitStabilityFunction(stable, PaulsonMomentumStabilityFunction()),
                            SplitStabilityFunction(stable, PaulsonScalarStabilityFunction()),
                            SplitStabilityFunction(stable, PaulsonScalarStabilityFunction()))


def atmosphere_sea_ice_stability_functions():  # :779-789
    m = SplitStabilityFunction(ShebaMomentumStabilityFunction(), PaulsonMomentumStabilityFunction())
    s = SplitStabilityFunction(ShebaScalarStabilityFunction(), PaulsonScalarStabilityFunction())
    return SimilarityScales(m, s, s)


# ---------------------------------------------------------------------------------------------
# roughness lengths (roughness_lengths.jl)
# ---------------------------------------------------------------------------------------------
@dataclass
class WindDependentWaveFormulation:  # :56-72
    Umax: float = 19
    C1: float = 0.0017
    C2: float = -0.005


@dataclass
class TemperatureDependentAirViscosity:  # :149-180
    C0: float = 1.326e-5
    C1: float = 1.326e-5 * 6.542e-3
    C2: float = 1.326e-5 * 8.301e-6
    C3: float = -1.326e-5 * 4.84e-9


@dataclass
class ReynoldsScalingFunction:  # :212-229
    A: float = 5.85e-5
    b: float = 0.72


def _viscosity_into(r, nu):
    if isinstance(nu, TemperatureDependentAirViscosity):
        r.visc_kind = A.NE_VISC_TEMPERATURE_DEPENDENT
        r.nu_C[0], r.nu_C[1], r.nu_C[2], r.nu_C[3] = nu.C0, nu.C1, nu.C2, nu.C3
    elif isinstance(nu, (int, float)):
        r.visc_kind = A.NE_VISC_CONSTANT
        r.visc_dtype = A.NE_F64  # the Float64 literal is never converted to FT (:94, 126)
        r.nu = float(nu)
    else:
        raise NoKernelVariantError(f"air_kinematic_viscosity {nu!r} has no kernel variant")


@dataclass
class MomentumRoughnessLength:  # :123-139
    gravitational_acceleration: float = 9.80665
    maximum_roughness_length: float = 1.0
    air_kinematic_viscosity: Any = 1.5e-5
    wave_formulation: Any = 0.02
    smooth_wall_parameter: float = 0.11

    def pod(self):
        r = A.NeRoughnessLength(kind=A.NE_ROUGH_MOMENTUM)
        r.gravitational_acceleration = self.gravitational_acceleration
        r.maximum_roughness_length = self.maximum_roughness_length
        r.smooth_wall_parameter = self.smooth_wall_parameter
        if isinstance(self.wave_formulation, WindDependentWaveFormulation):
            r.wave_kind = A.NE_WAVE_WIND_DEPENDENT
            w = self.wave_formulation
            r.wave_Umax, r.wave_C1, r.wave_C2 = w.Umax, w.C1, w.C2
        elif isinstance(self.wave_formulation, (int, float)):
            r.wave_kind = A.NE_WAVE_CONSTANT
            r.wave_constant = float(self.wave_formulation)
        else:
            raise NoKernelVariantError(f"wave_formulation {self.wave_formulation!r} has no kernel variant")
        _viscosity_into(r, self.air_kinematic_viscosity)
        return r


@dataclass
class ScalarRoughnessLength:  # :93-101
    air_kinematic_viscosity: Any = 1.5e-5
    reynolds_number_scaling_function: Any = field(default_factory=ReynoldsScalingFunction)
    maximum_roughness_length: float = 1.6e-4

    def pod(self):
        r = A.NeRoughnessLength(kind=A.NE_ROUGH_SCALAR)
        r.maximum_roughness_length = self.maximum_roughness_length
        s = self.reynolds_number_scaling_function
        if not isinstance(s, ReynoldsScalingFunction):
            raise NoKernelVariantError(f"reynolds_number_scaling_function {s!r} has no kernel variant")
        r.reynolds_A, r.reynolds_b = s.A, s.b
        _viscosity_into(r, self.air_kinematic_viscosity)
        return r


@dataclass
class LandRoughnessLength:  # roughness_lengths.jl:21-39
    """The land model's per-cell aerodynamic roughness field as a MOST roughness length (`multiplier` scales it: 0.1 for scalar
    lengths taken as a tenth of the momentum one).  `minimum_roughness_length = None` is the reference's eps(FT)."""
    multiplier: float = 1
    minimum_roughness_length: Any = None
    FT: str = "f64"

    def pod(self):
        import numpy as np
        npf = np.float64 if self.FT == "f64" else np.float32
        r = A.NeRoughnessLength(kind=A.NE_ROUGH_LAND)
        lmin = np.finfo(npf).eps if self.minimum_roughness_length is None else self.minimum_roughness_length
        r.land_multiplier = float(npf(self.multiplier))                 # convert(FT, …) in the constructor (:34-39)
        r.land_minimum_roughness_length = float(npf(lmin))
        return r


class LandZeroPlaneDisplacement:  # roughness_lengths.jl:44-51
    """Marker: the zero-plane displacement is the land model's per-cell field (0 where the land model provides none)."""


def roughness_pod(ell) -> A.NeRoughnessLength:
    if isinstance(ell, (MomentumRoughnessLength, ScalarRoughnessLength, LandRoughnessLength)):
        return ell.pod()
    if isinstance(ell, (int, float)):  # roughness_length(ℓ::Number, args...) = ℓ (:193)
        r = A.NeRoughnessLength(kind=A.NE_ROUGH_CONSTANT)
        r.constant = float(ell)
        return r
    raise NoKernelVariantError(f"roughness length {ell!r} has no kernel variant (callables are not supported)")


# ---------------------------------------------------------------------------------------------
# subgrid velocities
# ---------------------------------------------------------------------------------------------
@dataclass
class ConvectiveGustiness:  # :45-48
    gustiness_parameter: float = 1.2
    minimum_gustiness: float = 0.01


@dataclass
class SubgridVelocityCorrection:  # :69-80
    convective: Any = field(default_factory=ConvectiveGustiness)
    mesoscale: Any = None


def mahrt_sun_subgrid_velocity(dx, threshold=5e3):  # :114-117
    delta = max(dx / threshold - 1, 0)
    return 0.32 * delta ** 0.33


def _sgs_slot(x):
    if x is None:
        return A.NE_SGS_NONE, 0.0, None
    if isinstance(x, ConvectiveGustiness):
        return A.NE_SGS_CONVECTIVE, 0.0, x
    if isinstance(x, (int, float)):
        return A.NE_SGS_CONSTANT, float(x), None
    raise NoKernelVariantError(f"subgrid velocity formulation {x!r} has no kernel variant")


def subgrid_pod(sv) -> A.NeSubgridVelocity:
    s = A.NeSubgridVelocity()
    if isinstance(sv, SubgridVelocityCorrection):
        s.composite = 1
        ck, cc, cg = _sgs_slot(sv.convective)
        mk, mc, mg = _sgs_slot(sv.mesoscale)
        if mk == A.NE_SGS_CONVECTIVE and ck == A.NE_SGS_CONVECTIVE:
            raise NoKernelVariantError("two ConvectiveGustiness slots have no kernel variant")
        s.convective_kind, s.convective_constant = ck, cc
        s.mesoscale_kind, s.mesoscale_constant = mk, mc
        g = cg or mg
    else:
        ck, cc, g = _sgs_slot(sv)
        s.convective_kind, s.convective_constant = ck, cc
    if g is not None:
        s.gustiness_parameter, s.minimum_gustiness = g.gustiness_parameter, g.minimum_gustiness
    return s


# ---------------------------------------------------------------------------------------------
# stop criteria, similarity forms
# ---------------------------------------------------------------------------------------------
@dataclass
class ConvergenceStopCriteria:  # compute_interface_state.jl:5-8
    tolerance: float = 1e-8
    maxiter: int = 100


@dataclass
class FixedIterations:  # compute_interface_state.jl:21-25
    iterations: int = 5


def stop_pod(sc) -> A.NeStopCriteria:
    if isinstance(sc, ConvergenceStopCriteria):
        return A.NeStopCriteria(kind=A.NE_STOP_CONVERGENCE, maxiter=int(sc.maxiter), tolerance=float(sc.tolerance))
    if isinstance(sc, FixedIterations):
        return A.NeStopCriteria(kind=A.NE_STOP_FIXED_ITERATIONS, maxiter=int(sc.iterations), tolerance=0.0)
    raise NoKernelVariantError(f"solver_stop_criteria {sc!r} has no kernel variant")


class LogarithmicSimilarityProfile:  # :239
    pass


class COARELogarithmicSimilarityProfile:  # :240
    pass


# ---------------------------------------------------------------------------------------------
# flux formulations
# ---------------------------------------------------------------------------------------------
@dataclass
class SimilarityTheoryFluxes:  # :174-214
    von_karman_constant: float = 0.4
    turbulent_prandtl_number: float = 1
    subgrid_velocities: Any = field(default_factory=ConvectiveGustiness)
    stability_functions: Any = field(default_factory=atmosphere_ocean_stability_functions)
    momentum_roughness_length: Any = field(default_factory=MomentumRoughnessLength)
    temperature_roughness_length: Any = field(default_factory=ScalarRoughnessLength)
    water_vapor_roughness_length: Any = field(default_factory=ScalarRoughnessLength)
    zero_plane_displacement: Any = 0
    similarity_form: Any = field(default_factory=LogarithmicSimilarityProfile)
    solver_stop_criteria: Any = None
    solver_tolerance: float = 1e-8
    solver_maxiter: int = 100

    def __post_init__(self):
        if self.solver_stop_criteria is None:
            self.solver_stop_criteria = ConvergenceStopCriteria(self.solver_tolerance, self.solver_maxiter)
        if self.stability_functions is None:
            z = ZeroStabilityFunction()
            self.stability_functions = SimilarityScales(z, z, z)

    def pod(self) -> A.NeFluxFormulation:
        f = A.NeFluxFormulation(kind=A.NE_FLUX_SIMILARITY_THEORY)
        f.von_karman_constant = self.von_karman_constant
        f.subgrid_velocities = subgrid_pod(self.subgrid_velocities)
        sf = self.stability_functions
        if not isinstance(sf, SimilarityScales):
            raise NoKernelVariantError(f"stability_functions {sf!r} has no kernel variant")
        f.psi_momentum = stability_profile_pod(sf.momentum)
        f.psi_temperature = stability_profile_pod(sf.temperature)
        f.psi_water_vapor = stability_profile_pod(sf.water_vapor)
        f.ell_momentum = roughness_pod(self.momentum_roughness_length)
        f.ell_temperature = roughness_pod(self.temperature_roughness_length)
        f.ell_water_vapor = roughness_pod(self.water_vapor_roughness_length)
        if f.ell_momentum.kind == A.NE_ROUGH_SCALAR or f.ell_temperature.kind == A.NE_ROUGH_MOMENTUM \
                or f.ell_water_vapor.kind == A.NE_ROUGH_MOMENTUM:
            raise NoKernelVariantError("roughness length type not valid in this slot")
        if isinstance(self.zero_plane_displacement, LandZeroPlaneDisplacement):
            f.zero_plane_displacement_kind = A.NE_DISPLACEMENT_LAND
        elif isinstance(self.zero_plane_displacement, (int, float)):
            f.zero_plane_displacement = float(self.zero_plane_displacement)
        else:
            raise NoKernelVariantError(f"zero_plane_displacement {self.zero_plane_displacement!r} has no kernel variant")
        if isinstance(self.similarity_form, COARELogarithmicSimilarityProfile):
            f.similarity_form = A.NE_PROFILE_COARE
        elif isinstance(self.similarity_form, LogarithmicSimilarityProfile):
            f.similarity_form = A.NE_PROFILE_LOGARITHMIC
        else:
            raise NoKernelVariantError(f"similarity_form {self.similarity_form!r} has no kernel variant")
        f.stop = stop_pod(self.solver_stop_criteria)
        return f


def atmosphere_land_stability_functions():  # :776-777 (currently the Large-Yeager set)
    return large_yeager_stability_functions()


def default_atmosphere_land_fluxes(solver_stop_criteria=None):  # component_interfaces.jl:514-521
    kw = {} if solver_stop_criteria is None else {"solver_stop_criteria": solver_stop_criteria}
    return SimilarityTheoryFluxes(stability_functions=atmosphere_land_stability_functions(), momentum_roughness_length=0.1,
                                  temperature_roughness_length=0.01, water_vapor_roughness_length=0.01, **kw)


def atmosphere_sea_ice_similarity_theory():  # :791-794
    return SimilarityTheoryFluxes(stability_functions=atmosphere_sea_ice_stability_functions())


@dataclass
class PolynomialNeutralDragCoefficient:  # coefficient_based_turbulent_fluxes.jl:20-44
    a: float = 2.7
    b: float = 0.142
    c: float = 1 / 13.09
    d: float = 3.14807e-10
    high_wind_speed_threshold: float = 33
    high_wind_drag_coefficient: float = 2.34e-3
    minimum_wind_speed: float = 0.5

    def pod(self):
        return A.NePolynomialDrag(self.a, self.b, self.c, self.d, self.high_wind_speed_threshold,
                                  self.high_wind_drag_coefficient, self.minimum_wind_speed)


@dataclass
class LargeYeagerTransferCoefficients:  # :82-108
    von_karman_constant: float = 0.4
    neutral_drag_coefficient: Any = field(default_factory=PolynomialNeutralDragCoefficient)
    stability_functions: Any = field(default_factory=large_yeager_stability_functions)
    reference_height: float = 10
    stable_heat_transfer_coefficient: float = 18
    unstable_heat_transfer_coefficient: float = 32.7
    moisture_transfer_coefficient: float = 34.6


@dataclass
class CoefficientBasedFluxes:  # :218-232
    transfer_coefficients: Any = (1e-3, 1e-3, 1e-3)
    solver_stop_criteria: Any = None
    solver_tolerance: float = 1e-8
    solver_maxiter: int = 20

    def __post_init__(self):
        tc = self.transfer_coefficients
        if isinstance(tc, dict):  # NamedTuple: validate_coefficients :237-249
            required = ("momentum", "temperature", "water_vapor")
            missing = [k for k in required if k not in tc]
            if missing:
                raise ValueError(f"Transfer coefficients NamedTuple must contain keys {required}. Missing keys: {missing}.")
            self.transfer_coefficients = SimilarityScales(tc["momentum"], tc["temperature"], tc["water_vapor"])
        elif isinstance(tc, (tuple, list)):  # :251-258
            if len(tc) != 3:
                raise ValueError("Transfer coefficients must be a tuple of length 3: (momentum, temperature, "
                                 f"water_vapor). Got length {len(tc)} with value {tc}.")
            self.transfer_coefficients = SimilarityScales(*tc)
        if self.solver_stop_criteria is None:
            self.solver_stop_criteria = ConvergenceStopCriteria(self.solver_tolerance, self.solver_maxiter)

    def pod(self) -> A.NeFluxFormulation:
        f = A.NeFluxFormulation()
        tc = self.transfer_coefficients
        if isinstance(tc, LargeYeagerTransferCoefficients):
            f.kind = A.NE_FLUX_LARGE_YEAGER
            ly = f.large_yeager
            ly.von_karman_constant = tc.von_karman_constant
            if not isinstance(tc.neutral_drag_coefficient, PolynomialNeutralDragCoefficient):
                raise NoKernelVariantError("neutral_drag_coefficient has no kernel variant")
            ly.neutral_drag = tc.neutral_drag_coefficient.pod()
            ly.psi_momentum = stability_profile_pod(tc.stability_functions.momentum)
            ly.psi_temperature = stability_profile_pod(tc.stability_functions.temperature)
            ly.reference_height = tc.reference_height
            ly.stable_heat = tc.stable_heat_transfer_coefficient
            ly.unstable_heat = tc.unstable_heat_transfer_coefficient
            ly.moisture = tc.moisture_transfer_coefficient
        elif isinstance(tc, SimilarityScales):
            f.kind = A.NE_FLUX_COEFFICIENT_BASED
            for k, c in enumerate((tc.momentum, tc.temperature, tc.water_vapor)):
                if isinstance(c, PolynomialNeutralDragCoefficient):
                    f.coefficients[k].kind = A.NE_COEFF_POLYNOMIAL_DRAG
                    f.coefficients[k].polynomial = c.pod()
                elif isinstance(c, (int, float)):
                    f.coefficients[k].kind = A.NE_COEFF_CONSTANT
                    f.coefficients[k].constant = float(c)
                else:  # evaluate_coefficient(C::Function, ...) :270
                    raise NoKernelVariantError(
                        f"transfer coefficient {c!r} has no kernel variant (Function-valued coefficients are not supported)")
        else:
            raise NoKernelVariantError(f"transfer_coefficients {tc!r} has no kernel variant")
        f.stop = stop_pod(self.solver_stop_criteria)
        return f


def flux_formulation_pod(ff, land=False, FT="f64") -> A.NeFluxFormulation:
    """`land = False` (ocean / sea-ice interfaces): the land markers collapse to what local_roughness_length /
    local_zero_plane_displacement return for interior properties without land fields
    (similarity_theory_turbulent_fluxes.jl:265-303): max(multiplier * minimum, minimum) in FT, and 0."""
    if not isinstance(ff, (SimilarityTheoryFluxes, CoefficientBasedFluxes)):
        raise NoKernelVariantError(f"flux formulation {ff!r} has no kernel variant")
    f = ff.pod()
    if not land and f.kind == A.NE_FLUX_SIMILARITY_THEORY:
        import numpy as np
        npf = np.float64 if FT == "f64" else np.float32
        for name in ("ell_momentum", "ell_temperature", "ell_water_vapor"):
            r = getattr(f, name)
            if r.kind == A.NE_ROUGH_LAND:
                lmin = npf(r.land_minimum_roughness_length)
                c = A.NeRoughnessLength(kind=A.NE_ROUGH_CONSTANT)
                c.constant = float(max(npf(r.land_multiplier) * lmin, lmin))
                setattr(f, name, c)
        if f.zero_plane_displacement_kind == A.NE_DISPLACEMENT_LAND:
            f.zero_plane_displacement_kind, f.zero_plane_displacement = A.NE_DISPLACEMENT_CONSTANT, 0.0
    return f


# ---------------------------------------------------------------------------------------------
# interface properties (interface_states.jl)
# ---------------------------------------------------------------------------------------------
class Liquid:
    pass


class Ice:
    pass


@dataclass
class WaterMoleFraction:  # :236-253
    water_molar_mass: float = 18.02
    molar_masses: tuple = (35.45, 22.99, 96.06, 24.31)      # chloride, sodium, sulfate, magnesium
    mass_fractions: tuple = (0.56, 0.31, 0.08, 0.05)


@dataclass
class ImpureSaturationSpecificHumidity:  # :20-44
    phase: Any = field(default_factory=Liquid)
    water_mole_fraction: Any = None


# ---- land surface humidity closures (interface_states.jl:92-229) --------------------------------------------
@dataclass
class BulkHumidity:  # :107-126
    phase: Any = field(default_factory=Liquid)


@dataclass
class CriticalSaturation:  # :145-157 (Manabe 1969)
    critical_saturation: float = 0.75


@dataclass
class FractionalHumidity:  # :174-181
    efficiency: Any = None      # CriticalSaturation or a constant number
    phase: Any = field(default_factory=Liquid)


@dataclass
class SkinHumidity:  # :208-219, solved inside the iteration :625-651
    surface_thickness: float = 0.1
    vapor_diffusivity: float = 2e-2
    phase: Any = field(default_factory=Liquid)


@dataclass
class StorageBasedDryLayerDepth:  # dry_layer_humidity.jl: δᵛ(𝒮) = δᵛmax [1 − min(𝒮/𝒮ᶜ, 1)]^η
    maximum_dry_layer_depth: float = 0.05
    dry_layer_onset_saturation: float = 0.5
    dry_layer_exponent: float = 2.0


class ConstantTortuosity:
    pass


class PowerLawTortuosity:  # Millington–Quirk
    pass


@dataclass
class DryLayerVaporPistonVelocity:
    minimum_dry_layer_depth: float = 1e-4
    molecular_diffusivity: float = 2.5e-5
    wet_transition_width: Any = None            # default 5 δᵛmin
    tortuosity: Any = field(default_factory=ConstantTortuosity)

    def __post_init__(self):
        if self.wet_transition_width is None:
            self.wet_transition_width = 5 * self.minimum_dry_layer_depth


@dataclass
class DryLayerHumidity:  # dry_layer_humidity.jl
    dry_layer_depth: Any = field(default_factory=StorageBasedDryLayerDepth)
    vapor_exchange: Any = field(default_factory=DryLayerVaporPistonVelocity)
    thermal_exchange_depth: float = 0.10
    porosity: float = 0.4
    phase: Any = field(default_factory=Liquid)


def land_humidity_pod(q) -> A.NeLandHumidity:
    h = A.NeLandHumidity()
    if isinstance(q, BulkHumidity):
        h.kind = A.NE_LANDQ_BULK
    elif isinstance(q, FractionalHumidity):
        if isinstance(q.efficiency, CriticalSaturation):
            h.kind, h.critical_saturation = A.NE_LANDQ_FRACTIONAL_CRITICAL, float(q.efficiency.critical_saturation)
        elif isinstance(q.efficiency, (int, float)):
            h.kind, h.efficiency = A.NE_LANDQ_FRACTIONAL_CONSTANT, float(q.efficiency)
        else:
            raise NoKernelVariantError(f"evaporation efficiency {q.efficiency!r} has no kernel variant")
    elif isinstance(q, SkinHumidity):
        if not isinstance(q.surface_thickness, (int, float)):
            raise NoKernelVariantError("SkinHumidity with a wetness-dependent surface thickness has no kernel variant")
        h.kind = A.NE_LANDQ_SKIN
        h.surface_thickness, h.vapor_diffusivity = float(q.surface_thickness), float(q.vapor_diffusivity)
    elif isinstance(q, DryLayerHumidity):
        dd, vx = q.dry_layer_depth, q.vapor_exchange
        if not isinstance(dd, StorageBasedDryLayerDepth) or not isinstance(vx, DryLayerVaporPistonVelocity):
            raise NoKernelVariantError("DryLayerHumidity: only StorageBasedDryLayerDepth + DryLayerVaporPistonVelocity have a kernel variant")
        h.kind = A.NE_LANDQ_DRY_LAYER
        h.maximum_dry_layer_depth, h.dry_layer_onset_saturation = float(dd.maximum_dry_layer_depth), float(dd.dry_layer_onset_saturation)
        h.dry_layer_exponent = float(dd.dry_layer_exponent)
        h.minimum_dry_layer_depth, h.molecular_diffusivity = float(vx.minimum_dry_layer_depth), float(vx.molecular_diffusivity)
        h.wet_transition_width = float(vx.wet_transition_width)
        if isinstance(vx.tortuosity, ConstantTortuosity):
            h.tortuosity = A.NE_TORTUOSITY_CONSTANT
        elif isinstance(vx.tortuosity, PowerLawTortuosity):
            h.tortuosity = A.NE_TORTUOSITY_POWER_LAW
        else:
            raise NoKernelVariantError(f"tortuosity {vx.tortuosity!r} has no kernel variant")
        h.thermal_exchange_depth, h.porosity = float(q.thermal_exchange_depth), float(q.porosity)
    else:
        raise NoKernelVariantError(f"land humidity formulation {q!r} has no kernel variant")
    if isinstance(q.phase, Liquid):
        h.phase = A.NE_PHASE_LIQUID
    elif isinstance(q.phase, Ice):
        h.phase = A.NE_PHASE_ICE
    else:
        raise NoKernelVariantError(f"phase {q.phase!r} has no kernel variant")
    return h


class BulkTemperature:  # :330
    pass


@dataclass
class InteriorDiffusivity:  # :384-388
    minimum_diffusivity: float = 1.4e-7


@dataclass
class DiffusiveFlux:  # :371-374
    κ: Any = 1e-2
    δ: float = 1.0


@dataclass
class ConductiveFlux:  # ClimaSeaIce.SeaIceThermodynamics.ConductiveFlux
    conductivity: float = 2.0


@dataclass
class IceSnowConductiveFlux:  # ClimaSeaIce.SeaIceThermodynamics.IceSnowConductiveFlux
    snow_conductivity: float = 0.31
    ice_conductivity: float = 2.0


@dataclass
class SkinTemperature:  # :355-360
    internal_flux: Any = None
    max_ΔT: float = 5


class RelativeVelocity:  # :287
    pass


class WindVelocity:  # :284
    pass


@dataclass
class InterfaceProperties:  # :8-12
    specific_humidity_formulation: Any = field(default_factory=lambda: ImpureSaturationSpecificHumidity(Liquid(), 0.98))
    temperature_formulation: Any = field(default_factory=BulkTemperature)
    velocity_formulation: Any = field(default_factory=RelativeVelocity)

    def pod(self) -> A.NeInterfaceProperties:
        p = A.NeInterfaceProperties()
        q = self.specific_humidity_formulation
        if not isinstance(q, ImpureSaturationSpecificHumidity):
            raise NoKernelVariantError(f"specific humidity formulation {q!r} has no kernel variant "
                                       "(land humidity closures are a 'next' row)")
        if isinstance(q.phase, Liquid):
            p.phase = A.NE_PHASE_LIQUID
        elif isinstance(q.phase, Ice):
            p.phase = A.NE_PHASE_ICE
        else:
            raise NoKernelVariantError(f"phase {q.phase!r} has no kernel variant")
        x = q.water_mole_fraction
        if x is None:
            p.x_h2o_kind = A.NE_XH2O_ONE
        elif isinstance(x, (int, float)):
            p.x_h2o_kind, p.x_h2o = A.NE_XH2O_CONSTANT, float(x)
        elif isinstance(x, WaterMoleFraction):
            p.x_h2o_kind = A.NE_XH2O_SALINITY
            p.water_molar_mass = x.water_molar_mass
            for k in range(4):
                p.constituent_molar_mass[k] = x.molar_masses[k]
                p.constituent_mass_fraction[k] = x.mass_fractions[k]
        else:
            raise NoKernelVariantError(f"water_mole_fraction {x!r} has no kernel variant")
        v = self.velocity_formulation
        if isinstance(v, RelativeVelocity):
            p.velocity_formulation = A.NE_VEL_RELATIVE
        elif isinstance(v, WindVelocity):
            p.velocity_formulation = A.NE_VEL_WIND
        else:
            raise NoKernelVariantError(f"velocity formulation {v!r} has no kernel variant")
        t = self.temperature_formulation
        if isinstance(t, BulkTemperature):
            p.temperature_formulation = A.NE_TEMP_BULK
        elif isinstance(t, SkinTemperature):
            p.max_dT = float(t.max_ΔT)
            F = t.internal_flux
            if isinstance(F, DiffusiveFlux):
                p.delta = float(F.δ)
                if isinstance(F.κ, InteriorDiffusivity):
                    p.temperature_formulation = A.NE_TEMP_SKIN_DIFFUSIVE_INTERIOR
                    p.kappa = F.κ.minimum_diffusivity
                elif isinstance(F.κ, (int, float)):
                    p.temperature_formulation = A.NE_TEMP_SKIN_DIFFUSIVE
                    p.kappa = float(F.κ)
                else:
                    raise NoKernelVariantError(f"diffusivity {F.κ!r} has no kernel variant")
            elif isinstance(F, ConductiveFlux):
                p.temperature_formulation = A.NE_TEMP_SKIN_CONDUCTIVE
                p.ice_conductivity = F.conductivity
            elif isinstance(F, IceSnowConductiveFlux):
                p.temperature_formulation = A.NE_TEMP_SKIN_ICE_SNOW
                p.ice_conductivity, p.snow_conductivity = F.ice_conductivity, F.snow_conductivity
            else:
                raise NoKernelVariantError(f"internal flux {F!r} has no kernel variant")
        else:
            raise NoKernelVariantError(f"temperature formulation {t!r} has no kernel variant")
        return p


# ---------------------------------------------------------------------------------------------
# thermodynamics / media / radiation
# ---------------------------------------------------------------------------------------------
@dataclass
class AtmosphereThermodynamicsParameters:  # ../../Atmospheres/thermodynamic_parameters.jl:45-258
    FT: str = "f64"
    gas_constant: float = 8.3144598
    dry_air_molar_mass: float = 0.02897
    water_molar_mass: float = 0.018015
    dry_air_adiabatic_exponent: float = 2 / 7
    water_vapor_heat_capacity: float = 1859
    liquid_water_heat_capacity: float = 4181
    water_ice_heat_capacity: float = 2100
    reference_vaporization_enthalpy: float = 2500800
    reference_sublimation_enthalpy: float = 2834400
    reference_temperature: float = 273.16
    triple_point_temperature: float = 273.16
    triple_point_pressure: float = 611.657
    water_freezing_temperature: float = 273.15
    total_ice_nucleation_temperature: float = 233

    def pod(self) -> A.NeThermoParams:
        return A.NeThermoParams(
            dtype=A.NE_F64 if self.FT == "f64" else A.NE_F32, pad_=0,
            gas_constant=self.gas_constant, dry_air_molar_mass=self.dry_air_molar_mass,
            water_molar_mass=self.water_molar_mass, kappa_d=self.dry_air_adiabatic_exponent,
            cp_v=self.water_vapor_heat_capacity, cp_l=self.liquid_water_heat_capacity,
            cp_i=self.water_ice_heat_capacity, LH_v0=self.reference_vaporization_enthalpy,
            LH_s0=self.reference_sublimation_enthalpy, T_0=self.reference_temperature,
            T_triple=self.triple_point_temperature, press_triple=self.triple_point_pressure,
            T_freeze=self.water_freezing_temperature, T_icenuc=self.total_ice_nucleation_temperature)


@dataclass
class ElevationCorrection:  # atmosphere_state_correction.jl:39-59
    """ElevationCorrection(surface_elevation, atmosphere_elevation; lapse_rate = 6.5e-3): elevations are numbers or
    exchange-grid arrays (interior (ny, nx) or parent layout); g and Rᵈ come from the atmosphere when the correction
    is materialised on the exchange grid (atmosphere_state_correction.jl:89-110)."""
    surface_elevation: Any = 0.0
    atmosphere_elevation: Any = 0.0
    lapse_rate: float = 6.5e-3


class DegreesCelsius:  # ../components.jl:7
    pass


class DegreesKelvin:  # ../components.jl:8
    pass


@dataclass
class LinearLiquidus:  # ClimaSeaIce.SeaIceThermodynamics.LinearLiquidus (third party; defaults upstream)
    freshwater_melting_temperature: float = 0.0
    slope: float = 0.054


@dataclass
class MediumProperties:
    """ocean_properties / sea_ice_properties (component_interfaces.jl:460-480)."""
    reference_density: float = 1020.0
    heat_capacity: float = 3991.86795711963  # TEOS-10 reference heat capacity (third-party constant); pass explicitly
    temperature_units: Any = field(default_factory=DegreesCelsius)
    liquidus: LinearLiquidus = field(default_factory=LinearLiquidus)

    def pod(self) -> A.NeMediumProperties:
        m = A.NeMediumProperties()
        m.reference_density, m.heat_capacity = self.reference_density, self.heat_capacity
        m.temperature_units = A.NE_DEGREES_CELSIUS if isinstance(self.temperature_units, DegreesCelsius) else A.NE_DEGREES_KELVIN
        m.liquidus_slope = self.liquidus.slope
        m.liquidus_freshwater_melting_temperature = self.liquidus.freshwater_melting_temperature
        return m


@dataclass
class LatitudeDependentAlbedo:  # ../../Radiations/latitude_dependent_albedo.jl
    diffuse: float = 0.069
    direct: float = 0.011


@dataclass
class SeaIceAlbedo:  # ../../Radiations/sea_ice_albedo.jl:22-101 (CCSM3, Briegleb et al. 2004)
    ice_thickness: Any = None       # exchange-layout arrays of the sea-ice model
    snow_thickness: Any = None      # None: no snow model (get_snow_thickness(::Nothing) = 0)
    surface_temperature: Any = None
    ice_albedo: float = 0.54
    snow_albedo: float = 0.83
    ice_melt_reduction: float = 0.075
    snow_melt_reduction: float = 0.10
    melting_temperature: float = 0.0
    temperature_range: float = 1.0
    ocean_albedo: float = 0.06
    minimum_ice_thickness: float = 0.5
    minimum_snow_depth: float = 0.02


@dataclass
class TabulatedAlbedo:  # ../../Radiations/tabulated_albedo.jl:39-93
    """albedo(transmissivity, |latitude|) by bilinear lookup in `table` ((n_t, n_phi); the reference's default is the
    Payne (1972) table, which the caller supplies).  phi_values in radians, both axes with constant spacing."""
    table: Any = None               # device/host array of shape (n_phi, n_t) in C order == (n_t, n_phi) column-major
    phi_values: Any = None
    t_values: Any = None
    solar_constant: float = 1365.0
    day_to_radians: float = 2 * math.pi / 86400
    noon_in_seconds: int = 86400 // 2

    @staticmethod
    def clock_scalars(time_seconds):
        """simulation_day, seconds_in_day (:104-105) and the solar declination (:125-127) for a clock in seconds."""
        day = float(int(time_seconds // 86400)) if time_seconds >= 0 else -float(int((-time_seconds) // 86400))
        sec = time_seconds - day * 86400
        march_first = 80
        x = 360 * (day - march_first) / 365.25
        delta = math.radians((23 + 27 / 60) * math.sin(math.radians(math.fmod(x, 360.0))))
        return day, sec, delta


@dataclass
class SurfaceRadiationProperties:  # ../../Radiations/surface_radiation_properties.jl ; defaults prescribed_radiation.jl:65-72
    albedo: Any = 0.05
    emissivity: float = 0.97


# ---------------------------------------------------------------------------------------------
# sea-ice–ocean
# ---------------------------------------------------------------------------------------------
class MomentumBasedFrictionVelocity:  # friction_velocity.jl:24
    pass


@dataclass
class IceBathHeatFlux:  # sea_ice_ocean_heat_flux_formulations.jl:45-67
    heat_transfer_coefficient: float = 0.006
    friction_velocity: Any = 0.02


@dataclass
class ThreeEquationHeatFlux:  # :118-159
    heat_transfer_coefficient: float = 0.0095
    salt_transfer_coefficient: Optional[float] = None
    friction_velocity: Any = 0.002
    conductive_flux: Any = None          # ConductiveFlux => ConductiveFluxTEF (:198)
    internal_temperature: Any = None     # field (device array) when conductive_flux is set

    def __post_init__(self):
        if self.salt_transfer_coefficient is None:
            self.salt_transfer_coefficient = self.heat_transfer_coefficient / 35
