"""ctypes mirror of include/ne_b200.h (the C ABI of libne_b200.so).

Field order and types follow the header one to one; `tests/test_abi.py` checks every struct size
against the compiled library (`ne_struct_size`) and that every entry point the header declares is
exported.  Nothing here computes anything.
"""
import ctypes as C

NE_ABI_VERSION = 1

# error codes
NE_OK, NE_E_INVALID, NE_E_NO_VARIANT, NE_E_CUDA, NE_E_NO_DEVICE = 0, -1, -2, -3, -4
NE_F32, NE_F64 = 0, 1

NE_MAX_SUMMANDS = 4
NE_DIAG_MAX_FIELDS = 16

# stability functions
(NE_PSI_ZERO, NE_PSI_EDSON_MOMENTUM, NE_PSI_EDSON_SCALAR, NE_PSI_SHEBA_MOMENTUM, NE_PSI_SHEBA_SCALAR,
 NE_PSI_PAULSON_MOMENTUM, NE_PSI_PAULSON_SCALAR, NE_PSI_LINEAR_STABLE) = range(8)
NE_ROUGH_CONSTANT, NE_ROUGH_MOMENTUM, NE_ROUGH_SCALAR, NE_ROUGH_LAND = range(4)
NE_DISPLACEMENT_CONSTANT, NE_DISPLACEMENT_LAND = range(2)
NE_WAVE_CONSTANT, NE_WAVE_WIND_DEPENDENT = range(2)
NE_VISC_CONSTANT, NE_VISC_TEMPERATURE_DEPENDENT = range(2)
NE_SGS_NONE, NE_SGS_CONSTANT, NE_SGS_CONVECTIVE = range(3)
NE_PROFILE_LOGARITHMIC, NE_PROFILE_COARE = range(2)
NE_STOP_CONVERGENCE, NE_STOP_FIXED_ITERATIONS = range(2)
NE_COEFF_CONSTANT, NE_COEFF_POLYNOMIAL_DRAG = range(2)
NE_FLUX_SIMILARITY_THEORY, NE_FLUX_COEFFICIENT_BASED, NE_FLUX_LARGE_YEAGER = range(3)
NE_PHASE_LIQUID, NE_PHASE_ICE = range(2)
NE_XH2O_ONE, NE_XH2O_CONSTANT, NE_XH2O_SALINITY = range(3)
NE_VEL_RELATIVE, NE_VEL_WIND = range(2)
(NE_TEMP_BULK, NE_TEMP_SKIN_DIFFUSIVE, NE_TEMP_SKIN_DIFFUSIVE_INTERIOR, NE_TEMP_SKIN_CONDUCTIVE,
 NE_TEMP_SKIN_ICE_SNOW) = range(5)
NE_DEGREES_CELSIUS, NE_DEGREES_KELVIN = range(2)
NE_ALBEDO_CONSTANT, NE_ALBEDO_LATITUDE_DEPENDENT, NE_ALBEDO_FIELD, NE_ALBEDO_TABULATED, NE_ALBEDO_SEA_ICE = range(5)
NE_SIO_ICE_BATH, NE_SIO_THREE_EQUATION, NE_SIO_FREEZE_ONLY = range(3)
NE_USTAR_CONSTANT, NE_USTAR_MOMENTUM_BASED = range(2)
NE_REGION_BOX, NE_REGION_COLUMN = range(2)
NE_COLUMN_LINEAR, NE_COLUMN_NEAREST = range(2)

i32, i64, f64, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p


class NeSlot(C.Structure):
    _fields_ = [("ptr", vp), ("value", f64)]


class NeExchangeGrid(C.Structure):
    _fields_ = [("nx", i64), ("ny", i64), ("hx", i64), ("hy", i64),
                ("i_lo", i64), ("i_hi", i64), ("j_lo", i64), ("j_hi", i64)]


class NeTimeSeries(C.Structure):
    _fields_ = [("data", vp)]


class NeTimeInterp(C.Structure):
    _fields_ = [("frac", f64), ("frac_dtype", i32), ("m1", i32), ("m2", i32), ("same", i32)]


class NeInterpDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid), ("frac_i", vp), ("frac_j", vp), ("src_dtype", i32), ("n_fields", i32),
                ("src_nx", i64), ("src_ny", i64), ("src_hx", i64), ("src_hy", i64), ("src_nt", i64),
                ("time", NeTimeInterp), ("n_summands", i32 * 9), ("series", (NeTimeSeries * NE_MAX_SUMMANDS) * 9),
                ("out", vp * 9), ("potential", vp), ("potential_from", i32), ("ocean_reference_density", f64),
                ("rotation_cos", vp), ("rotation_sin", vp), ("rotate_u", i32), ("rotate_v", i32)]


class NeFracIndexDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid), ("nodes_2d", i32), ("lam", vp), ("phi", vp), ("src_dtype", i32),
                ("src_x_regular", i32), ("src_y_regular", i32), ("src_nx", i64), ("src_ny", i64),
                ("src_lam_nodes", vp), ("src_phi_nodes", vp), ("frac_i", vp), ("frac_j", vp)]


class NeThermoParams(C.Structure):
    _fields_ = [("dtype", i32), ("pad_", i32), ("gas_constant", f64), ("dry_air_molar_mass", f64),
                ("water_molar_mass", f64), ("kappa_d", f64), ("cp_v", f64), ("cp_l", f64), ("cp_i", f64),
                ("LH_v0", f64), ("LH_s0", f64), ("T_0", f64), ("T_triple", f64), ("press_triple", f64),
                ("T_freeze", f64), ("T_icenuc", f64)]


class NeStabilityFn(C.Structure):
    _fields_ = [("kind", i32), ("pad_", i32), ("p", f64 * 12)]


class NeStabilityProfile(C.Structure):
    _fields_ = [("split", i32), ("pad_", i32), ("a", NeStabilityFn), ("b", NeStabilityFn)]


class NeRoughnessLength(C.Structure):
    _fields_ = [("kind", i32), ("wave_kind", i32), ("visc_kind", i32), ("visc_dtype", i32), ("constant", f64),
                ("gravitational_acceleration", f64), ("wave_constant", f64), ("smooth_wall_parameter", f64),
                ("wave_Umax", f64), ("wave_C1", f64), ("wave_C2", f64), ("maximum_roughness_length", f64),
                ("nu", f64), ("nu_C", f64 * 4), ("reynolds_A", f64), ("reynolds_b", f64),
                ("land_multiplier", f64), ("land_minimum_roughness_length", f64)]


class NeSubgridVelocity(C.Structure):
    _fields_ = [("convective_kind", i32), ("mesoscale_kind", i32), ("composite", i32), ("pad_", i32),
                ("gustiness_parameter", f64), ("minimum_gustiness", f64), ("convective_constant", f64),
                ("mesoscale_constant", f64)]


class NeStopCriteria(C.Structure):
    _fields_ = [("kind", i32), ("maxiter", i32), ("tolerance", f64)]


class NePolynomialDrag(C.Structure):
    _fields_ = [("a", f64), ("b", f64), ("c", f64), ("d", f64), ("high_wind_speed_threshold", f64),
                ("high_wind_drag_coefficient", f64), ("minimum_wind_speed", f64)]


class NeTransferCoefficient(C.Structure):
    _fields_ = [("kind", i32), ("pad_", i32), ("constant", f64), ("polynomial", NePolynomialDrag)]


class NeLargeYeager(C.Structure):
    _fields_ = [("von_karman_constant", f64), ("neutral_drag", NePolynomialDrag),
                ("psi_momentum", NeStabilityProfile), ("psi_temperature", NeStabilityProfile),
                ("reference_height", f64), ("stable_heat", f64), ("unstable_heat", f64), ("moisture", f64)]


class NeFluxFormulation(C.Structure):
    _fields_ = [("kind", i32), ("similarity_form", i32), ("von_karman_constant", f64),
                ("subgrid_velocities", NeSubgridVelocity),
                ("psi_momentum", NeStabilityProfile), ("psi_temperature", NeStabilityProfile),
                ("psi_water_vapor", NeStabilityProfile),
                ("ell_momentum", NeRoughnessLength), ("ell_temperature", NeRoughnessLength),
                ("ell_water_vapor", NeRoughnessLength),
                ("zero_plane_displacement", f64), ("zero_plane_displacement_kind", i32), ("pad_", i32),
                ("coefficients", NeTransferCoefficient * 3),
                ("large_yeager", NeLargeYeager), ("stop", NeStopCriteria)]


class NeInterfaceProperties(C.Structure):
    _fields_ = [("phase", i32), ("x_h2o_kind", i32), ("velocity_formulation", i32), ("temperature_formulation", i32),
                ("x_h2o", f64), ("water_molar_mass", f64), ("constituent_molar_mass", f64 * 4),
                ("constituent_mass_fraction", f64 * 4), ("max_dT", f64), ("kappa", f64), ("delta", f64),
                ("ice_conductivity", f64), ("snow_conductivity", f64)]


class NeMediumProperties(C.Structure):
    _fields_ = [("reference_density", f64), ("heat_capacity", f64), ("temperature_units", i32), ("pad_", i32),
                ("liquidus_slope", f64), ("liquidus_freshwater_melting_temperature", f64)]


class NeSeaIceAlbedo(C.Structure):
    _fields_ = [("ice_albedo", f64), ("snow_albedo", f64), ("ice_melt_reduction", f64), ("snow_melt_reduction", f64),
                ("melting_temperature", f64), ("temperature_range", f64), ("ocean_albedo", f64),
                ("minimum_ice_thickness", f64), ("minimum_snow_depth", f64),
                ("ice_thickness", vp), ("snow_thickness", vp), ("surface_temperature", vp)]


class NeTabulatedAlbedo(C.Structure):
    _fields_ = [("table", vp), ("n_t", i32), ("n_phi", i32), ("t_values", f64 * 2), ("phi_values", f64 * 2),
                ("solar_constant", f64), ("day_to_radians", f64), ("noon_in_seconds", f64), ("seconds_in_day", f64),
                ("declination", f64), ("longitude", vp)]


class NeSurfaceRadiation(C.Structure):
    _fields_ = [("enabled", i32), ("albedo_kind", i32), ("stefan_boltzmann_constant", f64), ("albedo", f64),
                ("albedo_direct", f64), ("albedo_field", vp), ("latitude", vp), ("nodes_2d", i32), ("pad_", i32),
                ("emissivity", f64), ("downwelling_shortwave", vp), ("downwelling_longwave", vp),
                ("sea_ice_albedo", NeSeaIceAlbedo), ("tabulated_albedo", NeTabulatedAlbedo)]


class NeAtmosOceanDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid),
                ("ua", vp), ("va", vp), ("Ta", vp), ("pa", vp), ("qa", vp),
                ("surface_layer_height", NeSlot), ("boundary_layer_height", NeSlot),
                ("uo", NeSlot), ("vo", NeSlot), ("To", NeSlot), ("So", NeSlot),
                ("kappa", vp), ("inactive", vp),
                ("radiation", NeSurfaceRadiation), ("thermo", NeThermoParams),
                ("gravitational_acceleration", f64), ("flux", NeFluxFormulation),
                ("properties", NeInterfaceProperties), ("ocean", NeMediumProperties),
                ("latent_heat", vp), ("sensible_heat", vp), ("water_vapor", vp), ("x_momentum", vp), ("y_momentum", vp),
                ("interface_temperature", vp), ("friction_velocity", vp), ("temperature_scale", vp),
                ("water_vapor_scale", vp), ("iterations", vp)]


class NeAtmosSeaIceDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid),
                ("ua", vp), ("va", vp), ("Ta", vp), ("pa", vp), ("qa", vp),
                ("surface_layer_height", NeSlot), ("boundary_layer_height", NeSlot),
                ("To", NeSlot), ("So", NeSlot),
                ("hi", NeSlot), ("hs", NeSlot), ("hc", NeSlot), ("concentration", NeSlot),
                ("inactive", vp),
                ("radiation", NeSurfaceRadiation), ("thermo", NeThermoParams),
                ("gravitational_acceleration", f64), ("flux", NeFluxFormulation),
                ("properties", NeInterfaceProperties), ("ocean", NeMediumProperties), ("sea_ice", NeMediumProperties),
                ("latent_heat", vp), ("sensible_heat", vp), ("water_vapor", vp), ("x_momentum", vp), ("y_momentum", vp),
                ("interface_temperature", vp), ("iterations", vp)]


NE_LANDQ_BULK, NE_LANDQ_FRACTIONAL_CRITICAL, NE_LANDQ_FRACTIONAL_CONSTANT, NE_LANDQ_SKIN, NE_LANDQ_DRY_LAYER = 0, 1, 2, 3, 4
NE_TORTUOSITY_CONSTANT, NE_TORTUOSITY_POWER_LAW = 0, 1


class NeLandHumidity(C.Structure):
    _fields_ = [("kind", i32), ("phase", i32), ("critical_saturation", f64), ("efficiency", f64),
                ("surface_thickness", f64), ("vapor_diffusivity", f64),
                ("maximum_dry_layer_depth", f64), ("dry_layer_onset_saturation", f64), ("dry_layer_exponent", f64),
                ("minimum_dry_layer_depth", f64), ("molecular_diffusivity", f64), ("wet_transition_width", f64),
                ("thermal_exchange_depth", f64), ("porosity", f64), ("tortuosity", i32), ("pad_", i32)]


class NeAtmosLandDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid),
                ("ua", vp), ("va", vp), ("Ta", vp), ("pa", vp), ("qa", vp),
                ("surface_layer_height", NeSlot), ("boundary_layer_height", NeSlot),
                ("land_temperature", NeSlot), ("saturation", NeSlot),
                ("thermo", NeThermoParams), ("gravitational_acceleration", f64), ("flux", NeFluxFormulation),
                ("properties", NeInterfaceProperties), ("humidity", NeLandHumidity),
                ("latent_heat", vp), ("sensible_heat", vp), ("water_vapor", vp), ("x_momentum", vp), ("y_momentum", vp),
                ("interface_temperature", vp),
                ("friction_velocity", vp), ("temperature_scale", vp), ("water_vapor_scale", vp), ("iterations", vp),
                ("momentum_roughness_length", vp), ("scalar_roughness_length", vp), ("zero_plane_displacement", vp)]


class NeSeaIceOceanDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid), ("nz", i64), ("hz", i64), ("T", vp), ("S", vp), ("dz", vp), ("dt", f64),
                ("formulation", i32), ("friction_velocity_kind", i32),
                ("heat_transfer_coefficient", f64), ("salt_transfer_coefficient", f64), ("friction_velocity", f64),
                ("has_conductive_flux", i32), ("pad_", i32), ("conductivity", f64), ("internal_temperature", vp),
                ("latent_heat", f64), ("ocean", NeMediumProperties),
                ("hi", NeSlot), ("hc", NeSlot), ("concentration", NeSlot), ("ice_salinity", NeSlot),
                ("ice_mass_flux", NeSlot), ("snow_mass_flux", NeSlot),
                ("x_momentum_in", vp), ("y_momentum_in", vp),
                ("frazil_heat", vp), ("interface_heat", vp), ("salt", vp), ("freshwater", vp),
                ("interface_temperature", vp), ("interface_salinity", vp)]


class NeSeaIceOceanStressDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid), ("ui", vp), ("vi", vp), ("uo", vp), ("vo", vp),
                ("ocean_density", f64), ("drag_coefficient", f64), ("x_momentum", vp), ("y_momentum", vp)]


class NeAssembleOceanDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid),
                ("sensible_heat", NeSlot), ("latent_heat", NeSlot), ("water_vapor", NeSlot),
                ("x_momentum_ao", NeSlot), ("y_momentum_ao", NeSlot),
                ("interface_heat", NeSlot), ("salt_io", NeSlot), ("freshwater_io", NeSlot),
                ("x_momentum_io", NeSlot), ("y_momentum_io", NeSlot),
                ("ocean_surface_temperature", NeSlot), ("concentration", NeSlot), ("rainfall", NeSlot),
                ("snowfall", NeSlot), ("intercepted_snowfall", NeSlot), ("land_freshwater", NeSlot),
                ("inactive", vp), ("ocean", NeMediumProperties),
                ("tau_x", vp), ("tau_y", vp), ("JT", vp), ("JS", vp), ("Jw", vp), ("JH", vp)]


class NeAssembleSeaIceDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid),
                ("sensible_heat", NeSlot), ("latent_heat", NeSlot), ("x_momentum", NeSlot), ("y_momentum", NeSlot),
                ("frazil_heat", NeSlot), ("interface_heat", NeSlot), ("snowfall", NeSlot), ("concentration", NeSlot),
                ("inactive", vp),
                ("top_heat", vp), ("top_snowfall", vp), ("top_u", vp), ("top_v", vp), ("bottom_heat", vp)]


class NeApplyRadiationDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid), ("radiation", NeSurfaceRadiation), ("concentration", NeSlot),
                ("surface_temperature", vp), ("medium", NeMediumProperties), ("inactive", vp),
                ("over_sea_ice", i32), ("two_color", i32), ("heat_flux", vp), ("two_color_surface_flux", vp),
                ("upwelling_longwave", vp), ("downwelling_longwave", vp), ("downwelling_shortwave", vp)]


class NeDiagDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid), ("n_fields", i32), ("pad_", i32), ("fields", vp * NE_DIAG_MAX_FIELDS),
                ("area", vp), ("inactive", vp), ("partial", vp), ("n_blocks", i64), ("result", vp)]


class NeFusedStepDesc(C.Structure):
    _fields_ = [("atmosphere", NeInterpDesc), ("radiation", NeInterpDesc), ("ao", NeAtmosOceanDesc),
                ("assemble", NeAssembleOceanDesc), ("apply_radiation", NeApplyRadiationDesc), ("diag", NeDiagDesc)]


class NeElevationCorrectionDesc(C.Structure):
    _fields_ = [("grid", NeExchangeGrid), ("T", vp), ("p", vp), ("elevation_difference", vp), ("lapse_rate", f64),
                ("gravitational_acceleration", f64), ("dry_air_gas_constant", f64)]


NE_HOST_MAX_FIELDS = 8


class NeHostField(C.Structure):
    _fields_ = [("host", vp), ("device", vp)]


class NeHostStepDesc(C.Structure):
    _fields_ = [("step", NeFusedStepDesc), ("n_fields", i32), ("n_chunks", i32), ("fields", NeHostField * NE_HOST_MAX_FIELDS),
                ("row_bytes", i64), ("n_out_fields", i32), ("pad_", i32), ("out_fields", NeHostField * NE_HOST_MAX_FIELDS)]


NE_RING_MAX_SERIES = 16
NE_MANGLE_NONE, NE_MANGLE_SHIFT_SOUTH, NE_MANGLE_AVERAGE_NORTH_SOUTH = range(3)
NE_CONV_NONE, NE_CONV_NEGATE, NE_CONV_ADD, NE_CONV_SUB, NE_CONV_MUL, NE_CONV_DIV, NE_CONV_MUL_DIV = range(7)


class NeSeriesRingDesc(C.Structure):
    _fields_ = [("n_series", i32), ("n_slots", i32), ("dtype", i32), ("periodic_x", i32),
                ("nx", i64), ("ny", i64), ("hx", i64), ("hy", i64),
                ("ring", vp * NE_RING_MAX_SERIES), ("conv_kind", i32 * NE_RING_MAX_SERIES),
                ("conv_a", C.c_double * NE_RING_MAX_SERIES), ("conv_b", C.c_double * NE_RING_MAX_SERIES),
                ("has_missing", i32 * NE_RING_MAX_SERIES), ("missing_value", C.c_double * NE_RING_MAX_SERIES),
                ("raw_nx", i64), ("raw_ny", i64), ("di", i64), ("dj", i64), ("mangling", i32 * NE_RING_MAX_SERIES),
                ("region_kind", i32), ("column_interpolation", i32),
                ("col_i_minus", i64), ("col_i_plus", i64), ("col_j_minus", i64), ("col_j_plus", i64),
                ("col_wx", C.c_double), ("col_wy", C.c_double)]


STRUCTS = {c.__name__: c for c in [
    NeSlot, NeExchangeGrid, NeTimeSeries, NeTimeInterp, NeInterpDesc, NeFracIndexDesc, NeThermoParams, NeStabilityFn,
    NeStabilityProfile, NeRoughnessLength, NeSubgridVelocity, NeStopCriteria, NePolynomialDrag, NeTransferCoefficient,
    NeLargeYeager, NeFluxFormulation, NeInterfaceProperties, NeMediumProperties, NeSeaIceAlbedo, NeTabulatedAlbedo, NeSurfaceRadiation, NeAtmosOceanDesc,
    NeAtmosSeaIceDesc, NeLandHumidity, NeAtmosLandDesc, NeSeaIceOceanDesc, NeSeaIceOceanStressDesc, NeAssembleOceanDesc, NeAssembleSeaIceDesc,
    NeApplyRadiationDesc, NeElevationCorrectionDesc, NeFusedStepDesc, NeDiagDesc, NeHostField, NeHostStepDesc, NeSeriesRingDesc]}

# entry points declared in include/ne_b200.h: name -> descriptor struct (None: special signature)
DESC_ENTRY_POINTS = {
    "ne_frac_indices": NeFracIndexDesc,
    "ne_interp_state": NeInterpDesc,
    "ne_correct_atmosphere_elevation": NeElevationCorrectionDesc,
    "ne_atmosphere_ocean_fluxes": NeAtmosOceanDesc,
    "ne_atmosphere_sea_ice_fluxes": NeAtmosSeaIceDesc,
    "ne_atmosphere_land_fluxes": NeAtmosLandDesc,
    "ne_sea_ice_ocean_fluxes": NeSeaIceOceanDesc,
    "ne_sea_ice_ocean_stress": NeSeaIceOceanStressDesc,
    "ne_assemble_net_ocean_fluxes": NeAssembleOceanDesc,
    "ne_assemble_net_sea_ice_fluxes": NeAssembleSeaIceDesc,
    "ne_apply_radiative_fluxes": NeApplyRadiationDesc,
    "ne_fused_interface_step": NeFusedStepDesc,
    "ne_interp_and_ao_fluxes": NeFusedStepDesc,
    "ne_diag_reduce": NeDiagDesc,
}
OTHER_ENTRY_POINTS = ["ne_version", "ne_last_error", "ne_device_count", "ne_memcpy_h2d", "ne_memcpy_d2h",
                      "ne_stream_synchronize", "ne_measure_fp64_peak", "ne_struct_size", "ne_host_pipeline_create",
                      "ne_host_pipeline_destroy", "ne_host_pipelined_step_f64", "ne_host_pipelined_step_f32",
                      "ne_series_ring_create", "ne_series_ring_destroy", "ne_series_ring_load", "ne_series_ring_acquire",
                      "ne_series_ring_release", "ne_count_solve_ops_f64", "ne_diag_allreduce_f64"]


def all_entry_points():
    names = []
    for base in DESC_ENTRY_POINTS:
        names += [base + "_f64", base + "_f32"]
    return names + OTHER_ENTRY_POINTS
