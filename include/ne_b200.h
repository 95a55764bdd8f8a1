/*
 * ne_b200.h — C ABI of libne_b200.so: the atmosphere–surface interface path of
 * NumericalEarth.jl (PrescribedAtmosphere/Radiation interpolation onto the exchange
 * grid + similarity-theory / coefficient-based turbulent flux solve + sea-ice kernels +
 * net-flux assembly + radiative flux application) as hand-written sm_100a CUDA kernels.
 *
 * Boundary rules
 *  - plain C: PODs, raw DEVICE pointers, sizes; no torch / CUDA C++ types. `stream` is a
 *    cudaStream_t passed as void* (Julia: CUDA.stream().handle).
 *  - every call enqueues on `stream` and returns without synchronising (the reference's
 *    update_state! never synchronises: src/EarthSystemModels/time_step_earth_system_model.jl:38-83).
 *  - the library never allocates or frees caller arrays; descriptors are copied at call time.
 *  - return 0 on success, a negative NE_E* code otherwise; ne_last_error() gives the
 *    thread-local message. A plugin configuration with no kernel variant returns
 *    NE_E_NO_VARIANT — there is NO CPU fallback.
 *
 * All reference citations are relative to /root/reference/.
 *
 * Array layout. Every exchange-grid Field{Center,Center,Nothing} is the dense column-major
 * parent array (nx+2hx) x (ny+2hy) x 1 of an Oceananigans Field; reference index (i, j)
 * (1-based interior, 0 and N+1 = first halo ring) lives at
 *     base[(i + hx - 1) + (j + hy - 1) * (nx + 2*hx)].
 * Exchange arrays have the element type named by the entry point suffix (_f64 / _f32).
 */
#ifndef NE_B200_H
#define NE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NE_ABI_VERSION 1

/* ---- error codes ------------------------------------------------------------------- */
enum {
  NE_OK = 0,
  NE_E_INVALID = -1,     /* malformed descriptor (null pointer, bad size)                  */
  NE_E_NO_VARIANT = -2,  /* plugin type with no kernel variant (user closure etc.)          */
  NE_E_CUDA = -3,        /* CUDA runtime error (message holds cudaGetErrorString)           */
  NE_E_NO_DEVICE = -4    /* no CUDA device / extension built without device code            */
};

enum { NE_F32 = 0, NE_F64 = 1 };

/* ---- generic slots ------------------------------------------------------------------ */

/* {array | constant}: absent components arrive as ZeroField()/ConstantField/Number
 * (src/EarthSystemModels/components.jl:37-50, src/Oceans/slab_ocean.jl:104-116).
 * ptr != NULL: exchange-layout array of the entry point's element type; else `value`. */
typedef struct NeSlot {
  const void* ptr;
  double value;
} NeSlot;

/* Exchange grid parent layout + the kernel launch range.
 * interface_kernel_parameters (src/EarthSystemModels/InterfaceComputations/InterfaceComputations.jl:100-116)
 * gives (0:Nx+1)x(0:Ny+1); the `:xy` kernels use (1:Nx)x(1:Ny).  Latitude-band shards pass
 * their local ny and the same ranges relative to the band. */
typedef struct NeExchangeGrid {
  int64_t nx, ny;   /* interior size  */
  int64_t hx, hy;   /* halo size      */
  int64_t i_lo, i_hi, j_lo, j_hi; /* inclusive launch range, reference indices */
} NeExchangeGrid;

/* ---- interpolation (src/Atmospheres/interpolate_atmospheric_state.jl:9-182,
 *                      src/Radiations/interpolate_radiation_state.jl:4-69) --------------- */

/* One FieldTimeSeries `.data` parent: (nx+2hx) x (ny+2hy) x 1 x nt, column-major, element
 * type `dtype` (JRA55 is Float32: test/test_jra55.jl:40). data == NULL means the series is
 * `nothing` and contributes the literal 0 (interpolate_atmospheric_state.jl:143). */
typedef struct NeTimeSeries {
  const void* data;
} NeTimeSeries;

/* Host-computed TimeInterpolator (interpolate_atmospheric_state.jl:57-60): fractional time
 * index and the two in-memory slots (1-based along the 4th dimension, i.e. already passed
 * through memory_index). `same` = (n1 == n2) on the ORIGINAL time indices. */
typedef struct NeTimeInterp {
  double frac;      /* n-tilde                                                   */
  int32_t frac_dtype; /* NE_F32 / NE_F64: element type n-tilde has in the reference */
  int32_t m1, m2;   /* memory slots, 1-based                                      */
  int32_t same;     /* n1 == n2  => return psi1                                    */
} NeTimeInterp;

#define NE_MAX_SUMMANDS 4

typedef struct NeInterpDesc {
  NeExchangeGrid grid;
  /* fractional indices written by initialize! (prescribed_atmosphere_regridder.jl:41-71);
   * exchange layout, element type `src_dtype` (the ATMOSPHERE grid's eltype, :32-35).
   * NULL => Flat direction => FractionalIndices component `nothing` => (1,1,0).  */
  const void* frac_i;
  const void* frac_j;
  int32_t src_dtype;          /* element type of frac_i/frac_j and of every series      */
  int32_t n_fields;           /* number of output fields (7 atmosphere, 2 radiation)   */
  int64_t src_nx, src_ny, src_hx, src_hy, src_nt; /* shared by all series              */
  NeTimeInterp time;
  /* field f = sum over its summands (tuple-valued precipitation,
   * interpolate_atmospheric_state.jl:152-182); n_summands[f] in 0..NE_MAX_SUMMANDS     */
  int32_t n_summands[9];
  NeTimeSeries series[9][NE_MAX_SUMMANDS];
  void* out[9];               /* exchange-layout outputs (element type of the suffix)  */
  /* barotropic potential = p / rho_ocean (interpolate_atmospheric_state.jl:80-85):
   * optional; written over the same launch range from field index `potential_from`.     */
  void* potential;
  int32_t potential_from;
  double ocean_reference_density;
  /* intrinsic_vector (interpolate_atmospheric_state.jl:123-126): on a rotated exchange grid (tripolar, cubed sphere)
   * the interpolated vector (fields rotate_u, rotate_v) is turned into the grid's frame before it is stored:
   * u' = u cos(theta) + v sin(theta), v' = -u sin(theta) + v cos(theta).  Exchange-layout arrays of the exchange
   * element type holding the grid's rotation angle (Oceananigans rotation metrics, third party: the binding fills them
   * once).  NULL => latitude-longitude exchange grid: the vector is stored as interpolated.                         */
  const void* rotation_cos;
  const void* rotation_sin;
  int32_t rotate_u, rotate_v;
} NeInterpDesc;

/* Fractional indices of exchange nodes on a LatitudeLongitude source grid
 * (prescribed_atmosphere_regridder.jl:51-71 -> Oceananigans FractionalIndices, third party).
 * Exchange nodes: 1-D axes (lat-lon exchange grid) or 2-D exchange-layout arrays. */
typedef struct NeFracIndexDesc {
  NeExchangeGrid grid;
  int32_t nodes_2d;           /* 0: lam[(nx+2hx)], phi[(ny+2hy)]; 1: both exchange layout */
  const void* lam;            /* degrees, exchange element type                           */
  const void* phi;
  int32_t src_dtype;          /* element type of the source grid and of the outputs      */
  int32_t src_x_regular, src_y_regular;
  int64_t src_nx, src_ny;
  const void* src_lam_nodes;  /* src_nx center nodes (device, src_dtype)                 */
  const void* src_phi_nodes;  /* src_ny center nodes                                     */
  void* frac_i;               /* exchange layout, src_dtype                              */
  void* frac_j;
} NeFracIndexDesc;

/* ---- plugin types -> POD variants ----------------------------------------------------- */

/* AtmosphereThermodynamicsParameters (src/Atmospheres/thermodynamic_parameters.jl:30-258). */
typedef struct NeThermoParams {
  int32_t dtype;  /* eltype(thermodynamics_parameters) = CT (interface_states.jl:56-59)     */
  int32_t pad_;
  double gas_constant, dry_air_molar_mass, water_molar_mass;
  double kappa_d, cp_v, cp_l, cp_i;
  double LH_v0, LH_s0, T_0, T_triple, press_triple, T_freeze, T_icenuc;
} NeThermoParams;

/* stability functions (similarity_theory_turbulent_fluxes.jl:487-752) */
enum {
  NE_PSI_ZERO = 0,            /* Returns(zero(FT)) :201-204                                 */
  NE_PSI_EDSON_MOMENTUM = 1,  /* p = zmax,A+,B+,C+,D+,A-,B-,C-,D-,E-,F-           :487-532 */
  NE_PSI_EDSON_SCALAR = 2,    /* p = zmax,A+,B+,C+,D+,E+,A-,B-,C-,D-,E-,F-        :571-618 */
  NE_PSI_SHEBA_MOMENTUM = 3,  /* p = a,b                                          :637-657 */
  NE_PSI_SHEBA_SCALAR = 4,    /* p = a,b,c                                        :659-677 */
  NE_PSI_PAULSON_MOMENTUM = 5,/* p = a,b                                          :683-699 */
  NE_PSI_PAULSON_SCALAR = 6,  /* p = a                                            :701-710 */
  NE_PSI_LINEAR_STABLE = 7    /* p = coefficient, maximum_stability_parameter     :742-752 */
};
typedef struct NeStabilityFn {
  int32_t kind;
  int32_t pad_;
  double p[12];
} NeStabilityFn;
/* split != 0: SplitStabilityFunction(stable = a, unstable = b) :712-725 */
typedef struct NeStabilityProfile {
  int32_t split;
  int32_t pad_;
  NeStabilityFn a, b;
} NeStabilityProfile;

/* roughness lengths (roughness_lengths.jl) */
enum { NE_ROUGH_CONSTANT = 0, NE_ROUGH_MOMENTUM = 1, NE_ROUGH_SCALAR = 2,
       NE_ROUGH_LAND = 3 };   /* LandRoughnessLength :21-39: resolved per cell from the land model's roughness field
                                 (local_roughness_length, similarity_theory_turbulent_fluxes.jl:265-278), then a constant */
enum { NE_WAVE_CONSTANT = 0, NE_WAVE_WIND_DEPENDENT = 1 };          /* :56-75  */
enum { NE_VISC_CONSTANT = 0, NE_VISC_TEMPERATURE_DEPENDENT = 1 };   /* :149-189 */
typedef struct NeRoughnessLength {
  int32_t kind;
  int32_t wave_kind;
  int32_t visc_kind;
  int32_t visc_dtype;  /* constant nu keeps its own type (Float64 literal by default, :94,126) */
  double constant;     /* NE_ROUGH_CONSTANT                                                    */
  double gravitational_acceleration, wave_constant, smooth_wall_parameter; /* MOMENTUM :1-7   */
  double wave_Umax, wave_C1, wave_C2;
  double maximum_roughness_length;
  double nu;           /* constant viscosity                                                   */
  double nu_C[4];      /* TemperatureDependentAirViscosity C0..C3                              */
  double reynolds_A, reynolds_b; /* ReynoldsScalingFunction :212-231                           */
  double land_multiplier, land_minimum_roughness_length; /* NE_ROUGH_LAND: max(multiplier max(field, minimum), minimum);
                                                            no field (the land model provides none): the minimum stands in */
} NeRoughnessLength;

/* subgrid velocities (similarity_theory_turbulent_fluxes.jl:45-98) */
enum { NE_SGS_NONE = 0, NE_SGS_CONSTANT = 1, NE_SGS_CONVECTIVE = 2 };
typedef struct NeSubgridVelocity {
  int32_t convective_kind;   /* slot used when not composite, or the `convective` slot      */
  int32_t mesoscale_kind;    /* NONE / CONSTANT; only read when composite != 0              */
  int32_t composite;         /* SubgridVelocityCorrection :69-98                            */
  int32_t pad_;
  double gustiness_parameter, minimum_gustiness; /* ConvectiveGustiness                    */
  double convective_constant, mesoscale_constant;
} NeSubgridVelocity;

enum { NE_PROFILE_LOGARITHMIC = 0, NE_PROFILE_COARE = 1 };           /* :239-253 */
enum { NE_STOP_CONVERGENCE = 0, NE_STOP_FIXED_ITERATIONS = 1 };      /* compute_interface_state.jl:5-26 */
typedef struct NeStopCriteria {
  int32_t kind;
  int32_t maxiter;     /* or the fixed iteration count */
  double tolerance;
} NeStopCriteria;

/* coefficient-based fluxes (coefficient_based_turbulent_fluxes.jl) */
enum { NE_COEFF_CONSTANT = 0, NE_COEFF_POLYNOMIAL_DRAG = 1 };
typedef struct NePolynomialDrag {   /* :20-54 */
  double a, b, c, d, high_wind_speed_threshold, high_wind_drag_coefficient, minimum_wind_speed;
} NePolynomialDrag;
typedef struct NeTransferCoefficient {
  int32_t kind;
  int32_t pad_;
  double constant;
  NePolynomialDrag polynomial;
} NeTransferCoefficient;
typedef struct NeLargeYeager {      /* :82-108, 288-340 */
  double von_karman_constant;
  NePolynomialDrag neutral_drag;
  NeStabilityProfile psi_momentum, psi_temperature;
  double reference_height, stable_heat, unstable_heat, moisture;
} NeLargeYeager;

enum { NE_FLUX_SIMILARITY_THEORY = 0, NE_FLUX_COEFFICIENT_BASED = 1, NE_FLUX_LARGE_YEAGER = 2 };
enum { NE_DISPLACEMENT_CONSTANT = 0, NE_DISPLACEMENT_LAND = 1 };
typedef struct NeFluxFormulation {
  int32_t kind;
  int32_t similarity_form;
  /* SimilarityTheoryFluxes :9-18 */
  double von_karman_constant;
  NeSubgridVelocity subgrid_velocities;
  NeStabilityProfile psi_momentum, psi_temperature, psi_water_vapor;
  NeRoughnessLength ell_momentum, ell_temperature, ell_water_vapor;
  double zero_plane_displacement;
  int32_t zero_plane_displacement_kind;  /* NE_DISPLACEMENT_*: a Number | LandZeroPlaneDisplacement() (roughness_lengths.jl:44-51),
                                            read per cell from the land model's field, 0 without one (:296-303)         */
  int32_t pad_;
  /* CoefficientBasedFluxes :142-145 (tuple/SimilarityScales of constants or polynomial) */
  NeTransferCoefficient coefficients[3];
  NeLargeYeager large_yeager;
  NeStopCriteria stop;
} NeFluxFormulation;

/* InterfaceProperties (interface_states.jl:8-12) */
enum { NE_PHASE_LIQUID = 0, NE_PHASE_ICE = 1 };
enum { NE_XH2O_ONE = 0, NE_XH2O_CONSTANT = 1, NE_XH2O_SALINITY = 2 };  /* :44-47, 236-277 */
enum { NE_VEL_RELATIVE = 0, NE_VEL_WIND = 1 };                          /* :284-301 */
enum {
  NE_TEMP_BULK = 0,                 /* :330-333 */
  NE_TEMP_SKIN_DIFFUSIVE = 1,       /* DiffusiveFlux(kappa, delta) :371-374, 434-457 */
  NE_TEMP_SKIN_DIFFUSIVE_INTERIOR = 2, /* InteriorDiffusivity :384-391 (kappa array)  */
  NE_TEMP_SKIN_CONDUCTIVE = 3,      /* ClimaSeaIce ConductiveFlux :511-516            */
  NE_TEMP_SKIN_ICE_SNOW = 4         /* IceSnowConductiveFlux :519-524                 */
};
typedef struct NeInterfaceProperties {
  int32_t phase;
  int32_t x_h2o_kind;
  int32_t velocity_formulation;
  int32_t temperature_formulation;
  double x_h2o;                      /* NE_XH2O_CONSTANT (default 0.98, component_interfaces.jl:353-358) */
  double water_molar_mass;           /* WaterMoleFraction :236-253 */
  double constituent_molar_mass[4], constituent_mass_fraction[4];
  double max_dT;                     /* SkinTemperature.max_ΔT :355-360 */
  double kappa, delta;               /* DiffusiveFlux; kappa = minimum_diffusivity for INTERIOR */
  double ice_conductivity, snow_conductivity;
} NeInterfaceProperties;

enum { NE_DEGREES_CELSIUS = 0, NE_DEGREES_KELVIN = 1 };    /* components.jl:7-15 */
typedef struct NeMediumProperties {  /* ocean_properties / sea_ice_properties */
  double reference_density, heat_capacity;
  int32_t temperature_units;
  int32_t pad_;
  double liquidus_slope, liquidus_freshwater_melting_temperature; /* ClimaSeaIce LinearLiquidus */
} NeMediumProperties;

/* radiation properties of one surface (src/Radiations/air_sea_interface_radiation_state.jl:4-39) */
enum { NE_ALBEDO_CONSTANT = 0, NE_ALBEDO_LATITUDE_DEPENDENT = 1, NE_ALBEDO_FIELD = 2,
       NE_ALBEDO_TABULATED = 3, NE_ALBEDO_SEA_ICE = 4 };

/* SeaIceAlbedo (CCSM3; src/Radiations/sea_ice_albedo.jl:22-133): scalars converted to the exchange
 * element type in the kernel; snow_thickness NULL => `nothing` => zero(grid) (:132). */
typedef struct NeSeaIceAlbedo {
  double ice_albedo, snow_albedo, ice_melt_reduction, snow_melt_reduction;
  double melting_temperature, temperature_range, ocean_albedo;
  double minimum_ice_thickness, minimum_snow_depth;
  const void* ice_thickness;       /* exchange layout */
  const void* snow_thickness;      /* exchange layout or NULL */
  const void* surface_temperature; /* exchange layout, the units of the sea-ice model (deg C) */
} NeSeaIceAlbedo;

/* TabulatedAlbedo (src/Radiations/tabulated_albedo.jl:39-160): albedo(transmissivity, |latitude|) by
 * bilinear table lookup.  The clock-dependent scalars (day, seconds in the day, solar declination
 * :118-131 — host `sind`) are evaluated by the caller once per step and passed in. */
typedef struct NeTabulatedAlbedo {
  const void* table;               /* (n_t, n_phi) column-major, exchange element type          */
  int32_t n_t, n_phi;
  double t_values[2];              /* first two tabulated transmissivities (constant spacing)   */
  double phi_values[2];            /* first two tabulated latitudes, radians                     */
  double solar_constant;           /* S0, default 1365                                          */
  double day_to_radians;           /* 2 pi / 86400                                              */
  double noon_in_seconds;          /* 43200                                                     */
  double seconds_in_day;           /* time - day*86400 (:115)                                   */
  double declination;              /* delta, radians, already converted to the table eltype     */
  const void* longitude;           /* lam[(nx+2hx)] degrees (nodes_2d: exchange layout)         */
} NeTabulatedAlbedo;

typedef struct NeSurfaceRadiation {
  int32_t enabled;          /* 0 => radiation === nothing => zero radiation state            */
  int32_t albedo_kind;
  double stefan_boltzmann_constant;
  double albedo;            /* constant; or `diffuse` of LatitudeDependentAlbedo             */
  double albedo_direct;     /* latitude_dependent_albedo.jl:48-53                            */
  const void* albedo_field; /* NE_ALBEDO_FIELD: exchange-layout array (pre-evaluated)        */
  const void* latitude;     /* phi[(ny+2hy)] degrees (nodes_2d: exchange layout), LATITUDE_DEPENDENT / TABULATED */
  int32_t nodes_2d;         /* 0: 1-D lat-lon axes; 1: exchange-layout node arrays (curvilinear grids) */
  int32_t pad_;
  double emissivity;
  const void* downwelling_shortwave;  /* exchange layout                                     */
  const void* downwelling_longwave;
  NeSeaIceAlbedo sea_ice_albedo;       /* NE_ALBEDO_SEA_ICE   */
  NeTabulatedAlbedo tabulated_albedo;  /* NE_ALBEDO_TABULATED */
} NeSurfaceRadiation;

/* ---- atmosphere–ocean turbulent fluxes (atmosphere_ocean_fluxes.jl:17-197) -------------- */
typedef struct NeAtmosOceanDesc {
  NeExchangeGrid grid;
  /* interpolated atmosphere state on the exchange grid (:97-103) */
  const void *ua, *va, *Ta, *pa, *qa;
  NeSlot surface_layer_height;    /* state2dindex(...) :107 */
  NeSlot boundary_layer_height;   /* h_bl :115 */
  /* ocean surface state: POINTERS TO THE TOP-LEVEL (k = Nz) PLANE of each 3-D parent (:62-71).
   * u on (Face,Center): uo = (u[i]+u[i+1])/2;  v on (Center,Face): vo = (v[j]+v[j+1])/2.     */
  NeSlot uo, vo, To, So;
  const void* kappa;              /* interior diffusivity plane, NE_TEMP_SKIN_DIFFUSIVE_INTERIOR */
  const uint8_t* inactive;        /* exchange layout, 1 = inactive_node(i,j,Nz) (:142); NULL = all active */
  NeSurfaceRadiation radiation;
  NeThermoParams thermo;
  double gravitational_acceleration;
  NeFluxFormulation flux;
  NeInterfaceProperties properties;
  NeMediumProperties ocean;
  /* outputs (:184-196) */
  void *latent_heat, *sensible_heat, *water_vapor, *x_momentum, *y_momentum;
  void *interface_temperature;
  void *friction_velocity, *temperature_scale, *water_vapor_scale;
  int32_t* iterations;            /* optional diagnostics: iterate_interface_state calls per point */
} NeAtmosOceanDesc;

/* ---- atmosphere–sea-ice turbulent fluxes (atmosphere_sea_ice_fluxes.jl:18-185) ---------- */
typedef struct NeAtmosSeaIceDesc {
  NeExchangeGrid grid;
  const void *ua, *va, *Ta, *pa, *qa;
  NeSlot surface_layer_height, boundary_layer_height;
  NeSlot To, So;                  /* ocean top-level planes (:92-94)                       */
  NeSlot hi, hs, hc, concentration; /* sea-ice state (:99-102)                             */
  const uint8_t* inactive;
  NeSurfaceRadiation radiation;
  NeThermoParams thermo;
  double gravitational_acceleration;
  NeFluxFormulation flux;
  NeInterfaceProperties properties;
  NeMediumProperties ocean, sea_ice;
  void *latent_heat, *sensible_heat, *water_vapor, *x_momentum, *y_momentum;
  void *interface_temperature;    /* READ-MODIFY-WRITE: top_surface_temperature (:103-104,183) */
  int32_t* iterations;
} NeAtmosSeaIceDesc;

/* ---- atmosphere–land turbulent fluxes (atmosphere_land_fluxes.jl:48-251) -----------------------
 * Same solver as the ocean / sea-ice kernels with a third interface state (AirLandInterfaceState,
 * interface_states.jl:740-777): zero surface velocity, the bulk land temperature as interface temperature
 * (BulkTemperature, the reference's default for land), and a land surface-humidity closure
 * (interface_states.jl:92-229, 585-651).  No masking; launch range `:xy` = (1:nx) x (1:ny).  Temperatures in Kelvin.
 * A depth diagnostic other than StorageBasedDryLayerDepth has no kernel variant: NE_E_NO_VARIANT.                */
enum { NE_LANDQ_BULK = 0,                 /* BulkHumidity: q_sat(T_s) where saturation > 0, else 0                  */
       NE_LANDQ_FRACTIONAL_CRITICAL = 1,  /* FractionalHumidity(CriticalSaturation): min(S / S_c, 1) q_sat(T_s)      */
       NE_LANDQ_FRACTIONAL_CONSTANT = 2,  /* FractionalHumidity(beta::Number)                                        */
       NE_LANDQ_SKIN = 3,                 /* SkinHumidity: soil vapor-flux balance, re-solved every trip (:600-651) */
       NE_LANDQ_DRY_LAYER = 4 };          /* DryLayerHumidity (dry_layer_humidity.jl:81-366): Fick flux through a dry surface
                                             layer of depth dv(S), blended into the saturated skin with a logistic weight      */
enum { NE_TORTUOSITY_CONSTANT = 0, NE_TORTUOSITY_POWER_LAW = 1 };   /* dry_layer_humidity.jl: Constant / Millington-Quirk */
typedef struct NeLandHumidity {
  int32_t kind;
  int32_t phase;                 /* NE_PHASE_LIQUID / NE_PHASE_ICE */
  double critical_saturation;    /* FRACTIONAL_CRITICAL */
  double efficiency;             /* FRACTIONAL_CONSTANT */
  double surface_thickness;      /* SKIN: saturation depth d         */
  double vapor_diffusivity;      /* SKIN: soil vapor diffusivity     */
  /* DRY_LAYER: StorageBasedDryLayerDepth, DryLayerVaporPistonVelocity, thermal exchange depth, porosity */
  double maximum_dry_layer_depth, dry_layer_onset_saturation, dry_layer_exponent;
  double minimum_dry_layer_depth, molecular_diffusivity, wet_transition_width;
  double thermal_exchange_depth, porosity;
  int32_t tortuosity;
  int32_t pad_;
} NeLandHumidity;
typedef struct NeAtmosLandDesc {
  NeExchangeGrid grid;
  const void *ua, *va, *Ta, *pa, *qa;
  NeSlot surface_layer_height, boundary_layer_height;
  NeSlot land_temperature;       /* exchanger.land.state.T: bulk land temperature, Kelvin (:182-184)                  */
  NeSlot saturation;             /* exchanger.land.state.saturation                                                  */
  NeThermoParams thermo;
  double gravitational_acceleration;
  NeFluxFormulation flux;        /* default: Large-Yeager stability functions, constant roughness 0.1 / 0.01 / 0.01 m
                                    (component_interfaces.jl:514-521)                                                 */
  NeInterfaceProperties properties; /* temperature_formulation must be NE_TEMP_BULK; velocity formulation as usual     */
  NeLandHumidity humidity;
  void *latent_heat, *sensible_heat, *water_vapor, *x_momentum, *y_momentum;
  void *interface_temperature;
  void *friction_velocity, *temperature_scale, *water_vapor_scale;
  int32_t* iterations;
  /* atmosphere_land_surface_properties(land_state) (:114-115): per-cell fields a land model may provide (SlabLand provides
   * none: NULL = `hasproperty` is false).  Read by NE_ROUGH_LAND roughness lengths (momentum slot: the first; temperature
   * and water-vapor slots: the second) and by NE_DISPLACEMENT_LAND.  Exchange dtype, exchange-grid layout.               */
  const void *momentum_roughness_length, *scalar_roughness_length, *zero_plane_displacement;
} NeAtmosLandDesc;

/* ---- sea-ice–ocean fluxes (sea_ice_ocean_fluxes.jl:20-226, freezing_limited_ocean_temperature.jl:73-118) */
enum { NE_SIO_ICE_BATH = 0, NE_SIO_THREE_EQUATION = 1, NE_SIO_FREEZE_ONLY = 2 };
enum { NE_USTAR_CONSTANT = 0, NE_USTAR_MOMENTUM_BASED = 1 };   /* friction_velocity.jl:24-44 */
typedef struct NeSeaIceOceanDesc {
  NeExchangeGrid grid;            /* launch range (1:nx)x(1:ny)                              */
  int64_t nz, hz;                 /* ocean column: parent (nx+2hx)(ny+2hy)(nz+2hz)           */
  void* T;                        /* 3-D ocean temperature parent, READ-MODIFY-WRITE (:172)  */
  const void* S;                  /* 3-D ocean salinity parent                               */
  const void* dz;                 /* nz cell thicknesses Δzᶜᶜᶜ(k), k = 1..nz (exchange dtype) */
  double dt;                      /* sea_ice.Δt; Inf at iteration 0 for FREEZE_ONLY (:88)    */
  int32_t formulation, friction_velocity_kind;
  double heat_transfer_coefficient, salt_transfer_coefficient, friction_velocity;
  int32_t has_conductive_flux;    /* ConductiveFluxTEF (heat_flux_formulations.jl:198-203)   */
  int32_t pad_;
  double conductivity;
  const void* internal_temperature;
  double latent_heat;             /* phase_transitions.reference_latent_heat                 */
  NeMediumProperties ocean;       /* liquidus in .liquidus_*                                 */
  NeSlot hi, hc, concentration, ice_salinity, ice_mass_flux, snow_mass_flux;
  const void *x_momentum_in, *y_momentum_in; /* si–o stresses read by MomentumBased u* (C,C via ℑ) */
  void *frazil_heat, *interface_heat, *salt, *freshwater;
  void *interface_temperature, *interface_salinity; /* written for THREE_EQUATION only      */
} NeSeaIceOceanDesc;

/* _compute_sea_ice_ocean_stress! (sea_ice_ocean_fluxes.jl:79-104) with ClimaSeaIce
 * SemiImplicitStress: tau = rho_e * C_D * |du| * du (pinned by test/test_surface_fluxes.jl:294-336) */
typedef struct NeSeaIceOceanStressDesc {
  NeExchangeGrid grid;
  const void *ui, *vi;            /* sea-ice velocities (Face,Face)-located exchange layout  */
  const void *uo, *vo;            /* ocean surface velocity planes                           */
  double ocean_density, drag_coefficient;
  void *x_momentum, *y_momentum;
} NeSeaIceOceanStressDesc;

/* ---- net flux assembly (src/Oceans/assemble_net_ocean_fluxes.jl:74-153,
 *                         src/SeaIces/assemble_net_sea_ice_fluxes.jl:42-81) ------------------ */
typedef struct NeAssembleOceanDesc {
  NeExchangeGrid grid;            /* (1:nx)x(1:ny) */
  NeSlot sensible_heat, latent_heat, water_vapor, x_momentum_ao, y_momentum_ao;
  NeSlot interface_heat, salt_io, freshwater_io, x_momentum_io, y_momentum_io; /* ZeroFluxes => constants */
  NeSlot ocean_surface_temperature, concentration, rainfall, snowfall, intercepted_snowfall, land_freshwater;
  const uint8_t* inactive;
  NeMediumProperties ocean;
  void *tau_x, *tau_y, *JT, *JS, *Jw, *JH;
} NeAssembleOceanDesc;

typedef struct NeAssembleSeaIceDesc {
  NeExchangeGrid grid;
  NeSlot sensible_heat, latent_heat, x_momentum, y_momentum;  /* atmosphere–sea-ice */
  NeSlot frazil_heat, interface_heat;                         /* sea-ice–ocean      */
  NeSlot snowfall, concentration;
  const uint8_t* inactive;
  void *top_heat, *top_snowfall, *top_u, *top_v, *bottom_heat;
} NeAssembleSeaIceDesc;

/* ---- radiative flux application (src/Radiations/apply_air_sea_radiative_fluxes.jl:62-111,
 *                                  apply_air_sea_ice_radiative_fluxes.jl:55-90) ------------- */
typedef struct NeApplyRadiationDesc {
  NeExchangeGrid grid;
  NeSurfaceRadiation radiation;
  NeSlot concentration;
  const void* surface_temperature;  /* a–o interface temperature / sea-ice top temperature */
  NeMediumProperties medium;        /* ocean_properties or sea_ice_properties               */
  const uint8_t* inactive;
  int32_t over_sea_ice;             /* surface: 0 ocean (JT += ...), 1 sea ice (top heat += ... * concentration),
                                       2 land (surface_energy_flux += ...; apply_air_land_radiative_fluxes.jl:64-97,
                                       surface_temperature = the atmosphere-land interface temperature, medium units Kelvin) */
  int32_t two_color;                /* ocean: route SW to TwoColorRadiation.surface_flux (src/Oceans/radiative_forcing.jl:84-91) */
  void* heat_flux;                  /* READ-MODIFY-WRITE: net_ocean_fluxes.T or top_heat_flux */
  void* two_color_surface_flux;
  void *upwelling_longwave, *downwelling_longwave, *downwelling_shortwave;
} NeApplyRadiationDesc;

/* ---- ElevationCorrection (phase 1.5 of update_state!,
 * src/EarthSystemModels/InterfaceComputations/atmosphere_state_correction.jl:39-146):
 * T <- T - G dz ; p <- p exp(-g dz / (Rd (T - G dz / 2))) in place on the regridded exchange state;
 * q is conserved.  dz = surface - atmosphere elevation, materialised on the exchange grid
 * (:89-110).  lapse_rate / g / Rd are converted to the exchange element type in the kernel (:135-143). */
typedef struct NeElevationCorrectionDesc {
  NeExchangeGrid grid;
  void* T;                       /* READ-MODIFY-WRITE, exchange layout */
  void* p;                       /* READ-MODIFY-WRITE */
  const void* elevation_difference;
  double lapse_rate;             /* default 6.5e-3 K/m (:49) */
  double gravitational_acceleration;
  double dry_air_gas_constant;   /* Parameters.R_d of the atmosphere's thermodynamics (:71) */
} NeElevationCorrectionDesc;

/* ---- diagnostics reduction (src/Diagnostics/interface_fluxes.jl:90-195; conservation sums) */
#define NE_DIAG_MAX_FIELDS 16
typedef struct NeDiagDesc {
  NeExchangeGrid grid;
  int32_t n_fields;
  int32_t pad_;
  const void* fields[NE_DIAG_MAX_FIELDS];
  const void* area;          /* exchange-layout cell areas (NULL => weight 1)     */
  const uint8_t* inactive;
  double* partial;           /* device scratch: n_blocks * n_fields doubles        */
  int64_t n_blocks;          /* fixed block count => deterministic summation order */
  double* result;            /* device, n_fields doubles                           */
} NeDiagDesc;

/* ---- fused interface step: interpolation -> a–o solve -> assembly -> radiation in ONE pass
 * (update_state! phases 1-4 for an OceanOnlyModel,
 *  src/EarthSystemModels/time_step_earth_system_model.jl:38-83).  Intermediate atmosphere
 * state arrays in `ao` may be NULL: they are then never materialised in HBM.  Phases 3-4 (+ the optional
 * diagnostics sums) run as ONE kernel when assemble / apply_radiation / diag share a launch range and
 * apply_radiation.heat_flux is assemble.JT. */
typedef struct NeFusedStepDesc {
  NeInterpDesc atmosphere;   /* 7 fields: u v T q p rain snow */
  NeInterpDesc radiation;    /* 2 fields: SW LW (n_fields = 0 => radiation off) */
  NeAtmosOceanDesc ao;
  NeAssembleOceanDesc assemble;
  NeApplyRadiationDesc apply_radiation;
  /* optional (n_fields = 0: off): area-weighted sums of `diag.fields` as ne_diag_reduce computes them, accumulated
   * by the same kernel that assembles the net fluxes and applies the radiation (one pass over the exchange grid
   * instead of three; same point -> block assignment, so the sums equal ne_diag_reduce's bit for bit).            */
  NeDiagDesc diag;
} NeFusedStepDesc;

/* ---- host-buffer step (ne_pipeline.cu): the ocean surface state arrives in pinned host memory every
 * coupled step; chunked H2D copies on the pipeline's own stream overlap the band-restricted kernels. */
#define NE_HOST_MAX_FIELDS 8
typedef struct NeHostField {
  const void* host;          /* pinned host memory, exchange-layout parent */
  void* device;              /* the device array the kernels read           */
} NeHostField;
typedef struct NeHostStepDesc {
  NeFusedStepDesc step;      /* full launch ranges; the pipeline restricts them per band     */
  int32_t n_fields;          /* ocean surface fields copied every step (T, S, u, v)           */
  int32_t n_chunks;          /* latitude chunks (<= the handle's capacity)                    */
  NeHostField fields[NE_HOST_MAX_FIELDS];
  int64_t row_bytes;         /* (nx + 2 hx) * sizeof(exchange element)                        */
  /* results back to the host (optional): the parent rows of each out_field leave device -> pinned host on a second
   * copy stream as soon as the band that writes them is done (net ocean fluxes, radiative fluxes, ...); the compute
   * stream is made to wait for the last of these copies, so synchronising it also means "results are on the host". */
  int32_t n_out_fields;
  int32_t pad_;
  NeHostField out_fields[NE_HOST_MAX_FIELDS];
} NeHostStepDesc;

/* ---- FieldTimeSeries window on the device (ne_series_ring.cu).
 * The reference keeps Nt_mem slices of each series in memory and, when the two interpolating time indices leave
 * that window, reloads the WHOLE window synchronously (update_state!(::PrescribedAtmosphere),
 * src/Atmospheres/prescribed_atmosphere.jl:154-162 -> Oceananigans update_field_time_series! -> set!(fts),
 * src/DataWrangling/JRA55/JRA55_field_time_series.jl:60-76 and :78-124).  Here the device array of a series is a
 * RING of n_slots slices: one slice at a time is replaced, on the ring's own copy stream, while the step's kernels
 * run; `NeTimeInterp.m1/m2` name ring slots.  Loading one slot does what set!(fts) does to one slice:
 *   raw file slice (nx x ny, x fastest, no halos) in pinned HOST memory
 *     -> read_data (src/DataWrangling/set_region_data.jl:162-163): file index = grid index + BoundingBoxOffset (di, dj),
 *        lat-axis mangling (mangle, :50-53: ShiftSouth reads j - 1, AverageNorthSouth averages j and j + 1; indices
 *        clamped to the file extent), missing value -> NaN, then convert_units (_set_region_kernel!, :200-205;
 *        src/DataWrangling/metadata_field.jl:486-525).  A Column region (region_kind = NE_REGION_COLUMN: a 1 x 1 series whose
 *        value is the NaN-aware blend of the four file cells around a point, blend(::Linear / ::Nearest), :168-193, with the
 *        bracketing indices and weights of region_info(::Column), :113-118, resolved by the host) goes through the same pass.
 *     -> interior of the slot, then fill_halo_regions!(fts): periodic in x (test/test_jra55.jl:40-47: fts[Nx+1,..] ==
 *        fts[1,..]) or mirrored when the source grid is bounded in x, mirrored (zero-flux) in y.
 * Which time index lives in which slot, and what to prefetch, is host policy (the binding's; series_window.py here). */
#define NE_RING_MAX_SERIES 16
enum { NE_CONV_NONE = 0, NE_CONV_NEGATE = 1, NE_CONV_ADD = 2, NE_CONV_SUB = 3, NE_CONV_MUL = 4, NE_CONV_DIV = 5,
       NE_CONV_MUL_DIV = 6 };   /* d, -d, d + a, d - a, d * a, d / a, d * a / b: each one rounding in the series eltype */
enum { NE_MANGLE_NONE = 0, NE_MANGLE_SHIFT_SOUTH = 1, NE_MANGLE_AVERAGE_NORTH_SOUTH = 2 };
enum { NE_REGION_BOX = 0, NE_REGION_COLUMN = 1 };        /* whole globe / BoundingBox (di, dj) | Column             */
enum { NE_COLUMN_LINEAR = 0, NE_COLUMN_NEAREST = 1 };    /* Column(...; interpolation = Linear() | Nearest())       */
typedef struct NeSeriesRingDesc {
  int32_t n_series;            /* series that share the time axis (<= NE_RING_MAX_SERIES)          */
  int32_t n_slots;             /* Nt_mem: slices each ring holds                                    */
  int32_t dtype;               /* NE_F32 / NE_F64: element type of the raw slices and of the rings  */
  int32_t periodic_x;          /* 1: x halos wrap (full-longitude source grid); 0: mirrored         */
  int64_t nx, ny, hx, hy;      /* source grid; a ring slice is (ny + 2 hy) rows of (nx + 2 hx)      */
  void* ring[NE_RING_MAX_SERIES];          /* device, n_slots slices each                           */
  int32_t conv_kind[NE_RING_MAX_SERIES];   /* NE_CONV_*                                             */
  double conv_a[NE_RING_MAX_SERIES], conv_b[NE_RING_MAX_SERIES];   /* rounded to `dtype` before use */
  int32_t has_missing[NE_RING_MAX_SERIES]; /* 1: raw values equal to missing_value become NaN       */
  double missing_value[NE_RING_MAX_SERIES];
  /* the raw (file) slice when it is not nx x ny: its extent (0 => nx / ny), where the grid's first cell sits in it
   * (region_info(::BoundingBox), set_region_data.jl:88-111) and the lat-axis mangling (mangling_for, :153-158:
   * raw_ny == ny - 1 => ShiftSouth, raw_ny == ny + 1 => AverageNorthSouth).                                      */
  int64_t raw_nx, raw_ny, di, dj;
  int32_t mangling[NE_RING_MAX_SERIES];    /* NE_MANGLE_*                                           */
  /* Column region (ColumnInfo, set_region_data.jl:79-87; nx = ny = 1): 0-based bracketing file indices (i_plus wraps to 0
   * across the periodic seam), blend weights in [0, 1] (rounded to `dtype` before use, as ColumnInfo holds them in the
   * target's element type) and the interpolation kind.                                                             */
  int32_t region_kind;                     /* NE_REGION_*                                           */
  int32_t column_interpolation;            /* NE_COLUMN_*                                           */
  int64_t col_i_minus, col_i_plus, col_j_minus, col_j_plus;
  double col_wx, col_wy;
} NeSeriesRingDesc;

/* ---- entry points ---------------------------------------------------------------------- */
int ne_version(void);
const char* ne_last_error(void);
int ne_device_count(void);
/* sizeof(struct <name>) as compiled into the library, -1 if unknown: lets a binding verify its mirror. */
int64_t ne_struct_size(const char* name);

int ne_frac_indices_f64(const NeFracIndexDesc*, void* stream);
int ne_frac_indices_f32(const NeFracIndexDesc*, void* stream);

int ne_interp_state_f64(const NeInterpDesc*, void* stream);   /* atmosphere (7) or radiation (2) */
int ne_interp_state_f32(const NeInterpDesc*, void* stream);

int ne_correct_atmosphere_elevation_f64(const NeElevationCorrectionDesc*, void* stream);
int ne_correct_atmosphere_elevation_f32(const NeElevationCorrectionDesc*, void* stream);

int ne_atmosphere_ocean_fluxes_f64(const NeAtmosOceanDesc*, void* stream);
int ne_atmosphere_ocean_fluxes_f32(const NeAtmosOceanDesc*, void* stream);

int ne_atmosphere_sea_ice_fluxes_f64(const NeAtmosSeaIceDesc*, void* stream);
int ne_atmosphere_sea_ice_fluxes_f32(const NeAtmosSeaIceDesc*, void* stream);

int ne_atmosphere_land_fluxes_f64(const NeAtmosLandDesc*, void* stream);
int ne_atmosphere_land_fluxes_f32(const NeAtmosLandDesc*, void* stream);

int ne_sea_ice_ocean_fluxes_f64(const NeSeaIceOceanDesc*, void* stream);
int ne_sea_ice_ocean_fluxes_f32(const NeSeaIceOceanDesc*, void* stream);
int ne_sea_ice_ocean_stress_f64(const NeSeaIceOceanStressDesc*, void* stream);
int ne_sea_ice_ocean_stress_f32(const NeSeaIceOceanStressDesc*, void* stream);

int ne_assemble_net_ocean_fluxes_f64(const NeAssembleOceanDesc*, void* stream);
int ne_assemble_net_ocean_fluxes_f32(const NeAssembleOceanDesc*, void* stream);
int ne_assemble_net_sea_ice_fluxes_f64(const NeAssembleSeaIceDesc*, void* stream);
int ne_assemble_net_sea_ice_fluxes_f32(const NeAssembleSeaIceDesc*, void* stream);

int ne_apply_radiative_fluxes_f64(const NeApplyRadiationDesc*, void* stream);
int ne_apply_radiative_fluxes_f32(const NeApplyRadiationDesc*, void* stream);

int ne_fused_interface_step_f64(const NeFusedStepDesc*, void* stream);
int ne_fused_interface_step_f32(const NeFusedStepDesc*, void* stream);
/* Phases 1-2 of the fused step only (both interpolations + the a–o solve; `assemble` and
 * `apply_radiation` are ignored): interpolate_state! + compute_atmosphere_ocean_fluxes!
 * (time_step_earth_system_model.jl:38-62) as one kernel when the configuration qualifies. */
int ne_interp_and_ao_fluxes_f64(const NeFusedStepDesc*, void* stream);
int ne_interp_and_ao_fluxes_f32(const NeFusedStepDesc*, void* stream);

int ne_diag_reduce_f64(const NeDiagDesc*, void* stream);
int ne_diag_reduce_f32(const NeDiagDesc*, void* stream);

/* The one collective of the path (north_star: "NCCL over NVLink only for the global flux and conservation
 * diagnostics reduction"; the reference sums the lazy global integrals of src/Diagnostics/interface_fluxes.jl:90-195 over its
 * MPI ranks).  In-place sum of `n` doubles (`sums`: device memory, normally NeDiagDesc.result) over the ranks of an NCCL
 * communicator, enqueued on `stream` behind the kernels that produce the sums.  `nccl_comm` is the host's own ncclComm_t
 * (NCCL.jl: `comm.handle`; one process per GPU): the library does not create communicators.  NCCL is bound at run time
 * (dlopen of libnccl.so.2, NE_B200_NCCL_LIB overrides the name), so single-GPU hosts need no NCCL; returns NE_E_NO_VARIANT
 * when the library cannot be loaded and NE_E_CUDA with the NCCL error string when the call fails. */
int ne_diag_allreduce_f64(void* nccl_comm, double* sums, int32_t n, void* stream);

/* Host-buffer convenience used by the end-to-end benchmark and by hosts without device
 * arrays: copies a contiguous host block to/from the device on `stream`. */
int ne_memcpy_h2d(void* dst_device, const void* src_host, uint64_t bytes, void* stream);
int ne_memcpy_d2h(void* dst_host, const void* src_device, uint64_t bytes, void* stream);
int ne_stream_synchronize(void* stream);

/* Host-buffer interface step: create once per device (owns a non-blocking copy stream and max_chunks + 1
 * events), then one call per coupled step on the caller's compute stream; everything is enqueued, nothing
 * synchronises.  Results equal the unchunked ne_fused_interface_step bit for bit. */
int ne_host_pipeline_create(void** handle, int32_t max_chunks);
int ne_host_pipeline_destroy(void* handle);
int ne_host_pipelined_step_f64(void* handle, const NeHostStepDesc*, void* stream);
int ne_host_pipelined_step_f32(void* handle, const NeHostStepDesc*, void* stream);

/* FieldTimeSeries ring (see NeSeriesRingDesc).  create: owns a non-blocking copy stream, one device staging slice per
 * series and n_slots + 1 events.  load: enqueue, on the copy stream, the H2D copy of one raw slice of every series
 * (host_raw[k], pinned, valid until the copy has run) and the kernel that writes slot `slot` (0-based) of every ring;
 * the copy waits for the kernels enqueued on the compute stream before the last release.  acquire: the compute stream
 * waits until `slot` has been written.  release: the kernels enqueued so far on the compute stream are the readers a
 * later load has to wait for.  Nothing synchronises the host. */
int ne_series_ring_create(void** handle, const NeSeriesRingDesc*);
int ne_series_ring_destroy(void* handle);
int ne_series_ring_load(void* handle, int32_t slot, const void* const* host_raw);
int ne_series_ring_acquire(void* handle, int32_t slot, void* compute_stream);
int ne_series_ring_release(void* handle, void* compute_stream);

/* FP64 DFMA-issue microbenchmark used to measure the FP64 roofline denominator
 * (MEASURED_PEAKS.json has no FP64 figure).  Returns achieved TFLOP/s in *tflops. */
int ne_measure_fp64_peak(double* tflops, double* sm_clock_mhz_estimate);

/* FP64 instructions ONE launch of ne_atmosphere_ocean_fluxes_f64 executes on this descriptor (default plugin tree):
 * runs the counting instantiation of the shipped solve kernel (same source, every FP64 operation goes through a
 * counting policy) and synchronises.  out[0..5] = thread-level {fused multiply-adds, multiplies, adds/subtracts,
 * estimated instructions of library code the policy cannot see (IEEE division / libdevice on rare paths),
 * thread trips (iterate_interface_state calls), warp trips (32-lane loop passes)}; out[6..7] reserved.
 * Executed flop = 2 out[0] + out[1] + out[2].  The roofline numerator of bench.py; no profiler involved. */
int ne_count_solve_ops_f64(const NeAtmosOceanDesc*, uint64_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NE_B200_H */
