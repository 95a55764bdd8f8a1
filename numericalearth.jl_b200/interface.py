"""Host-side mirror of the reference's interface-computation API on top of the C ABI.

Mirrors (citations relative to /root/reference/src/):
    PrescribedAtmosphere                 Atmospheres/prescribed_atmosphere.jl:200-260
    PrescribedRadiation                  Radiations/prescribed_radiation.jl:40-110
    StateExchanger / ComponentExchanger  EarthSystemModels/InterfaceComputations/state_exchanger.jl:14-64
    ComponentInterfaces                  EarthSystemModels/InterfaceComputations/component_interfaces.jl:203-496
    update_state! phases                 EarthSystemModels/time_step_earth_system_model.jl:38-83
    cpu_interpolating_time_indices       Oceananigans.OutputReaders (third party; restated)

The functions keep the reference's names (`interpolate_state!` -> interpolate_state, etc.).  Each one
fills a POD descriptor with device pointers and enqueues ONE C-ABI call on the current stream; no
synchronisation happens here.  Arrays are exchange-layout parents stored as (ny+2hy, nx+2hx)
row-major tensors == Oceananigans' column-major (nx+2hx, ny+2hy, 1) parents.
"""
import math
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from . import abi as A
from . import formulations as F
from .lib import get_library

NE_DT = {"f64": A.NE_F64, "f32": A.NE_F32}


# -------------------------------------------------------------------------------------------------
# grids
# -------------------------------------------------------------------------------------------------
@dataclass
class ExchangeGrid:
    """Regular latitude-longitude exchange grid, Periodic x Bounded, with Oceananigans-style halos.

    `lam`/`phi` are the cell-centre nodes INCLUDING halos ((nx+2hx,), (ny+2hy,)), in degrees."""
    nx: int
    ny: int
    hx: int = 7
    hy: int = 7
    longitude: tuple = (0.0, 360.0)
    latitude: tuple = (-75.0, 75.0)
    FT: str = "f64"
    nz: int = 1
    hz: int = 0
    j_offset: int = 0          # latitude-band shards: global row index of local row 1, minus 1
    ny_global: Optional[int] = None
    # rotated exchange grids (tripolar, cubed sphere): exchange-layout (cos θ, sin θ) of the angle between the grid's
    # x direction and geographic east — what Oceananigans' intrinsic_vector uses (third party; the binding fills it once).
    # None: latitude-longitude grid, interpolated vectors are stored as they are.
    rotation: Any = None

    def __post_init__(self):
        npd = np.float64 if self.FT == "f64" else np.float32
        nyg = self.ny_global or self.ny
        dl = (self.longitude[1] - self.longitude[0]) / self.nx
        dp = (self.latitude[1] - self.latitude[0]) / nyg
        i = np.arange(1 - self.hx, self.nx + self.hx + 1, dtype=np.float64)
        j = np.arange(1 - self.hy, self.ny + self.hy + 1, dtype=np.float64) + self.j_offset
        self.lam = (self.longitude[0] + (i - 0.5) * dl).astype(npd)
        self.phi = (self.latitude[0] + (j - 0.5) * dp).astype(npd)

    @property
    def shape(self):
        return (self.ny + 2 * self.hy, self.nx + 2 * self.hx)

    def pod(self, with_halo_ring=True) -> A.NeExchangeGrid:
        """interface_kernel_parameters (0:N+1) when with_halo_ring, `:xy` (1:N) otherwise
        (EarthSystemModels/InterfaceComputations/InterfaceComputations.jl:100-116)."""
        if with_halo_ring:
            return A.NeExchangeGrid(self.nx, self.ny, self.hx, self.hy, 0, self.nx + 1, 0, self.ny + 1)
        return A.NeExchangeGrid(self.nx, self.ny, self.hx, self.hy, 1, self.nx, 1, self.ny)

    def launch_points(self, with_halo_ring=True):
        return (self.nx + 2) * (self.ny + 2) if with_halo_ring else self.nx * self.ny

    def interior(self, a):
        """View of the (0:N+1) launch range of a parent array (numpy or torch)."""
        return a[self.hy - 1:self.hy + self.ny + 1, self.hx - 1:self.hx + self.nx + 1]


@dataclass
class LatLonSourceGrid:
    """The atmosphere/radiation LatitudeLongitudeGrid (JRA55: 640x320, halo 3; DataWrangling/JRA55/JRA55_metadata.jl:23-48)."""
    nx: int = 640
    ny: int = 320
    hx: int = 3
    hy: int = 3
    FT: str = "f32"
    lam_nodes: Any = None      # nx centre nodes
    phi_nodes: Any = None      # ny centre nodes
    x_regular: bool = True
    y_regular: bool = True

    def __post_init__(self):
        npd = np.float64 if self.FT == "f64" else np.float32
        if self.lam_nodes is None:
            d = 360.0 / self.nx
            self.lam_nodes = ((np.arange(self.nx) + 0.5) * d).astype(npd)
        if self.phi_nodes is None:
            d = 180.0 / self.ny
            self.phi_nodes = (-90.0 + (np.arange(self.ny) + 0.5) * d).astype(npd)
        self.lam_nodes = np.asarray(self.lam_nodes, dtype=npd)
        self.phi_nodes = np.asarray(self.phi_nodes, dtype=npd)

    @property
    def shape(self):
        return (self.ny + 2 * self.hy, self.nx + 2 * self.hx)


# -------------------------------------------------------------------------------------------------
# time interpolation (host) — Oceananigans cpu_interpolating_time_indices, restated
# -------------------------------------------------------------------------------------------------
def interpolating_time_indices(times, t, time_indexing="cyclical"):
    """Return (ñ, n1, n2) with 1-based n1, n2.  Cyclical period = (t_N - t_1) + Δt
    (DataWrangling/metadata_field_time_series.jl:51-56)."""
    times = np.asarray(times, dtype=np.float64)
    N = len(times)
    if N == 1:
        return 0.0, 1, 1
    t1, tN = times[0], times[-1]
    if time_indexing == "cyclical":
        dt_wrap = times[1] - times[0]
        period = (tN - t1) + dt_wrap
        tau = (t - t1) % period + t1
        if tau >= tN:
            return float((tau - tN) / dt_wrap), N, 1
        n1 = int(np.searchsorted(times, tau, side="right"))
        return float((tau - times[n1 - 1]) / (times[n1] - times[n1 - 1])), n1, n1 + 1
    if time_indexing == "clamp":
        if t <= t1:
            return 0.0, 1, 1
        if t >= tN:
            return 0.0, N, N
    n1 = int(np.clip(np.searchsorted(times, t, side="right"), 1, N - 1))
    return float((t - times[n1 - 1]) / (times[n1] - times[n1 - 1])), n1, n1 + 1


# -------------------------------------------------------------------------------------------------
# components
# -------------------------------------------------------------------------------------------------
@dataclass
class PrescribedAtmosphere:
    grid: LatLonSourceGrid
    times: Any
    u: Any = None
    v: Any = None
    T: Any = None
    q: Any = None
    p: Any = None
    rain: Any = ()            # tuple of series (freshwater_flux.rain ...)
    snow: Any = ()
    surface_layer_height: float = 10.0      # prescribed_atmosphere.jl:222
    boundary_layer_height: float = 512.0    # :223
    thermodynamics_parameters: Any = None
    time_indexing: str = "cyclical"
    window: Any = None        # a series_window.SeriesWindow when the series are device rings (partly in memory)

    def __post_init__(self):
        if self.thermodynamics_parameters is None:
            self.thermodynamics_parameters = F.AtmosphereThermodynamicsParameters(FT=self.grid.FT)


@dataclass
class PrescribedRadiation:
    grid: LatLonSourceGrid
    times: Any
    downwelling_shortwave: Any = None
    downwelling_longwave: Any = None
    stefan_boltzmann_constant: float = 5.670374419e-8   # Radiations/Radiations.jl:13-16
    surface_properties: dict = field(default_factory=lambda: {
        "ocean": F.SurfaceRadiationProperties(0.05, 0.97),
        "sea_ice": F.SurfaceRadiationProperties(0.7, 1.0)})
    time_indexing: str = "cyclical"
    window: Any = None


@dataclass
class PrescribedLand:
    """PrescribedLand freshwater fluxes (Lands/prescribed_land.jl; JRA55 river + iceberg runoff on its own 1440x720
    daily grid, DataWrangling/JRA55/JRA55_metadata.jl:24-25): a tuple of series on one source grid, summed."""
    grid: LatLonSourceGrid
    times: Any
    freshwater_flux: Any = ()      # tuple of series (rivers, icebergs, ...), each (nt, ny+2hy, nx+2hx)
    time_indexing: str = "cyclical"
    window: Any = None


@dataclass
class SlabLandState:
    """The land exchanger state the atmosphere-land flux kernel reads (atmosphere_land_fluxes.jl:78-84): bulk land
    temperature (Kelvin) and surface saturation, exchange-layout arrays or numbers."""
    T: Any = 288.0
    saturation: Any = 1.0
    # atmosphere_land_surface_properties(land_state) (atmosphere_land_fluxes.jl:107-115): per-cell fields a land model may
    # provide for LandRoughnessLength / LandZeroPlaneDisplacement (exchange-layout arrays; None = not provided, as SlabLand)
    momentum_roughness_length: Any = None
    scalar_roughness_length: Any = None
    zero_plane_displacement: Any = None


class _Fields:
    """Bag of named exchange-layout arrays."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def names(self):
        return list(self.__dict__.keys())


def _slot(backend, x) -> A.NeSlot:
    if x is None:
        return A.NeSlot(None, 0.0)
    if isinstance(x, (int, float)):
        return A.NeSlot(None, float(x))
    return A.NeSlot(backend.ptr(x), 0.0)


def _ptr(backend, x):
    return None if x is None else backend.ptr(x)


class ComponentInterfaces:
    """Exchange-grid state, flux fields and flux formulations of one coupled model
    (component_interfaces.jl:203-214, 390-496) + the StateExchanger (state_exchanger.jl:41-63)."""

    def __init__(self, grid: ExchangeGrid, backend, atmosphere: Optional[PrescribedAtmosphere] = None,
                 radiation: Optional[PrescribedRadiation] = None, *, sea_ice=False, lib=None,
                 atmosphere_ocean_fluxes=None, atmosphere_sea_ice_fluxes=None,
                 atmosphere_ocean_interface_temperature=None, atmosphere_ocean_velocity_difference=None,
                 atmosphere_ocean_interface_specific_humidity=None,
                 atmosphere_sea_ice_interface_temperature=None, atmosphere_sea_ice_velocity_difference=None,
                 sea_ice_ocean_heat_flux=None,
                 ocean_properties=None, sea_ice_properties=None,
                 gravitational_acceleration=9.80665, inactive=None, with_iterations=False, atmosphere_correction=None,
                 land: Optional[PrescribedLand] = None, slab_land: Optional[SlabLandState] = None,
                 atmosphere_land_fluxes=None, atmosphere_land_interface_specific_humidity=None,
                 atmosphere_land_velocity_difference=None, barotropic_potential=False, two_color_radiation=False):
        self.grid, self.backend = grid, backend
        self.land = land
        self.slab_land = slab_land
        self.al_flux_formulation = atmosphere_land_fluxes or F.default_atmosphere_land_fluxes()
        self.al_humidity = atmosphere_land_interface_specific_humidity or F.BulkHumidity()   # component_interfaces.jl:501-503
        self.al_velocity = atmosphere_land_velocity_difference or F.RelativeVelocity()
        self.lib = lib if lib is not None else get_library()
        if getattr(self.lib, "is_device", True) != backend.is_device:
            raise RuntimeError("array back-end and compute library disagree on where memory lives "
                               "(the CUDA library takes device arrays only; there is no CPU fallback)")
        self.atmosphere, self.radiation = atmosphere, radiation
        self.has_sea_ice = bool(sea_ice)
        FT = grid.FT
        Z = lambda: backend.zeros(grid.shape, FT)  # noqa: E731
        self.g = gravitational_acceleration
        self.ao_flux_formulation = atmosphere_ocean_fluxes or F.SimilarityTheoryFluxes()
        self.asi_flux_formulation = atmosphere_sea_ice_fluxes or F.atmosphere_sea_ice_similarity_theory()
        self.ao_properties = F.InterfaceProperties(
            atmosphere_ocean_interface_specific_humidity or F.ImpureSaturationSpecificHumidity(F.Liquid(), F._conv(FT, 0.98)),
            atmosphere_ocean_interface_temperature or F.BulkTemperature(),
            atmosphere_ocean_velocity_difference or F.RelativeVelocity())
        self.asi_properties = F.InterfaceProperties(
            F.ImpureSaturationSpecificHumidity(F.Ice(), None),   # component_interfaces.jl:280-281
            atmosphere_sea_ice_interface_temperature or F.SkinTemperature(F.ConductiveFlux(2.0)),
            atmosphere_sea_ice_velocity_difference or F.RelativeVelocity())
        self.sio_formulation = sea_ice_ocean_heat_flux or F.ThreeEquationHeatFlux()
        self.ocean_properties = ocean_properties or F.MediumProperties()
        self.sea_ice_properties = sea_ice_properties or F.MediumProperties(reference_density=900.0, heat_capacity=2100.0)
        self.inactive = inactive  # uint8 exchange-layout array or None
        self.sea_ice_latent_heat = 334e3  # J/kg (test/test_sea_ice_ocean_heat_fluxes.jl:54)

        # StateExchanger state (prescribed_atmosphere_regridder.jl:1-22)
        self.atmos_state = _Fields(u=Z(), v=Z(), T=Z(), p=Z(), q=Z(), Jrn=Z(), Jsn=Z())
        self.atmosphere_correction = self._materialize_correction(atmosphere_correction)
        self.frac = None
        if atmosphere is not None:
            self.frac = _Fields(i=backend.zeros(grid.shape, atmosphere.grid.FT), j=backend.zeros(grid.shape, atmosphere.grid.FT))
        self.rad_state = None
        self.rad_frac = None
        if radiation is not None:
            self.rad_state = _Fields(sw=Z(), lw=Z())
            # same source grid as the atmosphere (JRA55 radiation is): one set of fractional indices serves both, and the
            # fused step then interpolates the 9 series in one launch
            sg, ag = radiation.grid, (atmosphere.grid if atmosphere is not None else None)
            self.shared_frac = ag is not None and (sg is ag or (
                sg.FT == ag.FT and (sg.nx, sg.ny, sg.hx, sg.hy, sg.x_regular, sg.y_regular) ==
                (ag.nx, ag.ny, ag.hx, ag.hy, ag.x_regular, ag.y_regular) and
                np.array_equal(sg.lam_nodes, ag.lam_nodes) and np.array_equal(sg.phi_nodes, ag.phi_nodes)))
            self.rad_frac = self.frac if self.shared_frac else \
                _Fields(i=backend.zeros(grid.shape, radiation.grid.FT), j=backend.zeros(grid.shape, radiation.grid.FT))
            self.phi_dev = backend.from_numpy(grid.phi)
            self.rad_fluxes_ocean = _Fields(upwelling_longwave=Z(), downwelling_longwave=Z(), downwelling_shortwave=Z())
            self.rad_fluxes_sea_ice = _Fields(upwelling_longwave=Z(), downwelling_longwave=Z(), downwelling_shortwave=Z()) if sea_ice else None
        # PrescribedLand exchanger state (Lands/prescribed_land_regridder.jl): the summed runoff on the exchange grid
        self.land_state = None
        self.land_frac = None
        if land is not None:
            self.land_state = _Fields(freshwater_flux=Z())
            self.land_frac = _Fields(i=backend.zeros(grid.shape, land.grid.FT), j=backend.zeros(grid.shape, land.grid.FT))
        if slab_land is not None:   # AtmosphereSurfaceFluxes of the atmosphere-land interface (atmosphere_land_fluxes.jl:33-37)
            self.al_fluxes = _Fields(latent_heat=Z(), sensible_heat=Z(), water_vapor=Z(), x_momentum=Z(), y_momentum=Z(),
                                     friction_velocity=Z(), temperature_scale=Z(), water_vapor_scale=Z())
            self.al_temperature = Z()
            self.al_iterations = backend.zeros(grid.shape, "i32") if with_iterations else None
            self.land_surface_energy_flux = Z()     # land.fluxes.surface_energy_flux (positive upward), READ-MODIFY-WRITE
            self.rad_fluxes_land = _Fields(upwelling_longwave=Z(), downwelling_longwave=Z(), downwelling_shortwave=Z())
        # ocean surface state (pointers to the top-level plane of the 3-D parents)
        self.ocean_state = _Fields(u=Z(), v=Z(), T=Z(), S=Z())
        self.kappa = None
        # AtmosphereSurfaceFluxes (component_interfaces.jl:22-47)
        self.ao_fluxes = _Fields(latent_heat=Z(), sensible_heat=Z(), water_vapor=Z(), x_momentum=Z(), y_momentum=Z(),
                                 friction_velocity=Z(), temperature_scale=Z(), water_vapor_scale=Z())
        self.ao_temperature = Z()
        self.ao_iterations = backend.zeros(grid.shape, "i32") if with_iterations else None
        # net ocean fluxes (Oceans/assemble_net_ocean_fluxes.jl:118-126)
        self.net_ocean = _Fields(u=Z(), v=Z(), T=Z(), S=Z(), eta=Z(), freshwater_heat_content=Z())
        if sea_ice:
            self.sea_ice_state = _Fields(hi=Z(), hs=Z(), hc=Z(), concentration=Z(), S=Z(), top_temperature=Z(),
                                         ice_mass_flux=Z(), snow_mass_flux=Z(), u=Z(), v=Z())
            self.asi_fluxes = _Fields(latent_heat=Z(), sensible_heat=Z(), water_vapor=Z(), x_momentum=Z(), y_momentum=Z())
            self.asi_iterations = backend.zeros(grid.shape, "i32") if with_iterations else None
            self.sio_fluxes = _Fields(frazil_heat=Z(), interface_heat=Z(), salt=Z(), freshwater=Z(), x_momentum=Z(), y_momentum=Z())
            self.sio_temperature, self.sio_salinity = Z(), Z()
            self.net_sea_ice = _Fields(top_heat=Z(), top_snowfall=Z(), top_u=Z(), top_v=Z(), bottom_heat=Z())
        # barotropic forcing of the ocean free surface (interpolate_atmospheric_state.jl:80-85): potential .= p / rho_ocean
        self.barotropic_potential = Z() if barotropic_potential else None
        # TwoColorRadiation.surface_flux (Oceans/radiative_forcing.jl:84-91): the penetrating shortwave leaves JT
        self.two_color_surface_flux = Z() if two_color_radiation else None
        # FreezingLimitedOceanTemperature.frazil_heat (SeaIces/freezing_limited_ocean_temperature.jl:73-118): the
        # OceanOnlyModel default "sea ice"; allocated by the first step that passes an ocean column
        self.frazil_heat = None
        self._keep = []   # device copies of node arrays etc.

    # ---- one-time: fractional indices (prescribed_atmosphere_regridder.jl:41-71) ---------------------
    def _frac_desc(self, src: LatLonSourceGrid, frac):
        b, g = self.backend, self.grid
        lam, phi = b.from_numpy(g.lam), b.from_numpy(g.phi)
        ln, pn = b.from_numpy(src.lam_nodes), b.from_numpy(src.phi_nodes)
        self._keep += [lam, phi, ln, pn]
        d = A.NeFracIndexDesc()
        d.grid = g.pod(True)
        d.nodes_2d = 0
        d.lam, d.phi = b.ptr(lam), b.ptr(phi)
        d.src_dtype = NE_DT[src.FT]
        d.src_x_regular, d.src_y_regular = int(src.x_regular), int(src.y_regular)
        d.src_nx, d.src_ny = src.nx, src.ny
        d.src_lam_nodes, d.src_phi_nodes = b.ptr(ln), b.ptr(pn)
        d.frac_i, d.frac_j = b.ptr(frac.i), b.ptr(frac.j)
        return d

    def initialize(self):
        """initialize!(exchanger) — launch _compute_fractional_indices! for atmosphere and radiation."""
        s = self.backend.stream()
        if self.atmosphere is not None:
            self.lib.call("frac_indices", self.grid.FT, self._frac_desc(self.atmosphere.grid, self.frac), s)
        if self.radiation is not None and not self.shared_frac:
            self.lib.call("frac_indices", self.grid.FT, self._frac_desc(self.radiation.grid, self.rad_frac), s)
        if self.land is not None:
            # the reference evaluates these fractional indices on the fly from the node (interpolate_land_state.jl:55-60);
            # they are the same function of the same node, so they are computed once
            self.lib.call("frac_indices", self.grid.FT, self._frac_desc(self.land.grid, self.land_frac), s)

    # ---- phase 1: interpolation ---------------------------------------------------------------------
    def _time_interp(self, src, t, frac_dtype="f64"):
        """Time weight and the two in-memory slices of a prescribed component at time t.  Fully in-memory series: the
        slices are the time indices; a windowed component (`src.window`) answers with ring slots and makes the
        compute stream wait for them (update_field_time_series!, prescribed_atmosphere.jl:154-162)."""
        if src.window is not None:
            nt, m1, m2, same = src.window.time_interp(t, self.backend.stream())
            return A.NeTimeInterp(frac=nt, frac_dtype=NE_DT[frac_dtype], m1=m1, m2=m2, same=same)
        nt, n1, n2 = interpolating_time_indices(src.times, t, src.time_indexing)
        return A.NeTimeInterp(frac=nt, frac_dtype=NE_DT[frac_dtype], m1=n1, m2=n2, same=int(n1 == n2))

    @staticmethod
    def _n_memory(src):
        return src.window.n_slots if src.window is not None else len(src.times)

    def release_windows(self):
        """After the interpolation kernels of a step have been enqueued: let every windowed component start loading
        the slices of the coming intervals behind them."""
        seen = []
        for src in (self.atmosphere, self.radiation, self.land):
            w = getattr(src, "window", None) if src is not None else None
            if w is not None and not any(w is x for x in seen):
                seen.append(w)
                w.after_launch(self.backend.stream())

    def atmosphere_interp_desc(self, t) -> A.NeInterpDesc:
        b, g, atm = self.backend, self.grid, self.atmosphere
        d = A.NeInterpDesc()
        d.grid = g.pod(True)
        d.frac_i, d.frac_j = b.ptr(self.frac.i), b.ptr(self.frac.j)
        d.src_dtype = NE_DT[atm.grid.FT]
        d.src_nx, d.src_ny, d.src_hx, d.src_hy = atm.grid.nx, atm.grid.ny, atm.grid.hx, atm.grid.hy
        d.src_nt = self._n_memory(atm)
        d.time = self._time_interp(atm, t)
        fields = [(atm.u,), (atm.v,), (atm.T,), (atm.q,), (atm.p,), tuple(atm.rain), tuple(atm.snow)]
        outs = [self.atmos_state.u, self.atmos_state.v, self.atmos_state.T, self.atmos_state.q, self.atmos_state.p,
                self.atmos_state.Jrn, self.atmos_state.Jsn]
        d.n_fields = 7
        for f, (series, out) in enumerate(zip(fields, outs)):
            series = [s for s in series]
            if len(series) > A.NE_MAX_SUMMANDS:
                raise F.NoKernelVariantError("more than 4 summands in a precipitation tuple")
            d.n_summands[f] = len(series)
            for k, s in enumerate(series):
                d.series[f][k].data = _ptr(b, s)
            d.out[f] = b.ptr(out)
        if self.barotropic_potential is not None:   # :80-85, from the interpolated pressure (field 4)
            d.potential, d.potential_from = b.ptr(self.barotropic_potential), 4
            d.ocean_reference_density = float(self.ocean_properties.reference_density)
        if g.rotation is not None:   # intrinsic_vector (interpolate_atmospheric_state.jl:123-126)
            if not hasattr(self, "_rotation_dev"):
                npd = np.float64 if g.FT == "f64" else np.float32
                self._rotation_dev = tuple(b.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=npd))) for a in g.rotation)
                for a in self._rotation_dev:
                    if tuple(a.shape) != tuple(g.shape):
                        raise ValueError("rotation arrays must have the exchange layout (ny + 2hy, nx + 2hx)")
            d.rotation_cos, d.rotation_sin = b.ptr(self._rotation_dev[0]), b.ptr(self._rotation_dev[1])
            d.rotate_u, d.rotate_v = 0, 1
        return d

    def radiation_interp_desc(self, t) -> A.NeInterpDesc:
        b, g, rad = self.backend, self.grid, self.radiation
        d = A.NeInterpDesc()
        d.grid = g.pod(True)
        d.frac_i, d.frac_j = b.ptr(self.rad_frac.i), b.ptr(self.rad_frac.j)
        d.src_dtype = NE_DT[rad.grid.FT]
        d.src_nx, d.src_ny, d.src_hx, d.src_hy = rad.grid.nx, rad.grid.ny, rad.grid.hx, rad.grid.hy
        d.src_nt = self._n_memory(rad)
        d.time = self._time_interp(rad, t)
        d.n_fields = 2
        for f, (s, out) in enumerate([(rad.downwelling_shortwave, self.rad_state.sw), (rad.downwelling_longwave, self.rad_state.lw)]):
            d.n_summands[f] = 1
            d.series[f][0].data = b.ptr(s)
            d.out[f] = b.ptr(out)
        return d

    def land_interp_desc(self, t) -> A.NeInterpDesc:
        """interpolate_state!(exchanger, grid, ::PrescribedLand, model) (Lands/interpolate_land_state.jl:6-61): ONE output
        field, the sum over the runoff series (interp_atmos_time_series of a tuple, interpolate_atmospheric_state.jl:152-182)."""
        b, g, land = self.backend, self.grid, self.land
        series = list(land.freshwater_flux)
        if not 1 <= len(series) <= A.NE_MAX_SUMMANDS:
            raise F.NoKernelVariantError("PrescribedLand needs 1 to 4 freshwater flux series")
        d = A.NeInterpDesc()
        d.grid = g.pod(True)
        d.frac_i, d.frac_j = b.ptr(self.land_frac.i), b.ptr(self.land_frac.j)
        d.src_dtype = NE_DT[land.grid.FT]
        d.src_nx, d.src_ny, d.src_hx, d.src_hy = land.grid.nx, land.grid.ny, land.grid.hx, land.grid.hy
        d.src_nt = self._n_memory(land)
        d.time = self._time_interp(land, t)
        d.n_fields = 1
        d.n_summands[0] = len(series)
        for k, sr in enumerate(series):
            d.series[0][k].data = b.ptr(sr)
        d.out[0] = b.ptr(self.land_state.freshwater_flux)
        return d

    # ---- ElevationCorrection (atmosphere_state_correction.jl) -----------------------------------------------
    def _materialize_correction(self, c):
        """materialize_correction (:89-110): Δz = zˢ − zᵃ on the exchange grid; g, Rᵈ from the atmosphere's
        thermodynamics (Rᵈ = R / Mᵈ in the thermodynamics element type, thermodynamic_parameters.jl:73)."""
        if c is None:
            return None
        if not isinstance(c, F.ElevationCorrection):
            raise F.NoKernelVariantError(f"atmosphere-state correction {type(c).__name__} has no kernel variant")
        g = self.grid
        ft = np.float64 if g.FT == "f64" else np.float32

        def field(x):
            out = np.zeros(g.shape, dtype=ft)
            if callable(x):
                raise F.NoKernelVariantError("function-valued elevations must be evaluated by the caller (no closures on the device)")
            x = np.asarray(x, dtype=ft)
            if x.shape == tuple(g.shape):
                out[...] = x
            else:   # number or interior array: set!(field, ·) fills the interior only
                out[g.hy:g.hy + g.ny, g.hx:g.hx + g.nx] = x
            return out
        dz = np.zeros(g.shape, dtype=ft)
        zs, za = field(c.surface_elevation), field(c.atmosphere_elevation)
        dz[g.hy:g.hy + g.ny, g.hx:g.hx + g.nx] = (zs - za)[g.hy:g.hy + g.ny, g.hx:g.hx + g.nx]
        th = self.atmosphere.thermodynamics_parameters if self.atmosphere is not None else F.AtmosphereThermodynamicsParameters(FT=g.FT)
        ct = np.float64 if th.FT == "f64" else np.float32
        R_d = ct(th.gas_constant) / ct(th.dry_air_molar_mass)
        return _Fields(dz=self.backend.from_numpy(dz), lapse_rate=float(c.lapse_rate),
                       gravitational_acceleration=float(ft(9.80665)),       # default_gravitational_acceleration (:72)
                       dry_air_gas_constant=float(ft(R_d)))

    def elevation_correction_desc(self) -> A.NeElevationCorrectionDesc:
        c, b = self.atmosphere_correction, self.backend
        d = A.NeElevationCorrectionDesc()
        d.grid = self.grid.pod(True)
        d.T, d.p, d.elevation_difference = b.ptr(self.atmos_state.T), b.ptr(self.atmos_state.p), b.ptr(c.dz)
        d.lapse_rate, d.gravitational_acceleration, d.dry_air_gas_constant = c.lapse_rate, c.gravitational_acceleration, c.dry_air_gas_constant
        return d

    def correct_state(self):
        """Phase 1.5: correct_state!(exchanger.atmosphere, grid) (time_step_earth_system_model.jl:56-62)."""
        if self.atmosphere_correction is not None:
            self.lib.call("correct_atmosphere_elevation", self.grid.FT, self.elevation_correction_desc(), self.backend.stream())

    def interpolate_state(self, t):
        """interpolate_state!(exchanger.radiation, ...) then (exchanger.atmosphere, ...)
        (time_step_earth_system_model.jl:50-51)."""
        s = self.backend.stream()
        if self.radiation is not None:
            self.lib.call("interp_state", self.grid.FT, self.radiation_interp_desc(t), s)
        if self.atmosphere is not None:
            self.lib.call("interp_state", self.grid.FT, self.atmosphere_interp_desc(t), s)
        if self.land is not None:
            self.lib.call("interp_state", self.grid.FT, self.land_interp_desc(t), s)
        self.release_windows()

    # ---- radiation POD ---------------------------------------------------------------------------------
    def _surface_radiation(self, surface, time_seconds=None) -> A.NeSurfaceRadiation:
        r = A.NeSurfaceRadiation()
        if time_seconds is None:
            time_seconds = getattr(self, "clock_time", 0.0)
        rad = self.radiation
        if rad is None or surface not in rad.surface_properties:
            r.enabled = 0
            return r
        b = self.backend
        sp = rad.surface_properties[surface]
        r.enabled = 1
        r.stefan_boltzmann_constant = rad.stefan_boltzmann_constant
        if isinstance(sp.albedo, (int, float)):
            r.albedo_kind, r.albedo = A.NE_ALBEDO_CONSTANT, float(sp.albedo)
        elif isinstance(sp.albedo, F.LatitudeDependentAlbedo):
            r.albedo_kind = A.NE_ALBEDO_LATITUDE_DEPENDENT
            r.albedo, r.albedo_direct = sp.albedo.diffuse, sp.albedo.direct
            r.latitude = b.ptr(self.phi_dev)
        elif isinstance(sp.albedo, F.SeaIceAlbedo):
            al, sa = sp.albedo, r.sea_ice_albedo
            r.albedo_kind = A.NE_ALBEDO_SEA_ICE
            for n in ("ice_albedo", "snow_albedo", "ice_melt_reduction", "snow_melt_reduction", "melting_temperature",
                      "temperature_range", "ocean_albedo", "minimum_ice_thickness", "minimum_snow_depth"):
                setattr(sa, n, float(getattr(al, n)))
            sa.ice_thickness, sa.snow_thickness = b.ptr(al.ice_thickness), _ptr(b, al.snow_thickness)
            sa.surface_temperature = b.ptr(al.surface_temperature)
        elif isinstance(sp.albedo, F.TabulatedAlbedo):
            al, ta = sp.albedo, r.tabulated_albedo
            r.albedo_kind = A.NE_ALBEDO_TABULATED
            ta.table = b.ptr(al.table)
            ta.n_phi, ta.n_t = int(al.table.shape[0]), int(al.table.shape[1])
            ft = np.float64 if self.grid.FT == "f64" else np.float32
            ta.t_values[0], ta.t_values[1] = float(ft(al.t_values[0])), float(ft(al.t_values[1]))
            ta.phi_values[0], ta.phi_values[1] = float(ft(al.phi_values[0])), float(ft(al.phi_values[1]))
            ta.solar_constant, ta.day_to_radians = float(al.solar_constant), float(ft(al.day_to_radians))
            ta.noon_in_seconds = float(al.noon_in_seconds)
            _, sec, delta = F.TabulatedAlbedo.clock_scalars(float(time_seconds))
            ta.seconds_in_day, ta.declination = sec, float(ft(delta))
            if not hasattr(self, "lam_dev"):
                self.lam_dev = b.from_numpy(self.grid.lam)
            ta.longitude = b.ptr(self.lam_dev)
            r.latitude = b.ptr(self.phi_dev)
        elif hasattr(sp.albedo, "shape"):
            r.albedo_kind = A.NE_ALBEDO_FIELD
            r.albedo_field = b.ptr(sp.albedo)
        else:
            raise F.NoKernelVariantError(f"albedo {sp.albedo!r} has no kernel variant")
        if not isinstance(sp.emissivity, (int, float)):
            raise F.NoKernelVariantError("field-valued emissivity has no kernel variant")
        r.emissivity = float(sp.emissivity)
        r.downwelling_shortwave = b.ptr(self.rad_state.sw)
        r.downwelling_longwave = b.ptr(self.rad_state.lw)
        return r

    # ---- phase 2: turbulent fluxes ---------------------------------------------------------------------
    def atmosphere_ocean_desc(self) -> A.NeAtmosOceanDesc:
        b, g = self.backend, self.grid
        d = A.NeAtmosOceanDesc()
        d.grid = g.pod(True)
        a = self.atmos_state
        d.ua, d.va, d.Ta, d.pa, d.qa = b.ptr(a.u), b.ptr(a.v), b.ptr(a.T), b.ptr(a.p), b.ptr(a.q)
        atm = self.atmosphere
        d.surface_layer_height = _slot(b, atm.surface_layer_height if atm else 10.0)
        d.boundary_layer_height = _slot(b, atm.boundary_layer_height if atm else 512.0)
        o = self.ocean_state
        d.uo, d.vo, d.To, d.So = _slot(b, o.u), _slot(b, o.v), _slot(b, o.T), _slot(b, o.S)
        d.kappa = _ptr(b, self.kappa)
        d.inactive = _ptr(b, self.inactive)
        d.radiation = self._surface_radiation("ocean")
        d.thermo = (atm.thermodynamics_parameters if atm else F.AtmosphereThermodynamicsParameters(FT=g.FT)).pod()
        d.gravitational_acceleration = self.g
        d.flux = F.flux_formulation_pod(self.ao_flux_formulation, FT=g.FT)
        d.properties = self.ao_properties.pod()
        d.ocean = self.ocean_properties.pod()
        f = self.ao_fluxes
        d.latent_heat, d.sensible_heat, d.water_vapor = b.ptr(f.latent_heat), b.ptr(f.sensible_heat), b.ptr(f.water_vapor)
        d.x_momentum, d.y_momentum = b.ptr(f.x_momentum), b.ptr(f.y_momentum)
        d.interface_temperature = b.ptr(self.ao_temperature)
        d.friction_velocity, d.temperature_scale, d.water_vapor_scale = \
            b.ptr(f.friction_velocity), b.ptr(f.temperature_scale), b.ptr(f.water_vapor_scale)
        d.iterations = _ptr(b, self.ao_iterations)
        return d

    def compute_atmosphere_ocean_fluxes(self):
        self.lib.call("atmosphere_ocean_fluxes", self.grid.FT, self.atmosphere_ocean_desc(), self.backend.stream())

    def atmosphere_sea_ice_desc(self) -> A.NeAtmosSeaIceDesc:
        b, g = self.backend, self.grid
        d = A.NeAtmosSeaIceDesc()
        d.grid = g.pod(True)
        a = self.atmos_state
        d.ua, d.va, d.Ta, d.pa, d.qa = b.ptr(a.u), b.ptr(a.v), b.ptr(a.T), b.ptr(a.p), b.ptr(a.q)
        atm = self.atmosphere
        d.surface_layer_height = _slot(b, atm.surface_layer_height if atm else 10.0)
        d.boundary_layer_height = _slot(b, atm.boundary_layer_height if atm else 512.0)
        o, si = self.ocean_state, self.sea_ice_state
        d.To, d.So = _slot(b, o.T), _slot(b, o.S)
        d.hi, d.hs, d.hc, d.concentration = _slot(b, si.hi), _slot(b, si.hs), _slot(b, si.hc), _slot(b, si.concentration)
        d.inactive = _ptr(b, self.inactive)
        d.radiation = self._surface_radiation("sea_ice")
        d.thermo = (atm.thermodynamics_parameters if atm else F.AtmosphereThermodynamicsParameters(FT=g.FT)).pod()
        d.gravitational_acceleration = self.g
        d.flux = F.flux_formulation_pod(self.asi_flux_formulation, FT=g.FT)
        d.properties = self.asi_properties.pod()
        d.ocean = self.ocean_properties.pod()
        d.sea_ice = self.sea_ice_properties.pod()
        f = self.asi_fluxes
        d.latent_heat, d.sensible_heat, d.water_vapor = b.ptr(f.latent_heat), b.ptr(f.sensible_heat), b.ptr(f.water_vapor)
        d.x_momentum, d.y_momentum = b.ptr(f.x_momentum), b.ptr(f.y_momentum)
        d.interface_temperature = b.ptr(si.top_temperature)
        d.iterations = _ptr(b, self.asi_iterations)
        return d

    def compute_atmosphere_sea_ice_fluxes(self):
        if self.has_sea_ice:
            self.lib.call("atmosphere_sea_ice_fluxes", self.grid.FT, self.atmosphere_sea_ice_desc(), self.backend.stream())

    def atmosphere_land_desc(self) -> A.NeAtmosLandDesc:
        """compute_atmosphere_land_fluxes!(model) (atmosphere_land_fluxes.jl:48-122): launch range `:xy`."""
        b, g, atm = self.backend, self.grid, self.atmosphere
        d = A.NeAtmosLandDesc()
        d.grid = g.pod(False)
        a = self.atmos_state
        d.ua, d.va, d.Ta, d.pa, d.qa = b.ptr(a.u), b.ptr(a.v), b.ptr(a.T), b.ptr(a.p), b.ptr(a.q)
        d.surface_layer_height = _slot(b, atm.surface_layer_height if atm else 10.0)
        d.boundary_layer_height = _slot(b, atm.boundary_layer_height if atm else 512.0)
        d.land_temperature, d.saturation = _slot(b, self.slab_land.T), _slot(b, self.slab_land.saturation)
        d.thermo = (atm.thermodynamics_parameters if atm else F.AtmosphereThermodynamicsParameters(FT=g.FT)).pod()
        d.gravitational_acceleration = self.g
        d.flux = F.flux_formulation_pod(self.al_flux_formulation, land=True)
        sl = self.slab_land
        d.momentum_roughness_length, d.scalar_roughness_length, d.zero_plane_displacement = \
            _ptr(b, sl.momentum_roughness_length), _ptr(b, sl.scalar_roughness_length), _ptr(b, sl.zero_plane_displacement)
        d.properties = F.InterfaceProperties(F.ImpureSaturationSpecificHumidity(F.Liquid(), None), F.BulkTemperature(),
                                             self.al_velocity).pod()
        d.humidity = F.land_humidity_pod(self.al_humidity)
        f = self.al_fluxes
        d.latent_heat, d.sensible_heat, d.water_vapor = b.ptr(f.latent_heat), b.ptr(f.sensible_heat), b.ptr(f.water_vapor)
        d.x_momentum, d.y_momentum = b.ptr(f.x_momentum), b.ptr(f.y_momentum)
        d.interface_temperature = b.ptr(self.al_temperature)
        d.friction_velocity, d.temperature_scale, d.water_vapor_scale = \
            b.ptr(f.friction_velocity), b.ptr(f.temperature_scale), b.ptr(f.water_vapor_scale)
        d.iterations = _ptr(b, self.al_iterations)
        return d

    def compute_atmosphere_land_fluxes(self):
        if self.slab_land is not None:
            self.lib.call("atmosphere_land_fluxes", self.grid.FT, self.atmosphere_land_desc(), self.backend.stream())

    def sea_ice_ocean_desc(self, T3, S3, dz, dt, nz, hz=0) -> A.NeSeaIceOceanDesc:
        b, g = self.backend, self.grid
        d = A.NeSeaIceOceanDesc()
        d.grid = g.pod(False)
        d.nz, d.hz = nz, hz
        d.T, d.S, d.dz, d.dt = b.ptr(T3), b.ptr(S3), b.ptr(dz), dt
        ff = self.sio_formulation if self.has_sea_ice else None
        if ff is None:   # FreezingLimitedOceanTemperature: the frazil clamp only (freezing_limited_ocean_temperature.jl:73-118)
            d.formulation = A.NE_SIO_FREEZE_ONLY
            if not self.has_sea_ice:
                if self.frazil_heat is None:
                    self.frazil_heat = b.zeros(g.shape, g.FT)
                d.frazil_heat = b.ptr(self.frazil_heat)
        else:
            if isinstance(ff, F.IceBathHeatFlux):
                d.formulation = A.NE_SIO_ICE_BATH
                d.heat_transfer_coefficient = ff.heat_transfer_coefficient
            elif isinstance(ff, F.ThreeEquationHeatFlux):
                d.formulation = A.NE_SIO_THREE_EQUATION
                d.heat_transfer_coefficient = ff.heat_transfer_coefficient
                d.salt_transfer_coefficient = ff.salt_transfer_coefficient
                if ff.conductive_flux is not None:
                    d.has_conductive_flux = 1
                    d.conductivity = ff.conductive_flux.conductivity
                    d.internal_temperature = b.ptr(ff.internal_temperature)
            else:
                raise F.NoKernelVariantError(f"sea-ice-ocean heat flux {ff!r} has no kernel variant")
            if isinstance(ff.friction_velocity, F.MomentumBasedFrictionVelocity):
                d.friction_velocity_kind = A.NE_USTAR_MOMENTUM_BASED
            elif isinstance(ff.friction_velocity, (int, float)):
                d.friction_velocity_kind, d.friction_velocity = A.NE_USTAR_CONSTANT, float(ff.friction_velocity)
            else:
                raise F.NoKernelVariantError(f"friction velocity {ff.friction_velocity!r} has no kernel variant")
        d.latent_heat = self.sea_ice_latent_heat   # sea_ice.model.phase_transitions.reference_latent_heat
        d.ocean = self.ocean_properties.pod()
        f = self.sio_fluxes if self.has_sea_ice else None
        if self.has_sea_ice:
            si = self.sea_ice_state
            d.hi, d.hc, d.concentration, d.ice_salinity = _slot(b, si.hi), _slot(b, si.hc), _slot(b, si.concentration), _slot(b, si.S)
            d.ice_mass_flux, d.snow_mass_flux = _slot(b, si.ice_mass_flux), _slot(b, si.snow_mass_flux)
            d.x_momentum_in, d.y_momentum_in = b.ptr(f.x_momentum), b.ptr(f.y_momentum)
            d.frazil_heat, d.interface_heat, d.salt, d.freshwater = \
                b.ptr(f.frazil_heat), b.ptr(f.interface_heat), b.ptr(f.salt), b.ptr(f.freshwater)
            d.interface_temperature, d.interface_salinity = b.ptr(self.sio_temperature), b.ptr(self.sio_salinity)
        return d

    def compute_sea_ice_ocean_fluxes(self, ocean_column):
        """compute_sea_ice_ocean_fluxes!(model) (sea_ice_ocean_fluxes.jl:20-77) with a sea-ice model; without one the
        OceanOnlyModel default FreezingLimitedOceanTemperature (freezing_limited_ocean_temperature.jl:73-93): clamp the
        T column to the liquidus and store the frazil heat.  ocean_column = (T3, S3, dz, dt, nz, hz); None: no 3-D ocean
        state was handed over (surface-only callers), nothing to do."""
        if ocean_column is not None:
            self.lib.call("sea_ice_ocean_fluxes", self.grid.FT, self.sea_ice_ocean_desc(*ocean_column), self.backend.stream())

    # ---- phase 3: net fluxes ----------------------------------------------------------------------------
    def assemble_ocean_desc(self) -> A.NeAssembleOceanDesc:
        b, g = self.backend, self.grid
        d = A.NeAssembleOceanDesc()
        d.grid = g.pod(False)
        f = self.ao_fluxes
        d.sensible_heat, d.latent_heat, d.water_vapor = _slot(b, f.sensible_heat), _slot(b, f.latent_heat), _slot(b, f.water_vapor)
        d.x_momentum_ao, d.y_momentum_ao = _slot(b, f.x_momentum), _slot(b, f.y_momentum)
        if self.has_sea_ice:
            s, si = self.sio_fluxes, self.sea_ice_state
            d.interface_heat, d.salt_io, d.freshwater_io = _slot(b, s.interface_heat), _slot(b, s.salt), _slot(b, s.freshwater)
            d.x_momentum_io, d.y_momentum_io = _slot(b, s.x_momentum), _slot(b, s.y_momentum)
            d.concentration = _slot(b, si.concentration)
        else:  # ZeroFluxes / ZeroField (component_interfaces.jl:181-201)
            for n in ("interface_heat", "salt_io", "freshwater_io", "x_momentum_io", "y_momentum_io", "concentration"):
                setattr(d, n, _slot(b, 0.0))
        d.ocean_surface_temperature = _slot(b, self.ocean_state.T)
        d.rainfall, d.snowfall = _slot(b, self.atmos_state.Jrn), _slot(b, self.atmos_state.Jsn)
        d.intercepted_snowfall = _slot(b, 0.0)
        d.land_freshwater = _slot(b, self.land_state.freshwater_flux if self.land is not None else 0.0)
        d.inactive = _ptr(b, self.inactive)
        d.ocean = self.ocean_properties.pod()
        n = self.net_ocean
        d.tau_x, d.tau_y, d.JT, d.JS, d.Jw, d.JH = b.ptr(n.u), b.ptr(n.v), b.ptr(n.T), b.ptr(n.S), b.ptr(n.eta), b.ptr(n.freshwater_heat_content)
        return d

    def assemble_sea_ice_desc(self) -> A.NeAssembleSeaIceDesc:
        b, g = self.backend, self.grid
        d = A.NeAssembleSeaIceDesc()
        d.grid = g.pod(False)
        f, s = self.asi_fluxes, self.sio_fluxes
        d.sensible_heat, d.latent_heat = _slot(b, f.sensible_heat), _slot(b, f.latent_heat)
        d.x_momentum, d.y_momentum = _slot(b, f.x_momentum), _slot(b, f.y_momentum)
        d.frazil_heat, d.interface_heat = _slot(b, s.frazil_heat), _slot(b, s.interface_heat)
        d.snowfall = _slot(b, self.atmos_state.Jsn)
        d.concentration = _slot(b, self.sea_ice_state.concentration)
        d.inactive = _ptr(b, self.inactive)
        n = self.net_sea_ice
        d.top_heat, d.top_snowfall, d.top_u, d.top_v, d.bottom_heat = \
            b.ptr(n.top_heat), b.ptr(n.top_snowfall), b.ptr(n.top_u), b.ptr(n.top_v), b.ptr(n.bottom_heat)
        return d

    def update_net_fluxes(self):
        s = self.backend.stream()
        if self.has_sea_ice:
            self.lib.call("assemble_net_sea_ice_fluxes", self.grid.FT, self.assemble_sea_ice_desc(), s)
        self.lib.call("assemble_net_ocean_fluxes", self.grid.FT, self.assemble_ocean_desc(), s)

    # ---- phase 4: radiation ---------------------------------------------------------------------------------
    def apply_radiation_desc(self, over_sea_ice=False) -> A.NeApplyRadiationDesc:
        b, g = self.backend, self.grid
        d = A.NeApplyRadiationDesc()
        d.grid = g.pod(False)
        d.radiation = self._surface_radiation("sea_ice" if over_sea_ice else "ocean")
        d.concentration = _slot(b, self.sea_ice_state.concentration if self.has_sea_ice else 0.0)
        d.inactive = _ptr(b, self.inactive)
        d.over_sea_ice = int(over_sea_ice)
        if over_sea_ice:
            d.surface_temperature = b.ptr(self.sea_ice_state.top_temperature)
            d.medium = self.sea_ice_properties.pod()
            d.heat_flux = b.ptr(self.net_sea_ice.top_heat)
            r = self.rad_fluxes_sea_ice
        else:
            d.surface_temperature = b.ptr(self.ao_temperature)
            d.medium = self.ocean_properties.pod()
            d.heat_flux = b.ptr(self.net_ocean.T)
            r = self.rad_fluxes_ocean
            if self.two_color_surface_flux is not None:
                d.two_color, d.two_color_surface_flux = 1, b.ptr(self.two_color_surface_flux)
        d.upwelling_longwave, d.downwelling_longwave, d.downwelling_shortwave = \
            b.ptr(r.upwelling_longwave), b.ptr(r.downwelling_longwave), b.ptr(r.downwelling_shortwave)
        return d

    def apply_air_land_radiative_fluxes(self):
        """apply_air_land_radiative_fluxes!(model) (Radiations/apply_air_land_radiative_fluxes.jl:17-97): needs a `land`
        entry in the radiation's surface properties."""
        if self.radiation is None or self.slab_land is None or "land" not in self.radiation.surface_properties:
            return
        b, g = self.backend, self.grid
        d = A.NeApplyRadiationDesc()
        d.grid = g.pod(False)
        d.radiation = self._surface_radiation("land")
        d.concentration = _slot(b, 0.0)
        d.inactive = _ptr(b, self.inactive)
        d.over_sea_ice = 2
        d.surface_temperature = b.ptr(self.al_temperature)
        d.medium = F.MediumProperties(temperature_units=F.DegreesKelvin()).pod()
        d.heat_flux = b.ptr(self.land_surface_energy_flux)
        r = self.rad_fluxes_land
        d.upwelling_longwave, d.downwelling_longwave, d.downwelling_shortwave = \
            b.ptr(r.upwelling_longwave), b.ptr(r.downwelling_longwave), b.ptr(r.downwelling_shortwave)
        self.lib.call("apply_radiative_fluxes", g.FT, d, b.stream())

    def apply_air_sea_radiative_fluxes(self):
        if self.radiation is not None:
            self.lib.call("apply_radiative_fluxes", self.grid.FT, self.apply_radiation_desc(False), self.backend.stream())

    def apply_air_sea_ice_radiative_fluxes(self):
        if self.radiation is not None and self.has_sea_ice:
            self.lib.call("apply_radiative_fluxes", self.grid.FT, self.apply_radiation_desc(True), self.backend.stream())

    # ---- the whole interface step --------------------------------------------------------------------------
    def update_state(self, t, ocean_column=None):
        """update_state!(model) phases 1-4 (time_step_earth_system_model.jl:38-83).
        ocean_column = (T3, S3, dz, dt, nz, hz) enables the sea-ice–ocean kernel."""
        self.clock_time = t   # clock-dependent surface properties (TabulatedAlbedo) read it
        self.interpolate_state(t)
        self.correct_state()
        self.compute_atmosphere_ocean_fluxes()
        self.compute_atmosphere_sea_ice_fluxes()
        self.compute_atmosphere_land_fluxes()
        self.compute_sea_ice_ocean_fluxes(ocean_column)
        self.update_net_fluxes()
        self.apply_air_sea_radiative_fluxes()
        self.apply_air_sea_ice_radiative_fluxes()
        self.apply_air_land_radiative_fluxes()

    def fused_step_desc(self, t, diagnostics=None) -> A.NeFusedStepDesc:
        """`diagnostics` (a sharding.FluxDiagnostics): its area-weighted sums are accumulated by the kernel that
        assembles the net fluxes and applies the radiation; follow the call with diagnostics.all_reduce() when
        world_size > 1."""
        d = A.NeFusedStepDesc()
        if diagnostics is not None:
            d.diag = diagnostics.desc
        d.atmosphere = self.atmosphere_interp_desc(t)
        if self.radiation is not None:
            d.radiation = self.radiation_interp_desc(t)
        d.ao = self.atmosphere_ocean_desc()
        d.assemble = self.assemble_ocean_desc()
        if self.radiation is not None:
            d.apply_radiation = self.apply_radiation_desc(False)
        return d

    def fused_interface_step(self, t, diagnostics=None, ocean_column=None):
        """Interpolation -> a–o solve -> net ocean flux assembly -> radiation (-> diagnostics sums) for an
        OceanOnlyModel, one C-ABI call.  `ocean_column` (T3, S3, dz, dt, nz, hz): the FreezingLimitedOceanTemperature clamp
        of the 3-D ocean temperature runs behind it (it only touches the column and frazil_heat, which nothing else in the
        step reads: its place in phase 2 or after phase 4 makes no difference)."""
        self.clock_time = t
        if self.atmosphere_correction is not None:   # phase 1.5 sits between the phases the fused call merges
            d, s, FT = self.fused_step_desc(t), self.backend.stream(), self.grid.FT
            self.interpolate_state(t)
            self.correct_state()
            self.lib.call("atmosphere_ocean_fluxes", FT, d.ao, s)
            self.lib.call("assemble_net_ocean_fluxes", FT, d.assemble, s)
            if self.radiation is not None:
                self.lib.call("apply_radiative_fluxes", FT, d.apply_radiation, s)
            if diagnostics is not None:
                diagnostics.reduce()
            self.compute_sea_ice_ocean_fluxes(ocean_column)
            return
        if self.land is not None:   # the runoff interpolation is independent of everything else in phase 1
            self.lib.call("interp_state", self.grid.FT, self.land_interp_desc(t), self.backend.stream())
        self.lib.call("fused_interface_step", self.grid.FT, self.fused_step_desc(t, diagnostics), self.backend.stream())
        self.release_windows()
        self.compute_sea_ice_ocean_fluxes(ocean_column)
        if diagnostics is not None:
            diagnostics.all_reduce()
