"""Synthetic JRA55/ECCO-shaped inputs (there is no network: SURVEY.md §8(d)).

Every array is generated on the host with numpy `default_rng(seed)` so the oracle, the CUDA library
and (if it ever becomes available) the Julia reference read identical bits.

Configs (BASELINE.json): C1 1° 360x150 (lat ±75), C2/C3 1/4° 1440x560 (lat ±70), C4 1/12° 4320x1680,
C5 1/48° 17280x6720; atmosphere 640x320, halo 3, 3-hourly, Float32 (JRA55-faithful) or Float64.
"""
import numpy as np

from . import formulations as F
from .interface import (ComponentInterfaces, ExchangeGrid, LatLonSourceGrid, PrescribedAtmosphere,
                        PrescribedLand, PrescribedRadiation)

CONFIGS = {
    "C1": dict(nx=360, ny=150, latitude=(-75.0, 75.0)),
    "C2": dict(nx=1440, ny=560, latitude=(-70.0, 70.0)),
    "C3": dict(nx=1440, ny=560, latitude=(-70.0, 70.0)),
    "C4": dict(nx=4320, ny=1680, latitude=(-70.0, 70.0)),
    "C5": dict(nx=17280, ny=6720, latitude=(-70.0, 70.0)),
    "tiny": dict(nx=48, ny=20, latitude=(-75.0, 75.0)),
}
BASE_SEED = 20261017


def _q_sat(T, p):
    """Clausius–Clapeyron q_sat used only to make plausible humidity inputs."""
    Rv, Rd = 8.3144598 / 0.018015, 8.3144598 / 0.02897
    dcp = 1859.0 - 4181.0
    psat = 611.657 * (T / 273.16) ** (dcp / Rv) * np.exp((2500800 - dcp * 273.16) / Rv * (1 / 273.16 - 1 / T))
    e = Rd / Rv
    return e * psat / (p - (1 - e) * psat)


def gaussian_like_latitudes(ny):
    """Stretched (Gaussian-grid-like) latitude centres: exercises the binary search path."""
    k = np.arange(ny) + 0.5
    return 90.0 * np.sin((k / ny - 0.5) * np.pi * 0.985) / np.sin(0.5 * np.pi * 0.985) * 0.9972


def atmosphere_arrays(src: LatLonSourceGrid, nt, rng, hours=3.0):
    """Dict of (nt, ny+2hy, nx+2hx) arrays in the source dtype with periodic-x halos filled
    (JRA55_field_time_series.jl:74) and y halos extended by replication."""
    npd = np.float64 if src.FT == "f64" else np.float32
    ny, nx = src.ny, src.nx
    phi = np.deg2rad(src.phi_nodes.astype(np.float64))[:, None]
    lam = np.deg2rad(src.lam_nodes.astype(np.float64))[None, :]
    out = {}
    # large-scale warm/cold anomaly so that both stable and unstable stratification occur over the ocean
    base_T = 296 - 30 * np.sin(phi) ** 2 + 9 * np.sin(2 * lam + 1.0) * np.cos(phi) + rng.normal(0, 2, (ny, nx))
    base_p = 101325 + rng.normal(0, 800, (ny, nx))
    rh = rng.uniform(0.5, 0.95, (ny, nx))
    base_u = 8 * np.cos(3 * phi) + rng.normal(0, 4, (ny, nx))
    base_v = rng.normal(0, 3, (ny, nx))
    calm = rng.uniform(size=(ny, nx)) < 0.02
    base_u[calm] = rng.uniform(-0.1, 0.1, calm.sum())
    base_v[calm] = rng.uniform(-0.1, 0.1, calm.sum())
    rain0 = 1e-5 * rng.uniform(0, 1, (ny, nx)) * (rng.uniform(size=(ny, nx)) < 0.2)
    snow0 = 1e-5 * rng.uniform(0, 1, (ny, nx)) * (rng.uniform(size=(ny, nx)) < 0.2) * (np.abs(np.rad2deg(phi)) > 60)
    lw0 = 340 + rng.normal(0, 40, (ny, nx))
    sw_noise = rng.normal(0, 30, (ny, nx))
    fields = {k: np.zeros((nt, ny, nx)) for k in ("u", "v", "T", "q", "p", "rain", "snow", "sw", "lw")}
    for n in range(nt):
        drift = np.sin(2 * np.pi * n * hours / 24.0 + lam)  # smooth 3-hourly drift so ñ matters
        T = base_T + 1.5 * drift
        p = base_p + 50 * drift
        fields["T"][n] = T
        fields["p"][n] = p
        fields["q"][n] = rh * _q_sat(T, p)
        fields["u"][n] = base_u + 0.8 * drift
        fields["v"][n] = base_v + 0.5 * np.cos(2 * np.pi * n * hours / 24.0 + lam)
        fields["rain"][n] = rain0 * (1 + 0.3 * drift)
        fields["snow"][n] = snow0 * (1 + 0.3 * drift)
        hour_angle = 2 * np.pi * n * hours / 24.0 + lam
        fields["sw"][n] = np.maximum(0.0, 320 * np.cos(phi) * np.cos(hour_angle) + sw_noise)
        fields["lw"][n] = lw0 + 5 * drift
    for k, a in fields.items():
        full = np.zeros((nt, ny + 2 * src.hy, nx + 2 * src.hx), dtype=npd)
        full[:, src.hy:src.hy + ny, src.hx:src.hx + nx] = a
        # periodic x halos
        full[:, :, :src.hx] = full[:, :, nx:nx + src.hx]
        full[:, :, nx + src.hx:] = full[:, :, src.hx:2 * src.hx]
        # y halos: replicate the boundary rows
        full[:, :src.hy, :] = full[:, src.hy:src.hy + 1, :]
        full[:, src.hy + ny:, :] = full[:, src.hy + ny - 1:src.hy + ny, :]
        out[k] = full
    return out


def inactive_mask(grid: ExchangeGrid):
    """~30 % inactive from a thresholded low-wavenumber field (+ everything beyond the Bounded-y domain).
    Deterministic (no random numbers): the mask of a latitude band is a slice of the global one."""
    shape = grid.shape
    phi = np.deg2rad(grid.phi.astype(np.float64))[:, None] * np.ones((1, shape[1]))
    lam = np.deg2rad(grid.lam.astype(np.float64))[None, :] * np.ones((shape[0], 1))
    low = (np.sin(2 * lam + 0.3) * np.cos(3 * phi) + 0.6 * np.sin(5 * lam - 1.0) * np.sin(2 * phi + 0.5) +
           0.4 * np.cos(lam * 3 + phi * 4))
    inactive = low > np.quantile(low, 0.70)
    j = np.arange(shape[0])
    outside = (j < grid.hy) | (j >= grid.hy + grid.ny)
    inactive[outside, :] = True
    return inactive.astype(np.uint8)


def row_cost_weights(config, FT="f64", active_cost=4.3):
    """Per-row cost estimate of the interface step on the GLOBAL grid of `config` for
    sharding.latitude_bands(weights=...): every point pays the HBM-bound kernels (weight 1), an active
    point additionally the similarity-theory solve (measured ~4.3x that on B200, profiles/r01_notes.md)."""
    cfg = CONFIGS[config] if isinstance(config, str) else config
    g = ExchangeGrid(nx=cfg["nx"], ny=cfg["ny"], hx=cfg.get("hx", 7), hy=cfg.get("hy", 7), latitude=cfg["latitude"], FT=FT)
    m = inactive_mask(g)[g.hy:g.hy + g.ny, g.hx:g.hx + g.nx]
    return g.nx + active_cost * (m == 0).sum(axis=1).astype(np.float64)


def land_arrays(src: LatLonSourceGrid, nt, rng):
    """JRA55-shaped runoff: two non-negative, sparse series (rivers, icebergs) on the land grid, (nt, ny+2hy, nx+2hx),
    periodic x halos filled."""
    npd = np.float64 if src.FT == "f64" else np.float32
    ny, nx = src.ny, src.nx
    out = []
    for scale, frac in ((2e-4, 0.04), (5e-5, 0.01)):
        base = scale * rng.uniform(0, 1, (ny, nx)) * (rng.uniform(size=(ny, nx)) < frac)
        a = np.stack([base * (1 + 0.2 * np.sin(2 * np.pi * n / max(nt, 1))) for n in range(nt)])
        full = np.zeros((nt, ny + 2 * src.hy, nx + 2 * src.hx), dtype=npd)
        full[:, src.hy:src.hy + ny, src.hx:src.hx + nx] = a
        full[:, :, :src.hx] = full[:, :, nx:nx + src.hx]
        full[:, :, nx + src.hx:] = full[:, :, src.hx:2 * src.hx]
        full[:, :src.hy, :] = full[:, src.hy:src.hy + 1, :]
        full[:, src.hy + ny:, :] = full[:, src.hy + ny - 1:src.hy + ny, :]
        out.append(full)
    return out


def rotation_arrays(grid: ExchangeGrid):
    """A smooth synthetic rotation angle θ(λ, φ) ∈ (−π, π) in the exchange layout, as (cos θ, sin θ): stands in for the
    rotation metrics of a tripolar grid (θ grows towards the northern fold)."""
    phi = np.deg2rad(grid.phi.astype(np.float64))[:, None]
    lam = np.deg2rad(grid.lam.astype(np.float64))[None, :]
    theta = 2.5 * np.sin(lam) * np.clip((np.rad2deg(phi) - 20.0) / 60.0, 0.0, 1.0) ** 2 + 0.05 * np.cos(2 * lam) * np.cos(phi)
    npd = np.float64 if grid.FT == "f64" else np.float32
    return np.cos(theta).astype(npd), np.sin(theta).astype(npd)


def ocean_arrays(grid: ExchangeGrid, rng, sea_ice=False):
    """Exchange-layout (ny+2hy, nx+2hx) ocean surface / sea-ice state + inactive mask."""
    npd = np.float64 if grid.FT == "f64" else np.float32
    shape = grid.shape
    phi = np.deg2rad(grid.phi.astype(np.float64))[:, None] * np.ones((1, shape[1]))
    lam = np.deg2rad(grid.lam.astype(np.float64))[None, :] * np.ones((shape[0], 1))
    o = {}
    o["T"] = 28 * np.cos(phi) ** 2 - 1.8 + rng.normal(0, 0.5, shape)
    o["S"] = 35 + rng.normal(0, 0.6, shape)
    o["u"] = rng.normal(0, 0.15, shape)
    o["v"] = rng.normal(0, 0.15, shape)
    o["inactive"] = inactive_mask(grid)
    if sea_ice:
        latd = np.abs(np.rad2deg(phi))
        conc = np.clip((latd - 62) / 6, 0, 1) * rng.uniform(0.8, 1.0, shape)
        o["concentration"] = conc
        o["hi"] = 1.5 * conc
        o["hs"] = 0.1 * conc
        o["hc"] = np.full(shape, 0.05)
        o["Si"] = np.full(shape, 4.0)
        o["top_temperature"] = np.full(shape, -5.0) + rng.normal(0, 1.0, shape)
        o["ice_mass_flux"] = 1e-5 * rng.normal(0, 1, shape) * (conc > 0)
        o["snow_mass_flux"] = 1e-6 * rng.uniform(0, 1, shape) * (conc > 0)
    return {k: (v if v.dtype == np.uint8 else v.astype(npd)) for k, v in o.items()}


def build_case(config, backend, FT="f64", atm_FT="f64", nt=2, seed_offset=0, sea_ice=False, radiation=True,
               stretched_latitude=False, lib=None, with_iterations=False, grid=None, land=False, rotated=False,
               local_surface=False, **interface_kwargs):
    """Construct a fully populated ComponentInterfaces for one BASELINE.json config on `backend`.
    local_surface: a latitude band draws its own surface state (seeded by its row offset) instead of slicing the global
    one — for throughput measurements of grids whose global state would not be worth generating on every rank."""
    cfg = CONFIGS[config] if isinstance(config, str) else config
    if grid is None:
        grid = ExchangeGrid(nx=cfg["nx"], ny=cfg["ny"], hx=cfg.get("hx", 7), hy=cfg.get("hy", 7),
                            latitude=cfg["latitude"], FT=FT)
    if rotated:
        grid.rotation = rotation_arrays(grid)
    rng = np.random.default_rng(BASE_SEED + seed_offset)
    src = LatLonSourceGrid(nx=cfg.get("src_nx", 640), ny=cfg.get("src_ny", 320), FT=atm_FT,
                           phi_nodes=gaussian_like_latitudes(cfg.get("src_ny", 320)) if stretched_latitude else None,
                           y_regular=not stretched_latitude)
    times = np.arange(nt, dtype=np.float64) * 10800.0
    a = atmosphere_arrays(src, nt, rng)
    dev = {k: backend.from_numpy(v) for k, v in a.items()}
    atm = PrescribedAtmosphere(grid=src, times=times, u=dev["u"], v=dev["v"], T=dev["T"], q=dev["q"], p=dev["p"],
                               rain=(dev["rain"],), snow=(dev["snow"],))
    rad = PrescribedRadiation(grid=src, times=times, downwelling_shortwave=dev["sw"], downwelling_longwave=dev["lw"]) \
        if radiation else None
    if grid.ny_global is not None and local_surface:
        o = ocean_arrays(grid, np.random.default_rng(BASE_SEED + seed_offset + 7919 * (1 + grid.j_offset)), sea_ice=sea_ice)
    elif grid.ny_global is not None:   # latitude band: generate the GLOBAL surface state, keep this band's rows
        gg = ExchangeGrid(nx=grid.nx, ny=grid.ny_global, hx=grid.hx, hy=grid.hy, latitude=grid.latitude, FT=grid.FT)
        og = ocean_arrays(gg, rng, sea_ice=sea_ice)
        o = {k: np.ascontiguousarray(v[grid.j_offset:grid.j_offset + grid.ny + 2 * grid.hy, :]) for k, v in og.items()}
    else:
        o = ocean_arrays(grid, rng, sea_ice=sea_ice)
    land_component = None
    if land:   # after every other draw, so the atmosphere / ocean inputs do not depend on it
        lsrc = LatLonSourceGrid(nx=cfg.get("land_nx", 1440), ny=cfg.get("land_ny", 720), FT=atm_FT)
        ltimes = np.arange(nt, dtype=np.float64) * 86400.0
        lrng = np.random.default_rng(BASE_SEED + 5000 + seed_offset)
        land_component = PrescribedLand(grid=lsrc, times=ltimes,
                                        freshwater_flux=tuple(backend.from_numpy(x) for x in land_arrays(lsrc, nt, lrng)))
    ci = ComponentInterfaces(grid, backend, atm, rad, sea_ice=sea_ice, lib=lib, inactive=backend.from_numpy(o["inactive"]),
                             with_iterations=with_iterations, land=land_component, **interface_kwargs)
    ci.ocean_state.u, ci.ocean_state.v = backend.from_numpy(o["u"]), backend.from_numpy(o["v"])
    ci.ocean_state.T, ci.ocean_state.S = backend.from_numpy(o["T"]), backend.from_numpy(o["S"])
    if sea_ice:
        s = ci.sea_ice_state
        s.concentration, s.hi, s.hs, s.hc = (backend.from_numpy(o[k]) for k in ("concentration", "hi", "hs", "hc"))
        s.S = backend.from_numpy(o["Si"])
        s.top_temperature = backend.from_numpy(o["top_temperature"])
        s.ice_mass_flux, s.snow_mass_flux = backend.from_numpy(o["ice_mass_flux"]), backend.from_numpy(o["snow_mass_flux"])
    ci._host_inputs = dict(atmosphere=a, ocean=o)
    return ci


def ocean_column(grid: ExchangeGrid, backend, nz=10, seed_offset=0):
    """3-D ocean T, S columns (nz, ny+2hy, nx+2hx) for the frazil kernels; ~5 % of cells supercooled."""
    npd = np.float64 if grid.FT == "f64" else np.float32
    rng = np.random.default_rng(BASE_SEED + 1000 + seed_offset)
    shape = (nz,) + grid.shape
    phi = np.deg2rad(grid.phi.astype(np.float64))[None, :, None]
    T = 28 * np.cos(phi) ** 2 - 3.6 + rng.normal(0, 0.4, shape)   # supercooled at the highest latitudes
    S = 35 + rng.normal(0, 0.6, shape)
    dz = np.linspace(50.0, 5.0, nz)   # k = 1 (bottom) .. nz (top)
    return backend.from_numpy(T.astype(npd)), backend.from_numpy(S.astype(npd)), backend.from_numpy(dz.astype(npd))
