"""Surface albedo kernel variants: SeaIceAlbedo known answers of the reference's own
test/test_sea_ice_albedo.jl:36-120 and TabulatedAlbedo against an independent numpy restatement of
src/Radiations/tabulated_albedo.jl:109-160 — on the oracle (CPU) and, marked gpu, CUDA vs oracle."""
import math

import numpy as np
import pytest

import ne_b200
from numericalearth_jl_b200 import synthetic

CFG = dict(nx=24, ny=12, hx=2, hy=2, latitude=(-80.0, 80.0))


def _albedo_through_the_kernel(ci, backend, over_sea_ice, t=0.0):
    """Run apply_radiative_fluxes with SW↓ = 1: the returned 'downwelling_shortwave' is (1 − α)·SW·(1 − ℵ or 1)."""
    ci.clock_time = t
    ci.rad_state.sw[...] = 1.0
    ci.rad_state.lw[...] = 0.0
    d = ci.apply_radiation_desc(over_sea_ice)
    ci.lib.call("apply_radiative_fluxes", ci.grid.FT, d, backend.stream())
    backend.synchronize()
    r = ci.rad_fluxes_sea_ice if over_sea_ice else ci.rad_fluxes_ocean
    g = ci.grid
    out = backend.to_numpy(r.downwelling_shortwave)[g.hy:g.hy + g.ny, g.hx:g.hx + g.nx]
    return 1.0 - out


def _sea_ice_case(lib, backend, hi, hs, Ts, snow=True):
    ci = synthetic.build_case(CFG, backend, FT="f64", atm_FT="f64", lib=lib, sea_ice=True)
    si = ci.sea_ice_state
    si.hi[...] = hi
    si.hs[...] = hs
    si.top_temperature[...] = Ts
    ci.radiation.surface_properties["sea_ice"] = ne_b200.SurfaceRadiationProperties(
        ne_b200.SeaIceAlbedo(si.hi, si.hs if snow else None, si.top_temperature), 1.0)
    return ci


@pytest.mark.parametrize("hi,hs,Ts,expected", [
    (2.0, 0.0, -20.0, 0.54),                        # cold thick ice, no snow            (:36-52)
    (2.0, 0.1, -20.0, 0.83),                        # deep snow                          (:54-68)
    (2.0, 0.0, 0.0, 0.54 - 0.075),                  # melting ice                        (:70-85)
    (0.0, 0.0, -10.0, 0.06),                        # no ice -> ocean albedo             (:87-101)
    (2.0, 0.01, -20.0, 0.5 * 0.83 + 0.5 * 0.54),    # partial snow cover                 (:103-119)
])
def test_sea_ice_albedo_known_answers(oracle_lib, host_backend, hi, hs, Ts, expected):
    ci = _sea_ice_case(oracle_lib, host_backend, hi, hs, Ts)
    alpha = _albedo_through_the_kernel(ci, host_backend, True)
    assert np.allclose(alpha, expected, rtol=1e-13, atol=0)


def test_sea_ice_albedo_without_snow_model(oracle_lib, host_backend):
    ci = _sea_ice_case(oracle_lib, host_backend, 2.0, 0.1, -20.0, snow=False)   # snow_thickness === nothing
    assert np.allclose(_albedo_through_the_kernel(ci, host_backend, True), 0.54, rtol=1e-13)


def _table(rng):
    phi_values = np.arange(0, 91, 2) / 180 * math.pi        # (0:2:90) ./ 180 * π
    t_values = np.arange(0, 1.0001, 0.05)                   # 0:0.05:1
    table = 0.03 + 0.4 * rng.random((len(phi_values), len(t_values)))   # [j (latitude), i (transmissivity)]
    return table, phi_values, t_values


def _tabulated_reference(table, phi_values, t_values, lam_deg, phi_deg, sw, t, S0=1365.0):
    """numpy restatement of tabulated_albedo.jl:109-160 (Float64)."""
    day = t // 86400
    sec = t - day * 86400
    lam, phi = np.deg2rad(lam_deg)[None, :], np.deg2rad(phi_deg)[:, None]
    h = (sec - 43200) * (2 * math.pi / 86400) + lam
    delta = math.radians((23 + 27 / 60) * math.sin(math.radians(360 * (day - 80) / 365.25)))
    cosz = np.maximum(0, np.sin(phi) * math.sin(delta) + np.cos(h) * math.cos(delta) * np.cos(phi))
    Qmax = S0 * cosz
    with np.errstate(divide="ignore", invalid="ignore"):
        tr = np.where(Qmax > 0, np.minimum(1, sw / Qmax), 0.0)
    fi = (tr - t_values[0]) / (t_values[1] - t_values[0])
    fj = (np.abs(phi) - phi_values[0]) / (phi_values[1] - phi_values[0]) + 0 * lam

    def interp(f):
        im = np.trunc(f).astype(int) + 1
        return im, im + np.sign(f).astype(int), np.mod(f, 1.0)
    im, ip, xi = interp(fi)
    jm, jp, eta = interp(fj)
    nphi, nt = table.shape   # zero-weighted i⁺/j⁺ past the edge: clamp the read

    def T(i, j):             # getindex(α_table, i, j), column-major (n_t, n_phi)
        return table[np.clip(j, 1, nphi) - 1, np.clip(i, 1, nt) - 1]
    return (1 - xi) * (1 - eta) * T(im, jm) + (1 - xi) * eta * T(im, jp) + xi * (1 - eta) * T(ip, jm) + xi * eta * T(ip, jp)


def _tabulated_case(lib, backend, rng):
    ci = synthetic.build_case(CFG, backend, FT="f64", atm_FT="f64", lib=lib)
    table, phi_values, t_values = _table(rng)
    ci.radiation.surface_properties["ocean"] = ne_b200.SurfaceRadiationProperties(
        ne_b200.TabulatedAlbedo(backend.from_numpy(table), phi_values, t_values), 0.97)
    return ci, table, phi_values, t_values


@pytest.mark.parametrize("t", [0.0, 30000.0, 86400.0 * 172 + 50000.0])
def test_tabulated_albedo_matches_numpy_restatement(oracle_lib, host_backend, t):
    rng = np.random.default_rng(3)
    ci, table, phi_values, t_values = _tabulated_case(oracle_lib, host_backend, rng)
    ci.clock_time = t
    g = ci.grid
    sw = 20.0 + 900.0 * rng.random(g.shape)
    ci.rad_state.sw[...] = sw
    ci.rad_state.lw[...] = 0.0
    ci.lib.call("apply_radiative_fluxes", g.FT, ci.apply_radiation_desc(False), host_backend.stream())
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    out = ci.rad_fluxes_ocean.downwelling_shortwave[rows, cols]          # (1 − α)·SW (no sea ice: ℵ = 0)
    alpha = 1.0 - out / sw[rows, cols]
    ref = _tabulated_reference(table, phi_values, t_values, g.lam[cols], g.phi[rows], sw[rows, cols], t)
    assert np.allclose(alpha, ref, rtol=0, atol=5e-13)
    assert 0.03 <= alpha.min() and alpha.max() <= 0.43


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sea_ice", "tabulated"])
def test_albedo_variants_gpu_parity(oracle_lib, host_backend, cuda_backend, cuda_lib, kind):
    rng = np.random.default_rng(5)
    outs = []
    for lib, backend in ((oracle_lib, host_backend), (cuda_lib, cuda_backend)):
        rng = np.random.default_rng(5)
        if kind == "sea_ice":
            ci = synthetic.build_case(CFG, backend, FT="f64", atm_FT="f64", lib=lib, sea_ice=True)
            g = ci.grid
            si = ci.sea_ice_state
            hi, hs, Ts = 1.2 * rng.random(g.shape), 0.05 * rng.random(g.shape), -3.0 + 3.5 * rng.random(g.shape)
            for dst, src in ((si.hi, hi), (si.hs, hs), (si.top_temperature, Ts)):
                dst.copy_(backend.from_numpy(src)) if backend.is_device else dst.__setitem__(Ellipsis, src)
            ci.radiation.surface_properties["sea_ice"] = ne_b200.SurfaceRadiationProperties(
                ne_b200.SeaIceAlbedo(si.hi, si.hs, si.top_temperature), 1.0)
            over = True
        else:
            ci, *_ = _tabulated_case(lib, backend, rng)
            over = False
        ci.initialize()
        ci.update_state(86400.0 * 200 + 41000.0)
        backend.synchronize()
        r = ci.rad_fluxes_sea_ice if over else ci.rad_fluxes_ocean
        outs.append({n: backend.to_numpy(getattr(r, n)).copy() for n in r.names()})
    for n in outs[0]:
        a, b = outs[0][n], outs[1][n]
        assert np.abs(a - b).max() <= 1e-10 * max(np.abs(a).max(), 1e-300), n
