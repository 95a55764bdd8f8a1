"""PrescribedLand runoff interpolation (Lands/interpolate_land_state.jl:6-61) and the
intrinsic_vector rotation of the interpolated wind on rotated exchange grids
(Atmospheres/interpolate_atmospheric_state.jl:123-126).  Oracle properties on the CPU, CUDA-vs-oracle parity on the GPU."""
import numpy as np
import pytest

import ne_b200
from numericalearth_jl_b200 import synthetic

T_STEP = 0.37 * 10800.0
CFG = dict(nx=96, ny=40, latitude=(-70.0, 70.0), land_nx=180, land_ny=90)


def _host_case(oracle_lib, **kw):
    ci = synthetic.build_case(CFG, ne_b200.NumpyHostBackend(), lib=oracle_lib, **kw)
    ci.initialize()
    ci.update_state(T_STEP)
    return ci


@pytest.mark.parametrize("FT,atm_FT", [("f64", "f64"), ("f64", "f32"), ("f32", "f32")])
def test_runoff_is_the_sum_of_its_series_and_enters_the_freshwater_flux(oracle_lib, FT, atm_FT):
    ci = _host_case(oracle_lib, FT=FT, atm_FT=atm_FT, land=True)
    g = ci.grid
    total = g.interior(ci.land_state.freshwater_flux).astype(np.float64)
    assert (total >= 0).all() and (total > 0).mean() > 0.01
    # each series alone (the other replaced by `nothing` → the literal 0, interpolate_atmospheric_state.jl:143)
    parts = []
    series = ci.land.freshwater_flux
    for k in range(len(series)):
        ci.land.freshwater_flux = (series[k],)
        ci.interpolate_state(T_STEP)
        parts.append(g.interior(ci.land_state.freshwater_flux).astype(np.float64).copy())
    ci.land.freshwater_flux = series
    ci.interpolate_state(T_STEP)
    tol = 0 if FT == "f64" else 2e-7    # Float64 output: the partial interpolants are exact doubles, their sum is the same add
    assert np.abs(parts[0] + parts[1] - total).max() <= tol * max(total.max(), 1e-300)
    # the runoff enters the net freshwater flux exactly like rain (assemble_net_ocean_fluxes.jl:103-110): ΔJw = +Jˡⁿ / ρ at active points
    ref = _host_case(oracle_lib, FT=FT, atm_FT=atm_FT, land=False)
    inner = (slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx))
    active = np.asarray(ci.inactive)[inner] == 0
    dJ = (np.asarray(ci.net_ocean.eta)[inner].astype(np.float64) - np.asarray(ref.net_ocean.eta)[inner].astype(np.float64))
    Jl = np.asarray(ci.land_state.freshwater_flux)[inner].astype(np.float64)
    rho = ci.ocean_properties.reference_density
    scale = float(np.abs(np.asarray(ci.net_ocean.eta)[inner]).max())
    assert np.abs(dJ[active] - Jl[active] / rho).max() <= (1e-15 if FT == "f64" else 1e-6) * scale
    assert (dJ[~active] == 0).all()


@pytest.mark.parametrize("FT,atm_FT", [("f64", "f64"), ("f64", "f32"), ("f32", "f32")])
def test_rotation_of_the_interpolated_wind(oracle_lib, FT, atm_FT):
    plain = _host_case(oracle_lib, FT=FT, atm_FT=atm_FT)
    rot = _host_case(oracle_lib, FT=FT, atm_FT=atm_FT, rotated=True)
    g = plain.grid
    c, s = (g.interior(a).astype(np.float64) for a in rot.grid.rotation)
    u, v = g.interior(plain.atmos_state.u).astype(np.float64), g.interior(plain.atmos_state.v).astype(np.float64)
    ur, vr = g.interior(rot.atmos_state.u).astype(np.float64), g.interior(rot.atmos_state.v).astype(np.float64)
    eps = 3e-16 if FT == "f64" and atm_FT == "f64" else 2e-7
    scale = max(np.abs(u).max(), np.abs(v).max())
    assert np.abs(ur - (u * c + v * s)).max() <= eps * scale
    assert np.abs(vr - (-u * s + v * c)).max() <= eps * scale
    assert np.abs((ur ** 2 + vr ** 2) - (u ** 2 + v ** 2)).max() <= 4 * eps * scale ** 2 + 1e-6 * (FT == "f32") * scale ** 2
    for n in ("T", "q", "p", "Jrn", "Jsn"):   # scalars are untouched
        assert np.array_equal(getattr(plain.atmos_state, n), getattr(rot.atmos_state, n)), n
    # θ = 0: bit-identical to the latitude-longitude path
    ident = synthetic.build_case(CFG, ne_b200.NumpyHostBackend(), lib=oracle_lib, FT=FT, atm_FT=atm_FT)
    ident.grid.rotation = (np.ones(g.shape), np.zeros(g.shape))
    ident.initialize(); ident.update_state(T_STEP)
    for n in ("u", "v"):
        assert np.array_equal(g.interior(getattr(plain.atmos_state, n)), g.interior(getattr(ident.atmos_state, n))), n


@pytest.mark.gpu
@pytest.mark.parametrize("FT,atm_FT", [("f64", "f64"), ("f64", "f32"), ("f32", "f32")])
def test_cuda_land_and_rotation_parity(oracle_lib, cuda_backend, cuda_lib, FT, atm_FT):
    """Runoff interpolation and rotated wind: bit-exact against the oracle; the fluxes downstream to the usual bar."""
    ref = synthetic.build_case(CFG, ne_b200.NumpyHostBackend(), lib=oracle_lib, FT=FT, atm_FT=atm_FT, land=True, rotated=True)
    dev = synthetic.build_case(CFG, cuda_backend, FT=FT, atm_FT=atm_FT, land=True, rotated=True)
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP); dev.update_state(T_STEP)
    cuda_backend.synchronize()
    g = ref.grid
    to = cuda_backend.to_numpy
    assert np.array_equal(g.interior(ref.land_state.freshwater_flux), g.interior(to(dev.land_state.freshwater_flux)))
    for n in ref.atmos_state.names():
        assert np.array_equal(g.interior(getattr(ref.atmos_state, n)), g.interior(to(getattr(dev.atmos_state, n)))), n
    tol = 1e-10 if (FT, atm_FT) == ("f64", "f64") else (2e-6 if FT == "f64" else 1e-5)
    for bag in ("ao_fluxes", "net_ocean"):
        for n in getattr(ref, bag).names():
            a, b = np.asarray(getattr(getattr(ref, bag), n), dtype=np.float64), to(getattr(getattr(dev, bag), n)).astype(np.float64)
            a, b = a[g.hy:g.hy + g.ny, g.hx:g.hx + g.nx], b[g.hy:g.hy + g.ny, g.hx:g.hx + g.nx]
            sc = float(np.abs(a).max()) or 1.0
            assert np.abs(a - b).max() / sc <= tol, f"{bag}.{n}: {np.abs(a - b).max() / sc}"
    # the fused step and the host pipeline take the same inputs
    dev2 = synthetic.build_case(CFG, cuda_backend, FT=FT, atm_FT=atm_FT, land=True, rotated=True)
    dev2.initialize()
    dev2.fused_interface_step(T_STEP)
    cuda_backend.synchronize()
    for n in dev.net_ocean.names():
        x, y = to(getattr(dev.net_ocean, n)), to(getattr(dev2.net_ocean, n))
        sc = float(np.abs(x).max()) or 1.0
        assert float(np.abs(x.astype(np.float64) - y).max()) <= 1e-13 * sc, n
