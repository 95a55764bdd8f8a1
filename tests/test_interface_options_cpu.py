"""Host-mirror wiring of the step options added in round 2, run on the oracle (no GPU) and pinned by plain numpy
restatements of the reference formulas — the same cases tests/test_gpu_paths.py runs on the device:

  FreezingLimitedOceanTemperature clamp + frazil heat   src/SeaIces/freezing_limited_ocean_temperature.jl:73-118
  barotropic potential = p / rho_ocean                  src/Atmospheres/interpolate_atmospheric_state.jl:80-85
  TwoColorRadiation shortwave routing                   src/Oceans/radiative_forcing.jl:84-91
  LatitudeDependentAlbedo                               src/Radiations/latitude_dependent_albedo.jl:48-53
  MomentumBasedFrictionVelocity / IceBathHeatFlux       …/friction_velocity.jl:24-44, …heat_flux_formulations.jl:176-195
"""
import numpy as np
import pytest

import ne_b200
from numericalearth_jl_b200 import synthetic

T_STEP = 0.37 * 10800.0
CFG = dict(nx=48, ny=20, latitude=(-75.0, 75.0))


def _case(oracle_lib, host_backend, **kw):
    ci = synthetic.build_case(CFG, host_backend, FT="f64", atm_FT="f64", lib=oracle_lib, **kw)
    ci.initialize()
    return ci


@pytest.mark.parametrize("dt", [1200.0, float("inf")])
def test_freezing_limited_clamp_known_answer(oracle_lib, host_backend, dt):
    ci = _case(oracle_lib, host_backend)
    g = ci.grid
    T3, S3, dz = synthetic.ocean_column(g, host_backend, nz=6)
    T3 -= 0.5                                                # supercool the polar rows
    T0 = T3.copy()
    ci.update_state(T_STEP, ocean_column=(T3, S3, dz, dt, 6, 0))
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    Tm = 0.0 - 0.054 * S3                                    # LinearLiquidus defaults (ClimaSeaIce, third party)
    freezing = T0 < Tm
    expect_T = np.where(freezing, Tm, T0)
    assert np.array_equal(T3[:, rows, cols], expect_T[:, rows, cols])
    assert np.array_equal(T3[:, :g.hy, :], T0[:, :g.hy, :])   # halos are not touched (launch range :xy)
    rho, c = ci.ocean_properties.reference_density, ci.ocean_properties.heat_capacity
    q = np.zeros(g.shape)
    for k in range(5, -1, -1):                               # k = Nz:-1:1, same accumulation order
        dE = freezing[k] * rho * c * (Tm[k] - T0[k])
        q = q - dE * dz[k] / dt
    assert np.array_equal(ci.frazil_heat[rows, cols], q[rows, cols])
    assert freezing[:, rows, cols].sum() > 10
    if np.isinf(dt):
        assert (ci.frazil_heat[rows, cols] == 0).all()


def test_clamp_is_its_own_phase(oracle_lib, host_backend):
    """compute_sea_ice_ocean_fluxes only touches the T column and frazil_heat: calling it on its own after a step
    without a column equals the step that was given the column."""
    a, b = _case(oracle_lib, host_backend), _case(oracle_lib, host_backend)
    ca = synthetic.ocean_column(a.grid, host_backend, nz=4) + (600.0, 4, 0)
    cb = synthetic.ocean_column(b.grid, host_backend, nz=4) + (600.0, 4, 0)
    a.update_state(T_STEP, ocean_column=ca)
    b.update_state(T_STEP)
    assert b.frazil_heat is None
    b.compute_sea_ice_ocean_fluxes(cb)
    assert np.array_equal(ca[0], cb[0]) and np.array_equal(a.frazil_heat, b.frazil_heat)
    for n in a.net_ocean.names():
        assert np.array_equal(getattr(a.net_ocean, n), getattr(b.net_ocean, n), equal_nan=True)


def test_barotropic_potential_known_answer(oracle_lib, host_backend):
    ci = _case(oracle_lib, host_backend, barotropic_potential=True)
    ci.interpolate_state(T_STEP)
    g = ci.grid
    assert np.array_equal(g.interior(ci.barotropic_potential), g.interior(ci.atmos_state.p) / ci.ocean_properties.reference_density)
    plain = _case(oracle_lib, host_backend)
    plain.interpolate_state(T_STEP)
    assert plain.barotropic_potential is None
    assert np.array_equal(plain.atmos_state.p, ci.atmos_state.p)


def test_two_color_routing_known_answer(oracle_lib, host_backend):
    tc, plain = _case(oracle_lib, host_backend, two_color_radiation=True), _case(oracle_lib, host_backend)
    tc.update_state(T_STEP); plain.update_state(T_STEP)
    g = tc.grid
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    rho, c = tc.ocean_properties.reference_density, tc.ocean_properties.heat_capacity
    tr = plain.rad_fluxes_ocean.downwelling_shortwave[rows, cols]          # transmitted shortwave, −(1 − α) SW (1 − ℵ)
    act = plain.inactive[rows, cols] == 0
    assert np.array_equal(tc.two_color_surface_flux[rows, cols][act], (tr / (rho * c))[act])   # surface_flux = −ℐtr/(ρc) with ℐtr = −tr …
    # J⁰ = −ℐₜ/(ρc) is what JT no longer receives: JT(plain) = JT(two colour) + ℐₜ/(ρc) = JT(two colour) − J⁰
    total = tc.net_ocean.T[rows, cols] - tc.two_color_surface_flux[rows, cols]
    # (J⁰ is stored at masked cells too, :84-91 runs before the inactive mask; JT is masked)
    assert np.abs(total - plain.net_ocean.T[rows, cols])[act].max() <= 2e-15 * np.abs(plain.net_ocean.T[rows, cols]).max()
    for n in ("upwelling_longwave", "downwelling_longwave", "downwelling_shortwave"):
        assert np.array_equal(getattr(tc.rad_fluxes_ocean, n), getattr(plain.rad_fluxes_ocean, n))


def test_latitude_dependent_albedo_known_answer(oracle_lib, host_backend):
    ci = synthetic.build_case(CFG, host_backend, FT="f64", atm_FT="f64", lib=oracle_lib)
    ci.radiation.surface_properties["ocean"] = ne_b200.SurfaceRadiationProperties(ne_b200.LatitudeDependentAlbedo(0.069, 0.011), 0.97)
    ci.initialize()
    ci.update_state(T_STEP)
    g = ci.grid
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    sw, tr = ci.rad_state.sw[rows, cols], ci.rad_fluxes_ocean.downwelling_shortwave[rows, cols]
    act = (ci.inactive[rows, cols] == 0) & (sw > 1.0)
    phi = np.deg2rad(g.phi[rows])[:, None] * np.ones_like(sw)
    alpha = 0.069 - 0.011 * np.cos(2 * phi)
    assert act.sum() > 50
    assert np.allclose(tr[act], ((1 - alpha) * sw)[act], rtol=1e-14, atol=0)   # the stored diagnostic is −ℐₜ = (1 − α) SW (:108)


@pytest.mark.parametrize("formulation", ["ice_bath", "ice_bath_momentum", "three_equation_momentum"])
def test_sea_ice_ocean_variants_known_answer(oracle_lib, host_backend, formulation):
    mom = ne_b200.MomentumBasedFrictionVelocity()
    ff = {"ice_bath": ne_b200.IceBathHeatFlux(), "ice_bath_momentum": ne_b200.IceBathHeatFlux(friction_velocity=mom),
          "three_equation_momentum": ne_b200.ThreeEquationHeatFlux(friction_velocity=mom)}[formulation]
    ci = _case(oracle_lib, host_backend, sea_ice=True, sea_ice_ocean_heat_flux=ff)
    g = ci.grid
    rng = np.random.default_rng(17)
    ci.sio_fluxes.x_momentum[...] = 0.05 * rng.standard_normal(g.shape)
    ci.sio_fluxes.y_momentum[...] = 0.05 * rng.standard_normal(g.shape)
    T3, S3, dz = synthetic.ocean_column(g, host_backend, nz=4)
    ci.update_state(T_STEP, ocean_column=(T3, S3, dz, 900.0, 4, 0))
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    rho, c = ci.ocean_properties.reference_density, ci.ocean_properties.heat_capacity
    tx, ty = ci.sio_fluxes.x_momentum, ci.sio_fluxes.y_momentum
    # τ at cell centres: ℑx of τx² (faces i, i+1), ℑy of τy² (faces j, j+1) (friction_velocity.jl:26-31)
    tx2 = 0.5 * (tx[rows, cols] ** 2 + tx[rows, g.hx + 1:g.hx + g.nx + 1] ** 2)
    ty2 = 0.5 * (ty[rows, cols] ** 2 + ty[g.hy + 1:g.hy + g.ny + 1, cols] ** 2)
    ustar_mom = np.sqrt(np.sqrt(tx2 + ty2) / rho)
    conc = ci.sea_ice_state.concentration[rows, cols]
    if formulation.startswith("ice_bath"):
        us = ustar_mom if formulation.endswith("momentum") else 0.02
        Ttop, Stop = T3[3][rows, cols], S3[3][rows, cols]       # after the clamp
        Q = rho * c * 0.006 * us * (Ttop - (-0.054 * Stop)) * conc
        got = ci.sio_fluxes.interface_heat[rows, cols]
        assert np.allclose(got, Q, rtol=1e-13, atol=1e-9 * np.abs(Q).max())
        assert (np.abs(got) > 0).sum() > 20
    else:
        q = ci.sio_fluxes.interface_heat[rows, cols]
        assert np.isfinite(q).all() and (q[conc > 0] != 0).any()
        # a constant u★ equal to the momentum-based one at a point must reproduce that point
        i, j = np.argwhere(conc > 0.5)[0]
        alt = _case(oracle_lib, host_backend, sea_ice=True,
                    sea_ice_ocean_heat_flux=ne_b200.ThreeEquationHeatFlux(friction_velocity=float(ustar_mom[i, j])))
        alt.sio_fluxes.x_momentum[...] = tx; alt.sio_fluxes.y_momentum[...] = ty
        T3b, S3b, dzb = synthetic.ocean_column(g, host_backend, nz=4)
        alt.update_state(T_STEP, ocean_column=(T3b, S3b, dzb, 900.0, 4, 0))
        assert alt.sio_fluxes.interface_heat[rows, cols][i, j] == pytest.approx(q[i, j], rel=1e-13)
