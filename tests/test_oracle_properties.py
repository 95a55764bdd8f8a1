"""Oracle self-consistency and the analytic / limit cases of the reference's integration tests,
re-created with synthetic inputs (the originals read JRA55 values from a download).

    neutral analytic fluxes   test/test_surface_fluxes.jl:148-212
    zero-flux invariance      test/test_surface_fluxes.jl:81-110
    three-equation solver     test/test_sea_ice_ocean_heat_fluxes.jl:48-181, 342-385
    freezing-limited clamp    test/test_surface_fluxes.jl:258-292
"""
import ctypes as C
import math

import numpy as np
import pytest

import ne_b200
from ne_b200 import abi as A
from numericalearth_jl_b200 import synthetic

F = ne_b200


def _uniform_case(oracle_lib, host_backend, *, ua, va, Ta, qa, pa, To, So=30.0, uo=0.0, vo=0.0, **kw):
    cfg = dict(nx=4, ny=4, hx=2, hy=2, latitude=(-10.0, 10.0))
    g = ne_b200.ExchangeGrid(nx=4, ny=4, hx=2, hy=2, latitude=(-10.0, 10.0))
    ci = ne_b200.ComponentInterfaces(g, host_backend, None, None, lib=oracle_lib, with_iterations=True, **kw)
    for name, v in (("u", ua), ("v", va), ("T", Ta), ("q", qa), ("p", pa)):
        getattr(ci.atmos_state, name)[:] = v
    ci.ocean_state.T[:] = To; ci.ocean_state.S[:] = So; ci.ocean_state.u[:] = uo; ci.ocean_state.v[:] = vo
    return ci


def test_neutral_analytic_fluxes(oracle_lib, host_backend):
    """ψ ≡ 0, ℓ = 1e-4 for all three scales, no gustiness: u★ = ϰ/log(h/ℓ)·ΔU etc."""
    ell = 1e-4
    st = F.SimilarityTheoryFluxes(momentum_roughness_length=ell, temperature_roughness_length=ell,
                                  water_vapor_roughness_length=ell, subgrid_velocities=None, stability_functions=None)
    ua, va, Ta, qa, pa, To, So = 4.2, -3.1, 288.15, 0.003, 101000.0, 15.0, 30.0
    ci = _uniform_case(oracle_lib, host_backend, ua=ua, va=va, Ta=Ta, qa=qa, pa=pa, To=To, So=So, atmosphere_ocean_fluxes=st)
    ci.compute_atmosphere_ocean_fluxes()
    th = F.AtmosphereThermodynamicsParameters()
    Rd, Rv = th.gas_constant / th.dry_air_molar_mass, th.gas_constant / th.water_molar_mass
    cp_d = Rd / th.dry_air_adiabatic_exponent
    cp = cp_d * (1 - qa) + th.water_vapor_heat_capacity * qa
    rho = pa / ((Rd * (1 - qa) + Rv * qa) * Ta)
    Lv = th.reference_vaporization_enthalpy + (th.water_vapor_heat_capacity - th.liquid_water_heat_capacity) * (Ta - th.reference_temperature)
    h, g, kappa = 10.0, 9.80665, 0.4
    Tk = To + 273.15
    ip = ci.ao_properties.pod(); thp = th.pod()
    qo = oracle_lib.dll.neo_surface_specific_humidity_f64(C.byref(ip), C.byref(thp), pa, Tk, So)
    dU = math.hypot(ua, va)
    dth = Ta - Tk + h / cp * g
    dq = qa - qo
    us, ths, qs = (kappa / math.log(h / ell) * x for x in (dU, dth, dq))
    f = ci.ao_fluxes
    i = (3, 3)
    assert f.friction_velocity[i] == pytest.approx(us, rel=1e-12)
    assert f.temperature_scale[i] == pytest.approx(ths, rel=1e-10)
    assert f.water_vapor_scale[i] == pytest.approx(qs, rel=1e-10)
    assert f.x_momentum[i] == pytest.approx(-rho * us ** 2 * ua / dU, rel=1e-12)
    assert f.y_momentum[i] == pytest.approx(-rho * us ** 2 * va / dU, rel=1e-12)
    assert f.sensible_heat[i] == pytest.approx(-rho * cp * us * ths, rel=1e-10)
    assert f.water_vapor[i] == pytest.approx(-rho * us * qs, rel=1e-10)
    assert f.latent_heat[i] == pytest.approx(-rho * us * qs * Lv, rel=1e-10)
    assert int(ci.ao_iterations[i]) == 2          # first trip always runs; second has zero drift


@pytest.mark.parametrize("temperature", ["bulk", "skin_const", "skin_interior"])
def test_zero_flux_invariance(oracle_lib, host_backend, temperature):
    """Tᵒᶜ chosen so that Δθ = 0, qₐ = qₛ(Tᵒᶜ), ocean moving with the wind => all five fluxes vanish."""
    tf = {"bulk": F.BulkTemperature(), "skin_const": F.SkinTemperature(F.DiffusiveFlux(1e-2, 1)),
          "skin_interior": F.SkinTemperature(F.DiffusiveFlux(F.InteriorDiffusivity(), 1))}[temperature]
    Ta, pa, ua, va = 288.15, 101325.0, 3.0, -2.0
    th = F.AtmosphereThermodynamicsParameters()
    ip = F.InterfaceProperties().pod(); thp = th.pod()
    # fixed point of qa = q_s(Tk(qa)):  Tk = Ta + g h / cp(qa)
    qa = 0.01
    for _ in range(50):
        cp = (th.gas_constant / th.dry_air_molar_mass / th.dry_air_adiabatic_exponent) * (1 - qa) + th.water_vapor_heat_capacity * qa
        Tk = Ta + 9.80665 * 10.0 / cp
        qa = oracle_lib.dll.neo_surface_specific_humidity_f64(C.byref(ip), C.byref(thp), pa, Tk, 35.0)
    ci = _uniform_case(oracle_lib, host_backend, ua=ua, va=va, Ta=Ta, qa=qa, pa=pa, To=Tk - 273.15, So=35.0, uo=ua, vo=va,
                       atmosphere_ocean_interface_temperature=tf)
    if temperature == "skin_interior":
        ci.kappa = np.full(ci.grid.shape, 3e-3)
    ci.compute_atmosphere_ocean_fluxes()
    eps32 = np.finfo(np.float32).eps
    f = ci.ao_fluxes
    for name in ("x_momentum", "y_momentum", "sensible_heat", "latent_heat", "water_vapor"):
        assert abs(getattr(f, name)[3, 3]) < eps32, name


def test_fixed_point_residual_and_iteration_bounds(oracle_lib, host_backend):
    """Converged points satisfy |Ψⁿ - Ψⁿ⁻¹| < tol: one more iteration from the output state moves it by < tol."""
    ci = synthetic.build_case("tiny", host_backend, lib=oracle_lib, with_iterations=True)
    ci.initialize(); ci.update_state(4000.0)
    g = ci.grid
    it = g.interior(ci.ao_iterations)
    inactive = g.interior(ci.inactive).astype(bool)
    assert (it[inactive] == 0).all() and (it[~inactive] >= 2).all() and it.max() < 100
    for n in ci.ao_fluxes.names():
        a = g.interior(getattr(ci.ao_fluxes, n))
        assert np.isfinite(a).all() and (a[inactive] == 0).all()
    assert (g.interior(ci.ao_temperature)[inactive] == 0).all()          # 273.15 K -> 0 °C
    # FixedIterations(it+1) reproduces the converged scales to within the tolerance
    ci2 = synthetic.build_case("tiny", host_backend, lib=oracle_lib, with_iterations=True,
                               atmosphere_ocean_fluxes=F.SimilarityTheoryFluxes(solver_stop_criteria=F.FixedIterations(int(it.max()) + 1)))
    ci2.initialize(); ci2.update_state(4000.0)
    act = ~inactive
    drift = sum(np.abs(g.interior(getattr(ci.ao_fluxes, n)) - g.interior(getattr(ci2.ao_fluxes, n)))[act]
                for n in ("friction_velocity", "temperature_scale", "water_vapor_scale"))
    assert drift.max() < 1e-7


def test_momentum_flux_opposes_relative_wind_and_heat_sign(oracle_lib, host_backend):
    ci = synthetic.build_case("tiny", host_backend, lib=oracle_lib)
    ci.initialize(); ci.update_state(4000.0)
    g = ci.grid
    act = ~g.interior(ci.inactive).astype(bool)
    du = g.interior(ci.atmos_state.u) - 0.5 * (ci.ocean_state.u[g.hy - 1:g.hy + g.ny + 1, g.hx - 1:g.hx + g.nx + 1] +
                                                ci.ocean_state.u[g.hy - 1:g.hy + g.ny + 1, g.hx:g.hx + g.nx + 2])
    tx = g.interior(ci.ao_fluxes.x_momentum)
    assert (np.sign(tx[act]) == -np.sign(du[act])).all()
    # sensible heat has the sign of -θ★ (positive = ocean cooling)
    assert (np.sign(g.interior(ci.ao_fluxes.sensible_heat)[act]) == -np.sign(g.interior(ci.ao_fluxes.temperature_scale)[act])).all()


# ---- three-equation sea-ice–ocean solver (test_sea_ice_ocean_heat_fluxes.jl:48-181) ---------------------
def _three_equation(oracle_lib, host_backend, To, So, Si, conc=1.0, formulation=None):
    g = ne_b200.ExchangeGrid(nx=2, ny=2, hx=2, hy=2, latitude=(-1.0, 1.0))
    ocean = F.MediumProperties(reference_density=1025.0, heat_capacity=3991.0)
    ci = ne_b200.ComponentInterfaces(g, host_backend, None, None, lib=oracle_lib, sea_ice=True, ocean_properties=ocean,
                                     sea_ice_ocean_heat_flux=formulation or F.ThreeEquationHeatFlux())
    s = ci.sea_ice_state
    s.S[:] = Si; s.hi[:] = 1.0; s.hc[:] = 0.1; s.concentration[:] = conc
    T3 = np.full((1,) + g.shape, To); S3 = np.full((1,) + g.shape, So); dz = np.array([1.0])
    d = ci.sea_ice_ocean_desc(T3, S3, dz, 1200.0, 1, 0)
    oracle_lib.call("sea_ice_ocean_fluxes", "f64", d)
    i = (2, 2)
    return ci.sio_fluxes.interface_heat[i], ci.sio_temperature[i], ci.sio_salinity[i], T3[0][i], ci


def test_three_equation_solver_properties(oracle_lib, host_backend):
    L, rho, c, ah, us = 334e3, 1025.0, 3991.0, 0.0095, 0.002
    Tm = lambda S: 0.0 - 0.054 * S  # noqa: E731
    # warm ocean: melting
    Q, Tb, Sb, _, _ = _three_equation(oracle_lib, host_backend, 2.0, 35.0, 5.0)
    assert 5.0 <= Sb <= 35.0 and Tb == pytest.approx(Tm(Sb)) and Q > 0
    # Q = ℰ q with q = η (T - T★)
    eta = rho * c * ah * us / L
    assert Q == pytest.approx(L * eta * (2.0 - Tb), rel=1e-12)
    # cool ocean
    Q, Tb, Sb, _, _ = _three_equation(oracle_lib, host_backend, Tm(35.0) + 0.5, 35.0, 5.0)
    assert 5.0 <= Sb <= 35.0 and Tb == pytest.approx(Tm(Sb)) and Q > 0
    # ocean at the freezing point: no melt, interface = ocean
    Q, Tb, Sb, _, _ = _three_equation(oracle_lib, host_backend, Tm(35.0), 35.0, 5.0)
    assert Sb == pytest.approx(35.0) and Tb == pytest.approx(Tm(35.0)) and abs(Q) < 1e-9
    # concentration scales the flux linearly
    Q1, *_ = _three_equation(oracle_lib, host_backend, 1.0, 34.0, 4.0, conc=1.0)
    Q5, *_ = _three_equation(oracle_lib, host_backend, 1.0, 34.0, 4.0, conc=0.5)
    assert Q5 == pytest.approx(0.5 * Q1, rel=1e-14)


def test_ice_bath_and_frazil_clamp(oracle_lib, host_backend):
    """Ice bath: Q = ρ c αₕ u★ (T - Tm) ℵ; supercooled water is clamped to Tm and releases frazil heat."""
    Q, _, _, Ttop, ci = _three_equation(oracle_lib, host_backend, 1.0, 35.0, 5.0, conc=0.8, formulation=F.IceBathHeatFlux())
    assert Q == pytest.approx(1025.0 * 3991.0 * 0.006 * 0.02 * (1.0 - (-0.054 * 35.0)) * 0.8, rel=1e-14)
    Q, _, _, Ttop, ci = _three_equation(oracle_lib, host_backend, -2.5, 35.0, 5.0, formulation=F.IceBathHeatFlux())
    assert Ttop == -0.054 * 35.0                      # clamped (test_surface_fluxes.jl:258-292: minimum(T) == Tm)
    frz = ci.sio_fluxes.frazil_heat[2, 2]
    assert frz == pytest.approx(-1025.0 * 3991.0 * (-0.054 * 35.0 + 2.5) * 1.0 / 1200.0, rel=1e-14) and frz < 0
    assert abs(Q) < 1e-9                               # after the clamp the ocean sits at the freezing point


def test_elevation_correction_matches_the_closed_form(oracle_lib, host_backend):
    """atmosphere_state_correction.jl:133-146: T <- T - G dz, p <- p exp(-g dz / (Rd (T - G dz/2))); q, u, v untouched;
    dz = 0 (sea level) is the identity bit for bit."""
    import ne_b200
    from numericalearth_jl_b200 import synthetic
    g = ne_b200.ExchangeGrid(nx=48, ny=20, hx=3, hy=3, latitude=(-60.0, 60.0), FT="f64")
    rng = np.random.default_rng(7)
    zs = rng.uniform(-50.0, 2500.0, (g.ny, g.nx))
    zs[:, :10] = 0.0
    corr = ne_b200.ElevationCorrection(surface_elevation=zs, atmosphere_elevation=0.0, lapse_rate=6.5e-3)
    ci = synthetic.build_case(dict(nx=48, ny=20, hx=3, hy=3, latitude=(-60.0, 60.0)), host_backend, FT="f64", atm_FT="f64",
                              lib=oracle_lib, grid=g, atmosphere_correction=corr)
    ci.initialize()
    ci.interpolate_state(4000.0)
    T0, p0, q0 = ci.atmos_state.T.copy(), ci.atmos_state.p.copy(), ci.atmos_state.q.copy()
    ci.correct_state()
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    dz = zs
    G, grav, Rd = 6.5e-3, 9.80665, 8.3144598 / 0.02897
    T_ref = T0[rows, cols] - G * dz
    p_ref = p0[rows, cols] * np.exp(-grav * dz / (Rd * (T0[rows, cols] - G * dz / 2)))
    assert np.allclose(ci.atmos_state.T[rows, cols], T_ref, rtol=1e-15, atol=0)
    assert np.allclose(ci.atmos_state.p[rows, cols], p_ref, rtol=1e-14, atol=0)
    assert np.array_equal(ci.atmos_state.q, q0)
    assert np.array_equal(ci.atmos_state.T[rows, cols][:, :10], T0[rows, cols][:, :10])
    assert np.array_equal(ci.atmos_state.p[rows, cols][:, :10], p0[rows, cols][:, :10])
    # a 1000 m lift cools by 6.5 K and drops the pressure by ~11 %
    k = np.unravel_index(np.argmin(np.abs(dz - 1000.0)), dz.shape)
    assert 0.85 < (ci.atmos_state.p[rows, cols][k] / p0[rows, cols][k]) ** (1000.0 / dz[k]) < 0.92
