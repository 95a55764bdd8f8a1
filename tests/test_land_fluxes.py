"""the atmosphere-land flux kernel (atmosphere_land_fluxes.jl:48-251) with the land
humidity closures BulkHumidity / FractionalHumidity / SkinHumidity (interface_states.jl:92-229, 585-651).
CPU: the reference's own known answers (test/test_surface_fluxes.jl:340-423) against the oracle, and properties of the
oracle kernel.  GPU: CUDA kernel against the oracle."""
import ctypes as C

import numpy as np
import pytest

import ne_b200
from numericalearth_jl_b200 import abi as A
from numericalearth_jl_b200 import formulations as F
from numericalearth_jl_b200 import synthetic

T_STEP = 0.37 * 10800.0
CFG = dict(nx=96, ny=40, latitude=(-70.0, 70.0))


def _thermo():
    return F.AtmosphereThermodynamicsParameters(FT="f64").pod()


def _q(oracle_lib, humidity, Ts, Td, S, ustar, qstar, qprev, p=101325.0, qa=0.005, Ta=290.0):
    h, th = F.land_humidity_pod(humidity), _thermo()
    return oracle_lib.dll.neo_land_interface_humidity(C.byref(h), C.byref(th), p, qa, Ta, Ts, Td, S, ustar, qstar, qprev)


def _qsat(oracle_lib, T, p=101325.0):
    th = _thermo()
    return oracle_lib.dll.neo_saturation_specific_humidity(C.byref(th), T, p, A.NE_PHASE_LIQUID)


def test_skin_humidity_vapor_flux_balance_reference_kat(oracle_lib):
    """test/test_surface_fluxes.jl:340-391."""
    qa, Td, Ts = 0.005, 295.0, 310.0
    qv = _qsat(oracle_lib, Td)
    assert qv > qa

    def converge(d):
        sh = F.SkinHumidity(surface_thickness=d, vapor_diffusivity=2e-2)
        q = qv
        for _ in range(100):
            q = _q(oracle_lib, sh, Ts, Td, 1.0, 0.3, -1e-4, q)
        return q

    thin, mid, thick = converge(1e-3), converge(1e-1), converge(1e2)
    for q in (thin, mid, thick):
        assert qa <= q <= qv
    assert abs(thin - qv) <= 1e-2 * qv and abs(thick - qa) <= 1e-2 * qa
    assert thin > mid > thick
    # zero turbulent flux (first iterate): saturated surface; and independent of the skin temperature
    sh = F.SkinHumidity(surface_thickness=0.1, vapor_diffusivity=2e-2)
    assert abs(_q(oracle_lib, sh, Ts, Td, 1.0, 0.0, 0.0, 0.0) - qv) <= 1e-12 * qv
    assert _q(oracle_lib, sh, 280.0, Td, 1.0, 0.3, -1e-4, 0.01) == _q(oracle_lib, sh, 310.0, Td, 1.0, 0.3, -1e-4, 0.01)


def test_fractional_and_bulk_humidity_reference_kat(oracle_lib):
    """test/test_surface_fluxes.jl:393-423 (Manabe critical wetness) + BulkHumidity (interface_states.jl:120-126)."""
    Ts = 295.0
    qv = _qsat(oracle_lib, Ts)
    fh = F.FractionalHumidity(efficiency=F.CriticalSaturation(0.75))
    assert _q(oracle_lib, fh, Ts, Ts, 0.0, 0.3, 0.0, 0.0) == 0.0
    assert abs(_q(oracle_lib, fh, Ts, Ts, 0.375, 0.3, 0.0, 0.0) - 0.5 * qv) <= 1e-15
    assert abs(_q(oracle_lib, fh, Ts, Ts, 0.75, 0.3, 0.0, 0.0) - qv) <= 1e-15
    assert abs(_q(oracle_lib, fh, Ts, Ts, 1.0, 0.3, 0.0, 0.0) - qv) <= 1e-15
    fc = F.FractionalHumidity(efficiency=0.4)
    assert abs(_q(oracle_lib, fc, Ts, Ts, 0.1, 0.3, 0.0, 0.0) - 0.4 * qv) <= 1e-15
    bh = F.BulkHumidity()
    assert _q(oracle_lib, bh, Ts, Ts, 0.0, 0.3, 0.0, 0.0) == 0.0
    assert _q(oracle_lib, bh, Ts, Ts, 1e-6, 0.3, 0.0, 0.0) == qv
    with pytest.raises(ne_b200.NoKernelVariantError):
        F.land_humidity_pod(F.DryLayerHumidity(dry_layer_depth=lambda S: 0.01))   # a user closure: no kernel variant


def _dry(eta=1.0, D0=2.5e-5, width=None, tort=None):
    return F.DryLayerHumidity(
        dry_layer_depth=F.StorageBasedDryLayerDepth(maximum_dry_layer_depth=0.05, dry_layer_onset_saturation=0.5, dry_layer_exponent=eta),
        vapor_exchange=F.DryLayerVaporPistonVelocity(minimum_dry_layer_depth=1e-4, molecular_diffusivity=D0, wet_transition_width=width,
                                                     tortuosity=tort or F.ConstantTortuosity()),
        thermal_exchange_depth=0.10, porosity=0.4)


def test_dry_layer_humidity_reference_kats(oracle_lib):
    """test/test_dry_layer_humidity.jl:30-140: wet branch, vapor divider, T_e interpolation, G_e -> 0, wet-transition blend."""
    p, qa, Ta, us, qst, qprev = 1.0e5, 1.0e-2, 295.0, 0.3, -2.0e-4, 0.005
    kw = dict(p=p, qa=qa, Ta=Ta)
    # wet branch: S = S_c -> dv = 0 -> q_in = q_sat(T_in) (sharp switch)
    q = _q(oracle_lib, _dry(eta=2.0, width=0.0), 300.0, 290.0, 0.5, us, qst, qprev, **kw)
    assert abs(q - _qsat(oracle_lib, 300.0, p)) <= 1e-15
    # vapor divider, fully dry: dv = dv_max = 0.05, chi = 0.5, T_e = (T_in + T_la) / 2
    q = _q(oracle_lib, _dry(), 300.0, 290.0, 0.0, us, qst, qprev, **kw)
    th = _thermo()
    Rd, Rv = th.gas_constant / th.dry_air_molar_mass, th.gas_constant / th.water_molar_mass
    rho = p / ((Rd * (1 - qa) + Rv * qa) * Ta)
    qe = _qsat(oracle_lib, 295.0, p)
    Ge, Ja, dq = rho * 2.5e-5 / 0.05, -rho * us * qst, qprev - qa
    expected = (Ge * qe + Ja / dq * qa) / (Ge + Ja / dq)
    # the logistic weight at dv = 0.05, dv_min = 1e-4, width 5e-4 is 1 to machine precision
    assert abs(q - expected) <= 1e-15
    # T_e interpolation: dry source is colder than the skin -> q_dry < q_wet = q_sat(T_in)
    qd = _q(oracle_lib, _dry(width=0.0), 310.0, 290.0, 0.0, us, qst, qprev, **kw)
    qw = _q(oracle_lib, _dry(width=0.0), 310.0, 290.0, 0.5, us, qst, qprev, **kw)
    assert abs(qw - _qsat(oracle_lib, 310.0, p)) <= 1e-15 and qd < qw
    # G_e -> 0: the atmospheric flux drives q_in to q_at
    q = _q(oracle_lib, _dry(D0=1e-14), 300.0, 290.0, 0.0, us, qst, qprev, **kw)
    assert abs(q - qa) <= 1e-6
    # wet-transition blend: monotone in dv between the saturated skin and the series solution, sigma = 1/2 at the centre
    w = 5e-3
    centre_S = 0.5 * (1 - (1e-4 + w / 2) / 0.05)          # eta = 1: dv = dv_max (1 - S / S_c)
    q_mid = _q(oracle_lib, _dry(width=w), 300.0, 290.0, centre_S, us, qst, qprev, **kw)
    q_sat = _qsat(oracle_lib, 300.0, p)
    q_ser = _q(oracle_lib, _dry(width=0.0), 300.0, 290.0, centre_S, us, qst, qprev, **kw)
    assert abs(q_mid - 0.5 * (q_sat + q_ser)) <= 1e-9
    # Millington-Quirk tortuosity lowers the diffusivity of a moist soil: closer to q_at than with constant tortuosity
    qc = _q(oracle_lib, _dry(), 300.0, 290.0, 0.2, us, qst, qprev, **kw)
    qp = _q(oracle_lib, _dry(tort=F.PowerLawTortuosity()), 300.0, 290.0, 0.2, us, qst, qprev, **kw)
    assert abs(qp - qa) < abs(qc - qa)


HUMIDITIES = {
    "bulk": lambda: F.BulkHumidity(),
    "fractional_critical": lambda: F.FractionalHumidity(efficiency=F.CriticalSaturation(0.75)),
    "fractional_constant": lambda: F.FractionalHumidity(efficiency=0.4),
    "skin": lambda: F.SkinHumidity(surface_thickness=0.05, vapor_diffusivity=2e-2),
    "dry_layer": lambda: F.DryLayerHumidity(
        dry_layer_depth=F.StorageBasedDryLayerDepth(maximum_dry_layer_depth=0.05, dry_layer_onset_saturation=0.6, dry_layer_exponent=2.0),
        vapor_exchange=F.DryLayerVaporPistonVelocity(minimum_dry_layer_depth=1e-4, molecular_diffusivity=2.5e-5,
                                                     tortuosity=F.PowerLawTortuosity()),
        thermal_exchange_depth=0.10, porosity=0.4),
}


def _land_state(grid, backend, FT):
    rng = np.random.default_rng(77)
    npd = np.float64 if FT == "f64" else np.float32
    phi = np.deg2rad(grid.phi.astype(np.float64))[:, None] * np.ones((1, grid.shape[1]))
    T = (300.0 - 35.0 * np.sin(phi) ** 2 + rng.normal(0, 3.0, grid.shape)).astype(npd)
    S = np.clip(rng.uniform(-0.2, 1.1, grid.shape), 0.0, 1.0).astype(npd)      # some cells completely dry, some saturated
    return ne_b200.SlabLandState(T=backend.from_numpy(T), saturation=backend.from_numpy(S))


def _case(backend, lib, FT, atm_FT, humidity, **kw):
    ci = synthetic.build_case(CFG, backend, FT=FT, atm_FT=atm_FT, lib=lib, with_iterations=True,
                              atmosphere_land_interface_specific_humidity=humidity, **kw)
    ci.slab_land = _land_state(ci.grid, backend, FT)
    # the land exchanger state is attached after construction: allocate the land flux fields the same way the constructor does
    Z = lambda: backend.zeros(ci.grid.shape, FT)  # noqa: E731
    from numericalearth_jl_b200.interface import _Fields
    ci.al_fluxes = _Fields(latent_heat=Z(), sensible_heat=Z(), water_vapor=Z(), x_momentum=Z(), y_momentum=Z(),
                           friction_velocity=Z(), temperature_scale=Z(), water_vapor_scale=Z())
    ci.al_temperature = Z()
    ci.al_iterations = backend.zeros(ci.grid.shape, "i32")
    ci.land_surface_energy_flux = Z()
    ci.rad_fluxes_land = _Fields(upwelling_longwave=Z(), downwelling_longwave=Z(), downwelling_shortwave=Z())
    ci.radiation.surface_properties["land"] = F.SurfaceRadiationProperties(0.23, 0.95)
    ci.initialize()
    ci.interpolate_state(T_STEP)
    ci.compute_atmosphere_land_fluxes()
    ci.land_surface_energy_flux[...] = 3.5          # READ-MODIFY-WRITE: the kernel adds to what is there
    ci.apply_air_land_radiative_fluxes()
    return ci


@pytest.mark.parametrize("name", sorted(HUMIDITIES))
def test_land_flux_kernel_oracle_properties(oracle_lib, name):
    ci = _case(ne_b200.NumpyHostBackend(), oracle_lib, "f64", "f64", HUMIDITIES[name]())
    g = ci.grid
    inner = (slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx))
    f = ci.al_fluxes
    it = ci.al_iterations[inner]
    assert it.min() >= 1 and it.max() <= 100
    us = f.friction_velocity[inner]
    conv = it < 100
    assert np.isfinite(us[conv]).all() and (us[conv] > 0).all()
    assert np.array_equal(ci.al_temperature[inner], np.asarray(ci.slab_land.T)[inner])       # BulkTemperature: T_s is the land temperature
    # momentum flux opposes the wind (surface at rest), Q_c = -rho c_p u* theta*, Q_v = L J_v with L > 0
    ua, va = ci.atmos_state.u[inner], ci.atmos_state.v[inner]
    assert (np.sign(f.x_momentum[inner][conv]) == -np.sign(ua[conv])).all() and (np.sign(f.y_momentum[inner][conv]) == -np.sign(va[conv])).all()
    Jv, Qv = f.water_vapor[inner][conv], f.latent_heat[inner][conv]
    assert (np.sign(Jv) == np.sign(Qv)).all()
    if name == "bulk":   # completely dry cells (S = 0) have q_s = 0: vapour can only flow downward (dew), J_v <= 0
        dry = (np.asarray(ci.slab_land.saturation)[inner] == 0) & conv
        assert dry.sum() > 50 and (f.water_vapor[inner][dry] <= 0).all()
    # the halo / ring outside `:xy` is not written
    assert (f.latent_heat[: g.hy, :] == 0).all() and (f.latent_heat[g.hy - 1, :] == 0).all()
    # apply_air_land_radiative_fluxes! (apply_air_land_radiative_fluxes.jl:64-97): Q += σ ε Tₛ⁴ − ε ℐ↓ˡʷ − (1 − α) ℐ↓ˢʷ at active
    # cells (positive upward), the three diagnostic fluxes everywhere
    sig, alpha, eps = ci.radiation.stefan_boltzmann_constant, 0.23, 0.95
    Ts = ci.al_temperature[inner]
    sw, lw = ci.rad_state.sw[inner], ci.rad_state.lw[inner]
    up = sig * eps * Ts ** 4
    act = np.asarray(ci.inactive)[inner] == 0
    expect = 3.5 + np.where(act, up - eps * lw - (1 - alpha) * sw, 0.0)
    assert np.abs(ci.land_surface_energy_flux[inner] - expect).max() <= 1e-12 * np.abs(expect).max()
    r = ci.rad_fluxes_land
    assert np.abs(r.upwelling_longwave[inner] - up).max() <= 1e-12 * up.max()
    assert np.abs(r.downwelling_longwave[inner] - eps * lw).max() <= 1e-12 * lw.max()
    assert np.abs(r.downwelling_shortwave[inner] - (1 - alpha) * sw).max() <= 1e-12 * max(sw.max(), 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("FT,atm_FT", [("f64", "f64"), ("f64", "f32"), ("f32", "f32")])
@pytest.mark.parametrize("name", sorted(HUMIDITIES))
def test_cuda_land_flux_kernel_parity(oracle_lib, cuda_backend, cuda_lib, monkeypatch, name, FT, atm_FT):
    """Float64 models take the work-queue kernel with the tabulated Large-Yeager functions (ne_flux_land_fast.cuh), Float32
    models the generic kernel; both against the oracle, and the Float64 fast path against the generic kernel."""
    ref = _case(ne_b200.NumpyHostBackend(), oracle_lib, FT, atm_FT, HUMIDITIES[name]())
    dev = _case(cuda_backend, None, FT, atm_FT, HUMIDITIES[name]())
    cuda_backend.synchronize()
    if FT == "f64":
        monkeypatch.setenv("NE_B200_FORCE_GENERIC", "1")
        gen = _case(cuda_backend, None, FT, atm_FT, HUMIDITIES[name]())
        cuda_backend.synchronize()
        monkeypatch.delenv("NE_B200_FORCE_GENERIC")
        gi = ref.grid
        win = (slice(gi.hy, gi.hy + gi.ny), slice(gi.hx, gi.hx + gi.nx))
        both = (cuda_backend.to_numpy(dev.al_iterations)[win] < 100) & (cuda_backend.to_numpy(gen.al_iterations)[win] < 100)
        differs = False
        for n in dev.al_fluxes.names():
            a, b = cuda_backend.to_numpy(getattr(dev.al_fluxes, n))[win], cuda_backend.to_numpy(getattr(gen.al_fluxes, n))[win]
            differs |= bool((a != b).any())
            sc = float(np.nanmax(np.abs(b))) or 1.0
            assert np.nanmax(np.abs(a - b)[both]) / sc <= 1e-10, f"fast vs generic {name}/{n}: {np.nanmax(np.abs(a - b)[both]) / sc}"
        assert differs, "the land fast path did not run (results identical to the generic kernel's)"
    g = ref.grid
    inner = (slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx))
    ri, di = ref.al_iterations[inner], cuda_backend.to_numpy(dev.al_iterations)[inner]
    tol = 1e-10 if (FT, atm_FT) == ("f64", "f64") else (2e-6 if FT == "f64" else 1e-5)
    if FT == "f64":
        assert float((ri != di).mean()) <= 2e-3
    conv = (ri < 100) & (di < 100)
    assert conv.mean() > 0.5
    for n in ref.al_fluxes.names():
        a = np.asarray(getattr(ref.al_fluxes, n))[inner].astype(np.float64)
        b = cuda_backend.to_numpy(getattr(dev.al_fluxes, n))[inner].astype(np.float64)
        assert np.array_equal(np.isnan(a), np.isnan(b)), n
        s = float(np.nanmax(np.abs(a))) or 1.0
        assert np.nanmax(np.abs(a - b)[conv]) / s <= tol, f"{name}/{n}: {np.nanmax(np.abs(a - b)[conv]) / s}"
    assert np.array_equal(ref.al_temperature[inner], cuda_backend.to_numpy(dev.al_temperature)[inner])
    # land radiation kernel: same arithmetic (-fmad=false translation unit): bit-exact
    assert np.array_equal(ref.land_surface_energy_flux[inner], cuda_backend.to_numpy(dev.land_surface_energy_flux)[inner])
    for n in ref.rad_fluxes_land.names():
        assert np.array_equal(getattr(ref.rad_fluxes_land, n)[inner], cuda_backend.to_numpy(getattr(dev.rad_fluxes_land, n))[inner]), n
    # an unknown closure is refused by the library itself
    d = dev.atmosphere_land_desc()
    d.humidity.kind = 9
    with pytest.raises(ne_b200.NoKernelVariantError):
        cuda_lib.call("atmosphere_land_fluxes", FT, d, cuda_backend.stream())


@pytest.mark.gpu
def test_cuda_land_example_configuration(oracle_lib, cuda_backend, cuda_lib):
    """The configuration of the reference's own land example (examples/era5_forced_slab_land.jl:164-193): DryLayerHumidity,
    default land fluxes with FixedIterations(8), BulkTemperature — work-queue kernel against the oracle, exactly 8 trips."""
    kw = dict(atmosphere_land_fluxes=F.default_atmosphere_land_fluxes(solver_stop_criteria=F.FixedIterations(8)))
    ref = _case(ne_b200.NumpyHostBackend(), oracle_lib, "f64", "f32", HUMIDITIES["dry_layer"](), **kw)
    dev = _case(cuda_backend, None, "f64", "f32", HUMIDITIES["dry_layer"](), **kw)
    cuda_backend.synchronize()
    g = ref.grid
    inner = (slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx))
    assert (ref.al_iterations[inner] == 8).all() and (cuda_backend.to_numpy(dev.al_iterations)[inner] == 8).all()
    for n in ref.al_fluxes.names():
        a = np.asarray(getattr(ref.al_fluxes, n))[inner]
        b = cuda_backend.to_numpy(getattr(dev.al_fluxes, n))[inner]
        s = float(np.abs(a).max()) or 1.0
        assert np.isfinite(b).all() and np.abs(a - b).max() / s <= 2e-6, f"{n}: {np.abs(a - b).max() / s}"   # Float32 q_sat


# ------------------------------------------------------------------------------ per-cell land roughness and displacement
def _neutral_fluxes(ell=0.1, displacement=0, scalar=None):
    """The closure of the reference's displacement tests (test/test_slab_land.jl:483-490, 523-529): no stability correction,
    no gustiness."""
    s = ell if scalar is None else scalar
    return F.SimilarityTheoryFluxes(momentum_roughness_length=ell, temperature_roughness_length=s, water_vapor_roughness_length=s,
                                    zero_plane_displacement=displacement, subgrid_velocities=None, stability_functions=None)


def _column(backend, lib, FT, fluxes, wind=5.0, land_T=288.0, fields=None, humidity=None):
    """Every cell of the grid is the reference's single column: uniform atmosphere (288 K, q = 0.003, 101325 Pa, u = wind),
    land at land_T, DryLand (saturation 0)."""
    ci = synthetic.build_case(CFG, backend, FT=FT, atm_FT=FT, lib=lib, with_iterations=True, atmosphere_land_fluxes=fluxes,
                              atmosphere_land_interface_specific_humidity=humidity or F.BulkHumidity())
    g = ci.grid
    npd = np.float64 if FT == "f64" else np.float32
    full = lambda v: backend.from_numpy(np.full(g.shape, v, npd))   # noqa: E731
    ci.slab_land = ne_b200.SlabLandState(T=full(land_T), saturation=0.0, **{k: backend.from_numpy(np.asarray(v, npd)) for k, v in (fields or {}).items()})
    from numericalearth_jl_b200.interface import _Fields
    Z = lambda: backend.zeros(g.shape, FT)  # noqa: E731
    ci.al_fluxes = _Fields(latent_heat=Z(), sensible_heat=Z(), water_vapor=Z(), x_momentum=Z(), y_momentum=Z(),
                           friction_velocity=Z(), temperature_scale=Z(), water_vapor_scale=Z())
    ci.al_temperature = Z()
    ci.al_iterations = backend.zeros(g.shape, "i32")
    ci.initialize()
    ci.interpolate_state(T_STEP)
    a = ci.atmos_state
    a.u, a.v, a.T, a.q, a.p = full(wind), full(0.0), full(288.0), full(0.003), full(101325.0)
    ci.compute_atmosphere_land_fluxes()
    return ci


def _ustar(ci, backend=None):
    g = ci.grid
    v = ci.al_fluxes.friction_velocity
    v = backend.to_numpy(v) if backend is not None else np.asarray(v)
    return v[g.hy:g.hy + g.ny, g.hx:g.hx + g.nx]


def test_zero_plane_displacement_known_answers_of_the_reference(oracle_lib):
    """test/test_slab_land.jl:463-508: the similarity profile is evaluated at h - d, floored at twice the roughness length."""
    host = ne_b200.NumpyHostBackend()
    ell, wind, kappa = 0.1, 5.0, 0.4

    def ustar(d):
        ci = _column(host, oracle_lib, "f64", _neutral_fluxes(ell, d), wind)
        u = _ustar(ci)
        assert (u == u[0, 0]).all()
        return float(u[0, 0]), float(ci.atmosphere.surface_layer_height)

    u4, h = ustar(4.0)
    assert u4 == pytest.approx(kappa / np.log((h - 4) / ell) * wind, rel=1e-12)
    assert ustar(0.0)[0] == pytest.approx(kappa / np.log(h / ell) * wind, rel=1e-12)
    assert ustar(6.0)[0] > ustar(3.0)[0] > ustar(0.0)[0]                 # displacement thins the layer, raising the drag
    assert ustar(2 * h)[0] == pytest.approx(kappa / np.log(2.0) * wind, rel=1e-12)   # floor at 2 l: finite


def test_land_markers_resolve_per_cell_like_constants(oracle_lib):
    """test/test_slab_land.jl:510-549 (LandZeroPlaneDisplacement) and local_roughness_length(::LandRoughnessLength)
    (similarity_theory_turbulent_fluxes.jl:265-278): a per-cell field reaches the solver as the constant would, bit for bit."""
    host = ne_b200.NumpyHostBackend()
    g = synthetic.build_case(CFG, host, lib=oracle_lib).grid
    # LandZeroPlaneDisplacement with a field of 4 m == the constant 4.0; without a field == 0
    per_cell = _ustar(_column(host, oracle_lib, "f64", _neutral_fluxes(0.1, F.LandZeroPlaneDisplacement()),
                              fields={"zero_plane_displacement": np.full(g.shape, 4.0)}))
    assert np.array_equal(per_cell, _ustar(_column(host, oracle_lib, "f64", _neutral_fluxes(0.1, 4.0))))
    assert per_cell[0, 0] == pytest.approx(0.4 / np.log((10.0 - 4) / 0.1) * 5.0, rel=1e-12)
    assert np.array_equal(_ustar(_column(host, oracle_lib, "f64", _neutral_fluxes(0.1, F.LandZeroPlaneDisplacement()))),
                          _ustar(_column(host, oracle_lib, "f64", _neutral_fluxes(0.1, 0.0))))
    # LandRoughnessLength: max(multiplier * max(field, minimum), minimum); the scalar slots read the scalar field
    lm = F.LandRoughnessLength(multiplier=1, minimum_roughness_length=1e-3)
    ls = F.LandRoughnessLength(multiplier=0.1, minimum_roughness_length=1e-4)
    fl = F.SimilarityTheoryFluxes(momentum_roughness_length=lm, temperature_roughness_length=ls, water_vapor_roughness_length=ls,
                                  stability_functions=F.atmosphere_land_stability_functions())
    z0m, z0s = np.full(g.shape, 0.25), np.full(g.shape, 0.05)
    z0m[:, ::2] = 1e-5                                                    # below the minimum: floored at 1e-3
    got = _column(host, oracle_lib, "f64", fl, land_T=294.0, fields={"momentum_roughness_length": z0m, "scalar_roughness_length": z0s})
    for cols, lu in ((slice(1, None, 2), 0.25), (slice(0, None, 2), 1e-3)):
        lsc = max(0.1 * 0.05, 1e-4)
        want = _column(host, oracle_lib, "f64", F.SimilarityTheoryFluxes(
            momentum_roughness_length=lu, temperature_roughness_length=lsc, water_vapor_roughness_length=lsc,
            stability_functions=F.atmosphere_land_stability_functions()), land_T=294.0)
        hx = g.hx % 2
        sel = slice((cols.start + hx) % 2, None, 2)     # interior column parity after removing the halo
        for n in got.al_fluxes.names():
            a = np.asarray(getattr(got.al_fluxes, n))[g.hy:g.hy + g.ny, g.hx:g.hx + g.nx]
            b = np.asarray(getattr(want.al_fluxes, n))[g.hy:g.hy + g.ny, g.hx:g.hx + g.nx]
            assert np.array_equal(a[:, sel], b[:, sel]), (n, lu)
    # a land model that provides no roughness field (SlabLand): the floor stands in, max(multiplier * minimum, minimum)
    bare = _column(host, oracle_lib, "f64", fl, land_T=294.0)
    want = _column(host, oracle_lib, "f64", F.SimilarityTheoryFluxes(
        momentum_roughness_length=1e-3, temperature_roughness_length=1e-4, water_vapor_roughness_length=1e-4,
        stability_functions=F.atmosphere_land_stability_functions()), land_T=294.0)
    assert np.array_equal(_ustar(bare), _ustar(want))
    # larger roughness raises the drag (test/test_slab_land.jl:617-618)
    assert _ustar(_column(host, oracle_lib, "f64", _neutral_fluxes(0.5)))[0, 0] > _ustar(_column(host, oracle_lib, "f64", _neutral_fluxes(0.05)))[0, 0]


def test_land_markers_collapse_to_constants_over_ocean_and_sea_ice():
    """Over the ocean `interior_properties` has no land fields: LandRoughnessLength returns max(multiplier * minimum, minimum),
    LandZeroPlaneDisplacement 0 (similarity_theory_turbulent_fluxes.jl:265-303)."""
    fl = F.SimilarityTheoryFluxes(momentum_roughness_length=F.LandRoughnessLength(multiplier=3, minimum_roughness_length=1e-3),
                                  temperature_roughness_length=F.LandRoughnessLength(multiplier=0.1, minimum_roughness_length=1e-4),
                                  water_vapor_roughness_length=1e-4, zero_plane_displacement=F.LandZeroPlaneDisplacement())
    land = F.flux_formulation_pod(fl, land=True)
    assert land.ell_momentum.kind == A.NE_ROUGH_LAND and land.zero_plane_displacement_kind == A.NE_DISPLACEMENT_LAND
    sea = F.flux_formulation_pod(fl)
    assert sea.ell_momentum.kind == A.NE_ROUGH_CONSTANT and sea.ell_momentum.constant == 3e-3
    assert sea.ell_temperature.kind == A.NE_ROUGH_CONSTANT and sea.ell_temperature.constant == 1e-4
    assert sea.zero_plane_displacement_kind == A.NE_DISPLACEMENT_CONSTANT and sea.zero_plane_displacement == 0.0
    eps32 = F.flux_formulation_pod(F.SimilarityTheoryFluxes(momentum_roughness_length=F.LandRoughnessLength(FT="f32")), FT="f32")
    assert eps32.ell_momentum.constant == float(np.finfo(np.float32).eps)      # minimum_roughness_length = eps(FT)


@pytest.mark.gpu
@pytest.mark.parametrize("FT", ["f64", "f32"])
def test_cuda_land_kernel_with_per_cell_roughness_and_displacement(oracle_lib, cuda_backend, cuda_lib, FT):
    """LandRoughnessLength + LandZeroPlaneDisplacement with random per-cell fields: generic CUDA kernel against the oracle;
    the displacement marker alone (constant roughness) must leave the work-queue fast path for the generic kernel."""
    host = ne_b200.NumpyHostBackend()
    g = synthetic.build_case(CFG, host, lib=oracle_lib).grid
    rng = np.random.default_rng(404)
    fields = {"momentum_roughness_length": 10.0 ** rng.uniform(-5, 0.3, g.shape), "scalar_roughness_length": 10.0 ** rng.uniform(-6, -1, g.shape),
              "zero_plane_displacement": np.where(rng.random(g.shape) < 0.15, rng.uniform(8.0, 25.0, g.shape), rng.uniform(0.0, 6.0, g.shape))}
    lm, ls = F.LandRoughnessLength(1, 1e-4, FT), F.LandRoughnessLength(0.1, 1e-5, FT)
    trees = {
        "markers": F.SimilarityTheoryFluxes(momentum_roughness_length=lm, temperature_roughness_length=ls, water_vapor_roughness_length=ls,
                                            zero_plane_displacement=F.LandZeroPlaneDisplacement(),
                                            stability_functions=F.atmosphere_land_stability_functions()),
        "displacement_only": F.SimilarityTheoryFluxes(momentum_roughness_length=0.1, temperature_roughness_length=0.01, water_vapor_roughness_length=0.01,
                                                      zero_plane_displacement=F.LandZeroPlaneDisplacement(),
                                                      stability_functions=F.atmosphere_land_stability_functions()),
    }
    tol = 1e-10 if FT == "f64" else 1e-5
    inner = (slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx))
    for name, fl in trees.items():
        ref = _column(host, oracle_lib, FT, fl, land_T=292.0, fields=fields, humidity=F.FractionalHumidity(efficiency=0.6))
        dev = _column(cuda_backend, None, FT, fl, land_T=292.0, fields=fields, humidity=F.FractionalHumidity(efficiency=0.6))
        cuda_backend.synchronize()
        ri, di = ref.al_iterations[inner], cuda_backend.to_numpy(dev.al_iterations)[inner]
        conv = (ri < 100) & (di < 100)
        assert conv.mean() > (0.9 if FT == "f64" else 0.8)     # Float32 models: one-ulp limit cycles run to maxiter
        if FT == "f64":
            assert float((ri != di).mean()) <= 2e-3
        us = _ustar(ref)
        assert np.unique(us[conv]).size > 1000, "the per-cell fields did not reach the solver"
        for n in ref.al_fluxes.names():
            a = np.asarray(getattr(ref.al_fluxes, n))[inner].astype(np.float64)
            b = cuda_backend.to_numpy(getattr(dev.al_fluxes, n))[inner].astype(np.float64)
            s = float(np.nanmax(np.abs(a))) or 1.0
            assert np.nanmax(np.abs(a - b)[conv]) / s <= tol, f"{name}/{n}: {np.nanmax(np.abs(a - b)[conv]) / s}"
    # the ocean entry point refuses the markers: the binding passes the collapsed constants there
    d = dev.atmosphere_ocean_desc()
    d.flux = F.flux_formulation_pod(trees["markers"], land=True)
    with pytest.raises(ne_b200.NoKernelVariantError):
        cuda_lib.call("atmosphere_ocean_fluxes", FT, d, cuda_backend.stream())
    # SkinTemperature over land has no method in the reference either (no heat capacity in the land properties)
    d = dev.atmosphere_land_desc()
    d.properties.temperature_formulation = A.NE_TEMP_SKIN_DIFFUSIVE
    with pytest.raises(ne_b200.NoKernelVariantError):
        cuda_lib.call("atmosphere_land_fluxes", FT, d, cuda_backend.stream())
