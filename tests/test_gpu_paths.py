"""GPU parity of kernel paths that round 1 built but never ran on the device ("missing" item 3), each
against the oracle under the pointwise criterion of parity.pointwise_rel:

  _freeze_ocean_temperature! (OceanOnlyModel default)   src/SeaIces/freezing_limited_ocean_temperature.jl:73-118
  MomentumBasedFrictionVelocity                         …/InterfaceComputations/friction_velocity.jl:24-44
  IceBathHeatFlux                                       …/sea_ice_ocean_heat_flux_formulations.jl:176-195
  SkinTemperature(DiffusiveFlux(InteriorDiffusivity))   …/interface_states.jl:384-391, 434-457
  LatitudeDependentAlbedo                               src/Radiations/latitude_dependent_albedo.jl:48-53
  TwoColorRadiation shortwave routing                   src/Oceans/radiative_forcing.jl:84-91
  barotropic potential = p / rho_ocean                  src/Atmospheres/interpolate_atmospheric_state.jl:80-85
  TabulatedAlbedo through the host pipeline at t > 0    (the pipeline froze the solar geometry at t = 0)
"""
import numpy as np
import pytest

import ne_b200
from numericalearth_jl_b200 import synthetic
from parity import ParityLog, build_pair, compare_pointwise, converged_mask, pointwise_rel

pytestmark = pytest.mark.gpu

T_STEP = 0.37 * 10800.0
TOL = 1e-10


def _assert_bag(tag, ref, dev, backend, name, ring, mask=None, tol=TOL):
    """Pointwise criterion; the only points allowed to miss it are a handful next to a sign change of the field, bounded by
    1e-8 (Float64 conditioning of Δq = qₐ − qₛ and Δθ: see tests/test_gpu_parity_full.py)."""
    floor = 1e-4 if name.startswith("net_") else 1e-6     # assembled sums: see ASSEMBLED_FLOOR in tests/test_gpu_parity_full.py
    for n, r in compare_pointwise(getattr(ref, name), getattr(dev, name), ref.grid, backend, with_halo_ring=ring, mask=mask, tol=tol, floor=floor).items():
        ParityLog.add(tag, bag=name, field=n, max_pointwise_rel=r["pw"], points=r["n"], exceed_tol=r["exceed"],
                      exceed_tol_near_sign_change=r["exceed_near_zero"], tol=tol)
        assert r["exceed"] == r["exceed_near_zero"] <= 5 and (r["exceed"] == 0 or r["pw"] <= 1e-8), f"{tag}: {name}.{n}: {r['pw']:.3e} ({r['exceed']} points)"


@pytest.mark.parametrize("entry", ["update_state", "fused_interface_step", "host_pipeline"])
@pytest.mark.parametrize("first_iteration", [False, True])
def test_freezing_limited_ocean_temperature(oracle_lib, cuda_backend, cuda_lib, entry, first_iteration):
    """OceanOnlyModel: no sea-ice component, so compute_sea_ice_ocean_fluxes! is the FreezingLimitedOceanTemperature
    clamp.  Reached from all three step entry points; dt = Inf at iteration 0 (:88) makes the frazil heat exactly -0."""
    import torch
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT="f64")
    ref.initialize(); dev.initialize()
    host = ne_b200.NumpyHostBackend()
    dt = float("inf") if first_iteration else 1200.0
    colr = synthetic.ocean_column(ref.grid, host, nz=10) + (dt, 10, 0)
    cold = synthetic.ocean_column(dev.grid, cuda_backend, nz=10) + (dt, 10, 0)
    colr[0][...] -= 0.5                                      # supercool the polar rows
    cold[0].sub_(0.5)
    T_before = colr[0].copy()
    ref.update_state(T_STEP, ocean_column=colr)
    if entry == "update_state":
        dev.update_state(T_STEP, ocean_column=cold)
    elif entry == "fused_interface_step":
        dev.fused_interface_step(T_STEP, ocean_column=cold)
    else:
        pinned = {k: torch.from_numpy(np.ascontiguousarray(dev._host_inputs["ocean"][k])).pin_memory() for k in ("T", "S", "u", "v")}
        ne_b200.HostPipelinedStep(dev, n_chunks=3).step(T_STEP, pinned, ocean_column=cold)
    cuda_backend.synchronize()
    g = ref.grid
    Tr, Td = colr[0], cuda_backend.to_numpy(cold[0])
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    assert np.array_equal(Tr, Td), "clamped T column differs"
    clamped = int((Tr[:, rows, cols] != T_before[:, rows, cols]).sum())
    assert clamped > 300, "the synthetic column should be supercooled somewhere"
    Sr = colr[1][:, rows, cols]
    assert (Tr[:, rows, cols] >= -0.054 * Sr - 1e-12).all()            # never below the liquidus (LinearLiquidus defaults)
    qr = ref.frazil_heat[rows, cols]
    qd = cuda_backend.to_numpy(dev.frazil_heat)[rows, cols]
    pw = pointwise_rel(qd, qr)
    ParityLog.add(f"freeze_only_{entry}_{'dtInf' if first_iteration else 'dt1200'}", field="frazil_heat", max_pointwise_rel=pw,
                  clamped_cells=clamped)
    if first_iteration:
        assert (qr == 0).all() and (qd == 0).all()
    else:
        assert (qr < 0).any() and pw <= TOL
    _assert_bag(f"freeze_only_{entry}", ref, dev, cuda_backend, "net_ocean", False)   # the rest of the step is untouched


@pytest.mark.parametrize("heat_flux", ["three_equation_momentum_ustar", "ice_bath", "ice_bath_momentum_ustar"])
def test_sea_ice_ocean_heat_flux_variants(oracle_lib, cuda_backend, cuda_lib, heat_flux):
    ff = {"three_equation_momentum_ustar": lambda: ne_b200.ThreeEquationHeatFlux(friction_velocity=ne_b200.MomentumBasedFrictionVelocity()),
          "ice_bath": lambda: ne_b200.IceBathHeatFlux(),
          "ice_bath_momentum_ustar": lambda: ne_b200.IceBathHeatFlux(friction_velocity=ne_b200.MomentumBasedFrictionVelocity())}[heat_flux]
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT="f64", sea_ice=True, sea_ice_ocean_heat_flux=ff())
    ref.initialize(); dev.initialize()
    # sea-ice–ocean stresses the momentum-based u★ averages to cell centres (friction_velocity.jl:26-27)
    rng = np.random.default_rng(17)
    tx, ty = 0.05 * rng.standard_normal(ref.grid.shape), 0.05 * rng.standard_normal(ref.grid.shape)
    ref.sio_fluxes.x_momentum[...] = tx; ref.sio_fluxes.y_momentum[...] = ty
    dev.sio_fluxes.x_momentum.copy_(cuda_backend.from_numpy(tx)); dev.sio_fluxes.y_momentum.copy_(cuda_backend.from_numpy(ty))
    host = ne_b200.NumpyHostBackend()
    colr = synthetic.ocean_column(ref.grid, host, nz=10) + (1200.0, 10, 0)
    cold = synthetic.ocean_column(dev.grid, cuda_backend, nz=10) + (1200.0, 10, 0)
    ref.update_state(T_STEP, ocean_column=colr); dev.update_state(T_STEP, ocean_column=cold)
    cuda_backend.synchronize()
    names = ["frazil_heat", "interface_heat", "salt", "freshwater"]
    for n, r in compare_pointwise(ref.sio_fluxes, dev.sio_fluxes, ref.grid, cuda_backend, names=names, with_halo_ring=False).items():
        ParityLog.add(f"sio_{heat_flux}", field=n, max_pointwise_rel=r["pw"], points=r["n"], exceed_tol=r["exceed"])
        assert r["exceed"] == r["exceed_near_zero"] <= 5 and (r["exceed"] == 0 or r["pw"] <= 1e-8), f"{heat_flux}: {n}: {r['pw']:.3e}"
    g = ref.grid
    q = g.interior(ref.sio_fluxes.interface_heat)
    assert np.isfinite(q).all() and (q != 0).sum() > 100
    assert np.array_equal(colr[0], cuda_backend.to_numpy(cold[0]))


def test_skin_temperature_interior_diffusivity(oracle_lib, cuda_backend, cuda_lib):
    """SkinTemperature(DiffusiveFlux(κ = InteriorDiffusivity)): κ read from the ocean's diffusivity field, floored at
    minimum_diffusivity (interface_states.jl:384-391)."""
    tf = lambda: ne_b200.SkinTemperature(ne_b200.DiffusiveFlux(ne_b200.InteriorDiffusivity(), 1.0))   # noqa: E731
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT="f64", atmosphere_ocean_interface_temperature=tf())
    rng = np.random.default_rng(23)
    kappa = 10.0 ** rng.uniform(-8.0, -2.0, ref.grid.shape)          # straddles the 1.4e-7 floor
    ref.kappa = kappa.copy()
    dev.kappa = cuda_backend.from_numpy(kappa)
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP); dev.update_state(T_STEP)
    cuda_backend.synchronize()
    m = converged_mask(ref.ao_iterations, ref.grid, 100, dilate=False)
    _assert_bag("skin_interior_diffusivity", ref, dev, cuda_backend, "ao_fluxes", True, mask=m)
    Tr, Td = ref.grid.interior(ref.ao_temperature), ref.grid.interior(cuda_backend.to_numpy(dev.ao_temperature))
    assert np.nanmax(np.abs(Tr - Td)[m]) <= 1e-9
    To = ref.grid.interior(ref.ocean_state.T)
    act = ref.grid.interior(ref.inactive) == 0
    assert float(np.abs(Tr - To)[act].max()) > 1e-3, "the skin temperature should differ from the bulk one"
    assert float(np.abs(Tr - To)[act].max()) <= 5.0 + 1e-12          # max_ΔT clamp (:452-456)


def test_latitude_dependent_albedo(oracle_lib, cuda_backend, cuda_lib):
    """α(φ) = α_diffuse − α_direct cos(2φ) (latitude_dependent_albedo.jl:48-53) in the a–o kernel's radiation state and in
    the radiative flux application."""
    outs = []
    for lib, backend in ((oracle_lib, ne_b200.NumpyHostBackend()), (cuda_lib, cuda_backend)):
        ci = synthetic.build_case("C1", backend, FT="f64", atm_FT="f64", lib=lib)
        ci.radiation.surface_properties["ocean"] = ne_b200.SurfaceRadiationProperties(ne_b200.LatitudeDependentAlbedo(), 0.97)
        ci.initialize()
        ci.update_state(T_STEP)
        backend.synchronize()
        outs.append(ci)
    ref, dev = outs
    _assert_bag("latitude_dependent_albedo", ref, dev, cuda_backend, "rad_fluxes_ocean", False)
    _assert_bag("latitude_dependent_albedo", ref, dev, cuda_backend, "net_ocean", False)
    g = ref.grid
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    sw, tr = ref.rad_state.sw[rows, cols], ref.rad_fluxes_ocean.downwelling_shortwave[rows, cols]
    act = (ref.inactive[rows, cols] == 0) & (sw > 1.0)
    alpha = 1.0 - (tr / np.where(sw > 0, sw, 1.0))            # the stored diagnostic is −ℐₜ = (1 − α) SW (apply_air_sea_radiative_fluxes.jl:108)
    phi = np.deg2rad(g.phi[rows])[:, None] * np.ones_like(sw)
    assert np.allclose(alpha[act], (0.069 - 0.011 * np.cos(2 * phi))[act], rtol=0, atol=1e-12)


def test_two_color_shortwave_routing(oracle_lib, cuda_backend, cuda_lib):
    """With a TwoColorRadiation ocean the transmitted shortwave leaves JT and is stored, divided by ρc, in the scheme's
    surface_flux (src/Oceans/radiative_forcing.jl:84-91)."""
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT="f64", two_color_radiation=True)
    plain = synthetic.build_case("C1", cuda_backend, FT="f64", atm_FT="f64")
    for ci in (ref, dev, plain):
        ci.initialize()
        ci.update_state(T_STEP)
    cuda_backend.synchronize()
    _assert_bag("two_color", ref, dev, cuda_backend, "net_ocean", False)
    _assert_bag("two_color", ref, dev, cuda_backend, "rad_fluxes_ocean", False)
    ref.tc = ne_b200.package.interface._Fields(surface_flux=ref.two_color_surface_flux)
    dev.tc = ne_b200.package.interface._Fields(surface_flux=dev.two_color_surface_flux)
    _assert_bag("two_color", ref, dev, cuda_backend, "tc", False)
    # J⁰ = −ℐₜ/(ρc) is what JT no longer receives: JT(plain) == JT(two colour) − J⁰, up to rounding
    g = dev.grid
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    a = cuda_backend.to_numpy(dev.net_ocean.T)[rows, cols] - cuda_backend.to_numpy(dev.two_color_surface_flux)[rows, cols]
    b = cuda_backend.to_numpy(plain.net_ocean.T)[rows, cols]
    act = cuda_backend.to_numpy(dev.inactive)[rows, cols] == 0    # J⁰ is stored at masked cells too, JT is masked
    assert np.abs(a - b)[act].max() <= 1e-15 * max(np.abs(b).max(), 1e-300) * 16
    assert np.abs(cuda_backend.to_numpy(dev.two_color_surface_flux)[rows, cols]).max() > 0


@pytest.mark.parametrize("atm_FT", ["f64", "f32"])
@pytest.mark.parametrize("entry", ["interpolate_state", "fused_interface_step"])
def test_barotropic_potential(oracle_lib, cuda_backend, cuda_lib, atm_FT, entry):
    """potential .= p ./ ρᵒᶜ (interpolate_atmospheric_state.jl:80-85): bit-exact (one IEEE division of the bit-exact
    interpolated pressure), through the direct / staged / merged interpolation kernels."""
    cfg = "C1" if entry == "interpolate_state" else dict(nx=600, ny=40, latitude=(-60.0, 60.0), src_nx=64, src_ny=32)
    ref, dev = build_pair(cfg, oracle_lib, cuda_backend, FT="f64", atm_FT=atm_FT, barotropic_potential=True)
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP)
    if entry == "interpolate_state":
        dev.interpolate_state(T_STEP)
    else:
        dev.fused_interface_step(T_STEP)
    cuda_backend.synchronize()
    g = ref.grid
    pr, pd = g.interior(ref.barotropic_potential), g.interior(cuda_backend.to_numpy(dev.barotropic_potential))
    assert np.array_equal(pr, pd)
    assert np.array_equal(pr, g.interior(ref.atmos_state.p) / 1020.0) or np.array_equal(pr, g.interior(ref.atmos_state.p) / ref.ocean_properties.reference_density)
    assert pr.min() > 90.0


def test_host_pipeline_refreshes_clock_dependent_albedo(cuda_backend, cuda_lib):
    """TabulatedAlbedo depends on the clock (seconds in the day, solar declination: tabulated_albedo.jl:104-131).  The host
    pipeline keeps one descriptor for its whole life: at t > 0 it must equal the fused step, which rebuilds its own."""
    import torch
    from test_albedo import _table
    rng = np.random.default_rng(3)
    table, phi_values, t_values = _table(rng)
    cases = []
    for _ in range(2):
        ci = synthetic.build_case("C1", cuda_backend, FT="f64", atm_FT="f32")
        ci.radiation.surface_properties["ocean"] = ne_b200.SurfaceRadiationProperties(
            ne_b200.TabulatedAlbedo(cuda_backend.from_numpy(table), phi_values, t_values), 0.97)
        ci.initialize()
        cases.append(ci)
    a, b = cases
    pinned = {k: torch.from_numpy(np.ascontiguousarray(b._host_inputs["ocean"][k])).pin_memory() for k in ("T", "S", "u", "v")}
    pipe = ne_b200.HostPipelinedStep(b, n_chunks=4)
    for t in (0.0, 86400.0 * 200 + 41000.0):
        a.fused_interface_step(t)
        pipe.step(t, pinned)
        cuda_backend.synchronize()
        for bag in ("net_ocean", "rad_fluxes_ocean", "ao_fluxes"):
            for n in getattr(a, bag).names():
                x, y = cuda_backend.to_numpy(getattr(getattr(a, bag), n)), cuda_backend.to_numpy(getattr(getattr(b, bag), n))
                assert np.array_equal(x, y, equal_nan=True), f"t = {t}: {bag}.{n} differs between the fused step and the host pipeline"
    sw0 = cuda_backend.to_numpy(a.rad_fluxes_ocean.downwelling_shortwave).copy()
    a.fused_interface_step(0.0)
    cuda_backend.synchronize()
    assert not np.array_equal(sw0, cuda_backend.to_numpy(a.rad_fluxes_ocean.downwelling_shortwave)), "the albedo should depend on the clock"


def test_host_pipeline_returns_results_to_host(cuda_backend, cuda_lib):
    """NeHostStepDesc.out_fields: the net ocean fluxes leave for pinned host memory band by band behind the kernels;
    after synchronising the compute stream the host copies equal the device arrays, and the step itself is unchanged."""
    import torch
    a = synthetic.build_case("C1", cuda_backend, FT="f64", atm_FT="f32")
    b = synthetic.build_case("C1", cuda_backend, FT="f64", atm_FT="f32")
    a.initialize(); b.initialize()
    a.fused_interface_step(T_STEP)
    pinned = {k: torch.from_numpy(np.ascontiguousarray(b._host_inputs["ocean"][k])).pin_memory() for k in ("T", "S", "u", "v")}
    back = {n: (getattr(b.net_ocean, n), torch.full(getattr(b.net_ocean, n).shape, float("nan"), dtype=torch.float64).pin_memory())
            for n in b.net_ocean.names()}
    for n_chunks in (1, 5):
        pipe = ne_b200.HostPipelinedStep(b, n_chunks=n_chunks, return_fields=back)
        assert pipe.d2h_bytes_per_step() == 6 * b.ocean_state.T.numel() * 8
        for _ in range(2):
            pipe.step(T_STEP, pinned)
        torch.cuda.current_stream().synchronize()       # the compute stream waits for the last D2H copy
        for n, (dev, host) in back.items():
            assert torch.equal(host, dev.cpu()), f"{n}: host copy differs (n_chunks={n_chunks})"
            assert np.array_equal(cuda_backend.to_numpy(getattr(a.net_ocean, n)), host.numpy(), equal_nan=True), n
            host.fill_(float("nan"))
