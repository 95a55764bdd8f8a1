"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): interpolation indices / weights / interpolated values bit-exact;
fluxes within 1e-10 relative in Float64 (1e-5 in Float32) at the same iteration count."""
import numpy as np
import pytest

import ne_b200
from numericalearth_jl_b200 import synthetic
from parity import build_pair, compare_fields, converged_mask

pytestmark = pytest.mark.gpu

F64_TOL = 1e-10
F32_TOL = 1e-5
T_STEP = 0.37 * 10800.0


def _check_iterations(ref, dev, backend, max_mismatch_rate=2e-4):
    g = ref.grid
    ri = g.interior(ref.ao_iterations)
    di = g.interior(backend.to_numpy(dev.ao_iterations))
    mism = float((ri != di).mean())
    assert mism <= max_mismatch_rate, f"iteration-count mismatch rate {mism}"
    assert np.abs(ri.astype(int) - di.astype(int)).max() <= 1
    return mism


@pytest.mark.parametrize("atm_FT,stretched", [("f64", False), ("f32", False), ("f64", True), ("f32", True)])
def test_fractional_indices_and_interpolation_bit_exact(oracle_lib, cuda_backend, cuda_lib, atm_FT, stretched):
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT=atm_FT, stretched_latitude=stretched)
    ref.initialize(); dev.initialize()
    for t in (0.0, T_STEP, 10800.0, 1.5 * 10800.0):   # ñ = 0, 0.37, exactly on a snapshot (wrap), mid wrap interval
        ref.interpolate_state(t); dev.interpolate_state(t)
        cuda_backend.synchronize()
        for bag_r, bag_d in ((ref.frac, dev.frac), (ref.rad_frac, dev.rad_frac), (ref.atmos_state, dev.atmos_state),
                             (ref.rad_state, dev.rad_state)):
            res = compare_fields(bag_r, bag_d, ref.grid, cuda_backend)
            for n, (_, _, exact) in res.items():
                assert exact, f"{n} not bit-exact (atm {atm_FT}, stretched={stretched}, t={t})"


@pytest.mark.parametrize("FT,atm_FT", [("f64", "f64"), ("f64", "f32"), ("f32", "f32")])
def test_staged_interpolation_bit_exact(oracle_lib, cuda_backend, cuda_lib, FT, atm_FT):
    """Exchange grid >= 4x finer than the source: the shared-memory staged kernel runs (blocks whose window wraps the
    periodic seam take its direct path).  Same bar as the direct kernel: bit-exact against the oracle."""
    cfg = dict(nx=600, ny=40, latitude=(-60.0, 60.0), src_nx=64, src_ny=32)
    ref, dev = build_pair(cfg, oracle_lib, cuda_backend, FT=FT, atm_FT=atm_FT)
    ref.initialize(); dev.initialize()
    for t in (0.0, T_STEP, 10800.0):
        ref.interpolate_state(t); dev.interpolate_state(t)
        cuda_backend.synchronize()
        for bag_r, bag_d in ((ref.atmos_state, dev.atmos_state), (ref.rad_state, dev.rad_state)):
            for n, (_, _, exact) in compare_fields(bag_r, bag_d, ref.grid, cuda_backend).items():
                assert exact, f"{n} not bit-exact (staged, atm {atm_FT}, t={t})"


@pytest.mark.parametrize("config", ["C1", "C2"])
def test_atmosphere_ocean_fluxes_f64(oracle_lib, cuda_backend, cuda_lib, config):
    ref, dev = build_pair(config, oracle_lib, cuda_backend, FT="f64", atm_FT="f64")
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP); dev.update_state(T_STEP)
    cuda_backend.synchronize()
    res = compare_fields(ref.ao_fluxes, dev.ao_fluxes, ref.grid, cuda_backend)
    for n, (r, fr, _) in res.items():
        assert fr <= F64_TOL, f"{n}: field-relative error {fr}"
    # interface temperature and the net fluxes downstream
    t = compare_fields(ne_b200.package.interface._Fields(T=ref.ao_temperature), ne_b200.package.interface._Fields(T=dev.ao_temperature),
                       ref.grid, cuda_backend)
    assert t["T"][1] <= F64_TOL
    res = compare_fields(ref.net_ocean, dev.net_ocean, ref.grid, cuda_backend, with_halo_ring=False)
    for n, (r, fr, _) in res.items():
        assert fr <= F64_TOL, f"net {n}: {fr}"
    res = compare_fields(ref.rad_fluxes_ocean, dev.rad_fluxes_ocean, ref.grid, cuda_backend, with_halo_ring=False)
    for n, (r, fr, _) in res.items():
        assert fr <= F64_TOL, f"radiative {n}: {fr}"
    _check_iterations(ref, dev, cuda_backend)


def test_generic_kernel_on_default_tree(oracle_lib, cuda_backend, cuda_lib, monkeypatch):
    """The default plugin tree normally takes the specialised kernel; NE_B200_FORCE_GENERIC routes it
    through the generic variant kernel, which must agree with the oracle as well."""
    monkeypatch.setenv("NE_B200_FORCE_GENERIC", "1")
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT="f64")
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP); dev.update_state(T_STEP)
    cuda_backend.synchronize()
    res = compare_fields(ref.ao_fluxes, dev.ao_fluxes, ref.grid, cuda_backend)
    for n, (r, fr, _) in res.items():
        assert fr <= F64_TOL, f"{n}: {fr}"
    _check_iterations(ref, dev, cuda_backend)


def test_closed_form_kernel_on_default_tree(oracle_lib, cuda_backend, cuda_lib, monkeypatch):
    """The default tree normally runs the table-driven iteration (ne_flux_tab.cuh); NE_B200_CLOSED_FORM_PSI
    keeps the libdevice closed-form iteration (the fallback for ψ parameters the tables cannot fit)."""
    monkeypatch.setenv("NE_B200_CLOSED_FORM_PSI", "1")
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT="f64")
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP); dev.update_state(T_STEP)
    cuda_backend.synchronize()
    res = compare_fields(ref.ao_fluxes, dev.ao_fluxes, ref.grid, cuda_backend)
    for n, (r, fr, _) in res.items():
        assert fr <= F64_TOL, f"{n}: {fr}"
    _check_iterations(ref, dev, cuda_backend)


def test_sea_ice_ocean_stress_known_answer(cuda_backend, cuda_lib):
    """test/test_surface_fluxes.jl:294-336: ocean (0.1, 0.2), ice at rest, Cᴰ=1e-3, ρₑ=1000
    => τˣ == sqrt(0.1²+0.2²)·0.1, τʸ == sqrt(0.1²+0.2²)·0.2 (exact)."""
    import ctypes
    A = ne_b200.abi
    g = ne_b200.ExchangeGrid(nx=4, ny=4, hx=2, hy=2)
    b = cuda_backend
    uo, vo = b.from_numpy(np.full(g.shape, 0.1)), b.from_numpy(np.full(g.shape, 0.2))
    ui, vi = b.zeros(g.shape, "f64"), b.zeros(g.shape, "f64")
    tx, ty = b.zeros(g.shape, "f64"), b.zeros(g.shape, "f64")
    d = A.NeSeaIceOceanStressDesc()
    d.grid = g.pod(True)
    d.ui, d.vi, d.uo, d.vo = b.ptr(ui), b.ptr(vi), b.ptr(uo), b.ptr(vo)
    d.ocean_density, d.drag_coefficient = 1000.0, 0.001
    d.x_momentum, d.y_momentum = b.ptr(tx), b.ptr(ty)
    cuda_lib.call("sea_ice_ocean_stress", "f64", d, b.stream())
    b.synchronize()
    assert (g.interior(b.to_numpy(tx)) == np.sqrt(0.1 ** 2 + 0.2 ** 2) * 0.1).all()
    assert (g.interior(b.to_numpy(ty)) == np.sqrt(0.1 ** 2 + 0.2 ** 2) * 0.2).all()


def test_atmosphere_ocean_fluxes_jra55_faithful_mixed_precision(oracle_lib, cuda_backend, cuda_lib):
    """Float64 ocean + Float32 atmosphere: q_sat is evaluated in Float32 (interface_states.jl:56-59)."""
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT="f32")
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP); dev.update_state(T_STEP)
    cuda_backend.synchronize()
    res = compare_fields(ref.ao_fluxes, dev.ao_fluxes, ref.grid, cuda_backend)
    for n, (r, fr, _) in res.items():
        assert fr <= 2e-6, f"{n}: {fr}"   # Float32 p_sat (powf/expf differ by ulps of Float32 between libms)


def test_atmosphere_ocean_fluxes_f32(oracle_lib, cuda_backend, cuda_lib):
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f32", atm_FT="f32")
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP); dev.update_state(T_STEP)
    cuda_backend.synchronize()
    res = compare_fields(ref.ao_fluxes, dev.ao_fluxes, ref.grid, cuda_backend)
    for n, (r, fr, _) in res.items():
        assert fr <= F32_TOL, f"{n}: {fr}"


VARIANTS = {
    "large_yeager_psi": dict(atmosphere_ocean_fluxes=lambda: ne_b200.SimilarityTheoryFluxes(
        stability_functions=ne_b200.large_yeager_stability_functions())),
    "fixed_iterations": dict(atmosphere_ocean_fluxes=lambda: ne_b200.SimilarityTheoryFluxes(
        solver_stop_criteria=ne_b200.FixedIterations(5))),
    "wind_velocity_coare_profile": dict(
        atmosphere_ocean_fluxes=lambda: ne_b200.SimilarityTheoryFluxes(similarity_form=ne_b200.COARELogarithmicSimilarityProfile()),
        atmosphere_ocean_velocity_difference=lambda: ne_b200.WindVelocity()),
    "wave_formulation_temperature_viscosity": dict(atmosphere_ocean_fluxes=lambda: ne_b200.SimilarityTheoryFluxes(
        momentum_roughness_length=ne_b200.MomentumRoughnessLength(wave_formulation=ne_b200.WindDependentWaveFormulation(),
                                                                  air_kinematic_viscosity=ne_b200.TemperatureDependentAirViscosity()),
        temperature_roughness_length=ne_b200.ScalarRoughnessLength(air_kinematic_viscosity=ne_b200.TemperatureDependentAirViscosity()),
        water_vapor_roughness_length=ne_b200.ScalarRoughnessLength(air_kinematic_viscosity=ne_b200.TemperatureDependentAirViscosity()))),
    "wind_dependent_waves": dict(atmosphere_ocean_fluxes=lambda: ne_b200.SimilarityTheoryFluxes(
        momentum_roughness_length=ne_b200.MomentumRoughnessLength(wave_formulation=ne_b200.WindDependentWaveFormulation()))),
    "large_yeager_psi_mesoscale_wind_waves": dict(atmosphere_ocean_fluxes=lambda: ne_b200.SimilarityTheoryFluxes(
        stability_functions=ne_b200.large_yeager_stability_functions(),
        subgrid_velocities=ne_b200.SubgridVelocityCorrection(mesoscale=ne_b200.mahrt_sun_subgrid_velocity(100e3)),
        momentum_roughness_length=ne_b200.MomentumRoughnessLength(wave_formulation=ne_b200.WindDependentWaveFormulation()))),
    "constant_roughness_no_gustiness_neutral": dict(atmosphere_ocean_fluxes=lambda: ne_b200.SimilarityTheoryFluxes(
        stability_functions=None, subgrid_velocities=None, momentum_roughness_length=1e-4,
        temperature_roughness_length=1e-4, water_vapor_roughness_length=1e-4)),
    "mesoscale_subgrid_velocity": dict(atmosphere_ocean_fluxes=lambda: ne_b200.SimilarityTheoryFluxes(
        subgrid_velocities=ne_b200.SubgridVelocityCorrection(mesoscale=ne_b200.mahrt_sun_subgrid_velocity(100e3)))),
    "skin_temperature_diffusive": dict(
        atmosphere_ocean_interface_temperature=lambda: ne_b200.SkinTemperature(ne_b200.DiffusiveFlux(1e-2, 1.0))),
    "salinity_dependent_mole_fraction": dict(
        atmosphere_ocean_interface_specific_humidity=lambda: ne_b200.ImpureSaturationSpecificHumidity(
            ne_b200.Liquid(), ne_b200.WaterMoleFraction())),
    "coefficient_based_constant": dict(atmosphere_ocean_fluxes=lambda: ne_b200.CoefficientBasedFluxes(
        transfer_coefficients=(1e-2, 1e-3, 1e-3))),
    "coefficient_based_polynomial_drag": dict(atmosphere_ocean_fluxes=lambda: ne_b200.CoefficientBasedFluxes(
        transfer_coefficients=(ne_b200.PolynomialNeutralDragCoefficient(), 1.1e-3, 1.2e-3))),
    "large_yeager_coefficients": dict(atmosphere_ocean_fluxes=lambda: ne_b200.CoefficientBasedFluxes(
        transfer_coefficients=ne_b200.LargeYeagerTransferCoefficients(), solver_stop_criteria=ne_b200.FixedIterations(5))),
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_atmosphere_ocean_variants(oracle_lib, cuda_backend, cuda_lib, name):
    kw = {k: v() for k, v in VARIANTS[name].items()}
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT="f64", **kw)
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP); dev.update_state(T_STEP)
    cuda_backend.synchronize()
    maxiter = dev.ao_flux_formulation.solver_stop_criteria.maxiter if hasattr(dev.ao_flux_formulation.solver_stop_criteria, "maxiter") else 10**9
    res = compare_fields(ref.ao_fluxes, dev.ao_fluxes, ref.grid, cuda_backend,
                         mask=converged_mask(ref.ao_iterations, ref.grid, maxiter, dilate=False))
    for n, (r, fr, _) in res.items():
        assert fr <= F64_TOL, f"{name}/{n}: {fr}"
    _check_iterations(ref, dev, cuda_backend, max_mismatch_rate=1e-3)


@pytest.mark.parametrize("asi_temperature", ["conductive", "ice_snow", "bulk"])
def test_ocean_sea_ice_model_step(oracle_lib, cuda_backend, cuda_lib, asi_temperature):
    """Config C3: atmosphere-ocean, atmosphere-sea-ice and sea-ice-ocean kernels together."""
    tf = {"conductive": lambda: ne_b200.SkinTemperature(ne_b200.ConductiveFlux(2.0)),
          "ice_snow": lambda: ne_b200.SkinTemperature(ne_b200.IceSnowConductiveFlux(0.31, 2.0)),
          "bulk": lambda: ne_b200.BulkTemperature()}[asi_temperature]
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT="f64", sea_ice=True,
                          atmosphere_sea_ice_interface_temperature=tf())
    ref.initialize(); dev.initialize()
    host = ne_b200.NumpyHostBackend()
    colr = synthetic.ocean_column(ref.grid, host, nz=10) + (1200.0, 10, 0)
    cold = synthetic.ocean_column(dev.grid, cuda_backend, nz=10) + (1200.0, 10, 0)
    ref.update_state(T_STEP, ocean_column=colr); dev.update_state(T_STEP, ocean_column=cold)
    cuda_backend.synchronize()
    it = np.maximum(ref.asi_iterations, ref.ao_iterations)
    for bag_r, bag_d, ring in ((ref.asi_fluxes, dev.asi_fluxes, True), (ref.sio_fluxes, dev.sio_fluxes, False),
                               (ref.net_sea_ice, dev.net_sea_ice, False), (ref.net_ocean, dev.net_ocean, False),
                               (ref.rad_fluxes_sea_ice, dev.rad_fluxes_sea_ice, False)):
        res = compare_fields(bag_r, bag_d, ref.grid, cuda_backend, with_halo_ring=ring,
                             mask=converged_mask(it, ref.grid, 100, with_halo_ring=ring))
        for n, (r, fr, _) in res.items():
            assert fr <= F64_TOL, f"{asi_temperature}/{n}: {fr}"
        res = compare_fields(bag_r, bag_d, ref.grid, cuda_backend, with_halo_ring=ring)   # incl. limit-cycle points
        for n, (r, fr, _) in res.items():
            assert fr <= 1e-3, f"{asi_temperature}/{n} (non-converged points): {fr}"
    ri, di = ref.grid.interior(ref.asi_iterations), ref.grid.interior(cuda_backend.to_numpy(dev.asi_iterations))
    assert float((ri != di).mean()) <= 1e-3
    # frazil clamp: the whole T column is bit-exact (pure compare/select)
    assert np.array_equal(colr[0], cuda_backend.to_numpy(cold[0]))
    Ts_r, Ts_d = ref.sea_ice_state.top_temperature, cuda_backend.to_numpy(dev.sea_ice_state.top_temperature)
    m = converged_mask(it, ref.grid, 100)
    assert np.nanmax(np.abs(ref.grid.interior(Ts_r) - ref.grid.interior(Ts_d))[m]) <= 1e-8


@pytest.mark.parametrize("atm_FT", ["f64", "f32"])
def test_sea_ice_work_queue_kernel_against_generic_kernel(cuda_backend, cuda_lib, monkeypatch, atm_FT):
    """Default a-si tree at 1/4 degree: the work-queue kernel with the tabulated Split(SHEBA, Paulson) functions
    against the generic kernel (libdevice closed forms).  Converged points to 1e-10, same trip counts."""
    dev = synthetic.build_case("C2", cuda_backend, FT="f64", atm_FT=atm_FT, sea_ice=True, with_iterations=True)
    dev.initialize()
    dev.interpolate_state(T_STEP)
    g = dev.grid
    Ts0 = dev.sea_ice_state.top_temperature.clone()

    def run():
        dev.sea_ice_state.top_temperature.copy_(Ts0)   # READ-MODIFY-WRITE input
        dev.compute_atmosphere_sea_ice_fluxes()
        cuda_backend.synchronize()
        out = {n: g.interior(cuda_backend.to_numpy(getattr(dev.asi_fluxes, n))).copy() for n in dev.asi_fluxes.names()}
        out["Ts"] = g.interior(cuda_backend.to_numpy(dev.sea_ice_state.top_temperature)).copy()
        return out, g.interior(cuda_backend.to_numpy(dev.asi_iterations)).copy()

    fast, fit = run()
    monkeypatch.setenv("NE_B200_FORCE_GENERIC", "1")
    gen, git = run()
    monkeypatch.delenv("NE_B200_FORCE_GENERIC")
    assert any((fast[n] != gen[n]).any() for n in fast), "the a-si fast path did not run (results identical to the generic kernel's)"
    assert (fit > 0).sum() > 10000 and ((fit > 0) == (git > 0)).all()
    assert float((fit != git).mean()) <= 1e-3, float((fit != git).mean())
    conv = (git < 100) & (fit < 100)
    for n in fast:
        a, b = fast[n], gen[n]
        assert np.isfinite(a).all(), n
        s = float(np.abs(b).max()) or 1.0
        assert float(np.abs(a - b)[conv].max()) / s <= F64_TOL, f"{n}: {float(np.abs(a - b)[conv].max()) / s}"
        assert float(np.abs(a - b).max()) / s <= 1e-3, f"{n} (limit-cycle points): {float(np.abs(a - b).max()) / s}"


@pytest.mark.parametrize("atm_FT", ["f64", "f32"])
@pytest.mark.parametrize("single_pass", [False, True])
def test_fused_interface_step_matches_unfused(oracle_lib, cuda_backend, cuda_lib, atm_FT, single_pass, monkeypatch):
    """One C-ABI call (single-pass interpolation + solve kernel, then assembly and radiation) against the
    oracle's phase-by-phase update_state!, and to 1e-13 against this library's own component kernels."""
    if single_pass:   # opt-in single-pass interpolation + solve kernel
        monkeypatch.setenv("NE_B200_FUSE_INTERP", "1")
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT=atm_FT)
    _, dev2 = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT=atm_FT)
    ref.initialize(); dev.initialize(); dev2.initialize()
    ref.update_state(T_STEP); dev.fused_interface_step(T_STEP); dev2.update_state(T_STEP)
    cuda_backend.synchronize()
    for bag_r, bag_d in ((ref.atmos_state, dev.atmos_state), (ref.rad_state, dev.rad_state)):
        for n, (_, _, exact) in compare_fields(bag_r, bag_d, ref.grid, cuda_backend).items():
            assert exact, f"fused interpolation of {n} not bit-exact"
    # Float32 atmosphere => q_sat in Float32 (interface_states.jl:56-59): powf/expf differ by Float32 ulps between libms
    tol = F64_TOL if atm_FT == "f64" else 2e-6
    for bag_r, bag_d, ring in ((ref.ao_fluxes, dev.ao_fluxes, True), (ref.net_ocean, dev.net_ocean, False),
                               (ref.rad_fluxes_ocean, dev.rad_fluxes_ocean, False)):
        res = compare_fields(bag_r, bag_d, ref.grid, cuda_backend, with_halo_ring=ring)
        for n, (r, fr, _) in res.items():
            assert fr <= tol, f"{n}: {fr}"
    if atm_FT == "f64":
        _check_iterations(ref, dev, cuda_backend)
    for bag in ("ao_fluxes", "net_ocean", "rad_fluxes_ocean"):
        a, b = getattr(dev, bag), getattr(dev2, bag)
        for n in a.names():
            x, y = cuda_backend.to_numpy(getattr(a, n)), cuda_backend.to_numpy(getattr(b, n))
            scale = float(np.abs(y).max()) or 1.0   # same arithmetic, separately compiled: FMA contraction may differ
            assert float(np.abs(x - y).max()) <= 1e-13 * scale, f"fused vs component kernels differ in {bag}.{n}"


def test_fused_interface_step_without_materialised_atmosphere_state(oracle_lib, cuda_backend, cuda_lib, monkeypatch):
    """Single-pass kernel: outputs of the interpolation left NULL are never written (they keep their sentinel),
    the fluxes are unchanged."""
    monkeypatch.setenv("NE_B200_FUSE_INTERP", "1")
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT="f32")
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP)
    sentinel = -12345.0
    for n in ("u", "v", "T", "q", "p"):
        getattr(dev.atmos_state, n).fill_(sentinel)
    d = dev.fused_step_desc(T_STEP)
    for f in range(5):
        d.atmosphere.out[f] = None
    for name in ("ua", "va", "Ta", "pa", "qa"):
        setattr(d.ao, name, None)
    cuda_lib.call("fused_interface_step", "f64", d, cuda_backend.stream())
    cuda_backend.synchronize()
    for n in ("u", "v", "T", "q", "p"):
        assert bool((cuda_backend.to_numpy(getattr(dev.atmos_state, n)) == sentinel).all()), f"{n} was materialised"
    for bag_r, bag_d, ring in ((ref.ao_fluxes, dev.ao_fluxes, True), (ref.net_ocean, dev.net_ocean, False)):
        res = compare_fields(bag_r, bag_d, ref.grid, cuda_backend, with_halo_ring=ring)
        for n, (r, fr, _) in res.items():
            assert fr <= 2e-6, f"{n}: {fr}"   # Float32 q_sat, see the mixed-precision test


@pytest.mark.parametrize("n_blocks", [3, 296, 1184])
def test_diagnostics_reduction_matches_numpy_and_is_deterministic(cuda_backend, cuda_lib, n_blocks):
    """ne_diag_reduce (src/Diagnostics/interface_fluxes.jl:90-195 integrals): area-weighted masked sums against
    numpy in Float64, bit-identical from run to run (fixed two-stage order)."""
    from numericalearth_jl_b200 import sharding
    dev = synthetic.build_case("C1", cuda_backend, FT="f64", atm_FT="f64")
    dev.initialize()
    dev.update_state(T_STEP)
    f = dev.ao_fluxes
    fields = [f.latent_heat, f.sensible_heat, f.water_vapor, f.x_momentum, f.y_momentum, dev.net_ocean.T]
    diag = sharding.FluxDiagnostics(dev, fields, n_blocks=n_blocks)
    r1 = cuda_backend.to_numpy(diag.reduce()).copy()
    cuda_backend.synchronize()
    r1 = cuda_backend.to_numpy(diag.result).copy()
    diag.reduce()
    cuda_backend.synchronize()
    r2 = cuda_backend.to_numpy(diag.result).copy()
    assert np.array_equal(r1, r2)
    g = dev.grid
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    act = ~cuda_backend.to_numpy(dev.inactive)[rows, cols].astype(bool)
    area = cuda_backend.to_numpy(diag.area)[rows, cols]
    for k, x in enumerate(fields):
        v = cuda_backend.to_numpy(x)[rows, cols]
        ref = float((v * area)[act].sum(dtype=np.float64))
        scale = float(np.abs(v * area)[act].sum())
        assert abs(r1[k] - ref) <= 1e-12 * scale, (k, r1[k], ref)


@pytest.mark.parametrize("n_chunks", [1, 3, 8, 37])
def test_host_pipelined_step_equals_update_state(cuda_backend, cuda_lib, n_chunks):
    """Ocean state in pinned host buffers, chunked H2D overlapped with the band-restricted kernels: every flux field is
    bit-identical to the plain device-resident update_state!."""
    import torch
    a = synthetic.build_case("C1", cuda_backend, FT="f64", atm_FT="f32")
    b = synthetic.build_case("C1", cuda_backend, FT="f64", atm_FT="f32")
    a.initialize(); b.initialize()
    a.update_state(T_STEP)
    host = {k: torch.from_numpy(np.ascontiguousarray(b._host_inputs["ocean"][k])).pin_memory() for k in ("T", "S", "u", "v")}
    for k in ("T", "S", "u", "v"):
        getattr(b.ocean_state, k).fill_(float("nan"))      # the device copies must really come from the host buffers
    pipe = ne_b200.HostPipelinedStep(b, n_chunks=n_chunks)
    pipe.step(T_STEP, host)
    pipe.step(T_STEP, host)                                 # a second step reuses the events/streams
    cuda_backend.synchronize()
    for bag in ("ao_fluxes", "net_ocean", "rad_fluxes_ocean", "atmos_state"):
        x, y = getattr(a, bag), getattr(b, bag)
        for n in x.names():
            assert np.array_equal(cuda_backend.to_numpy(getattr(x, n)), cuda_backend.to_numpy(getattr(y, n)), equal_nan=True), \
                f"{bag}.{n} differs (n_chunks={n_chunks})"


@pytest.mark.parametrize("FT", ["f64", "f32"])
def test_elevation_correction_parity(oracle_lib, cuda_backend, cuda_lib, FT):
    """Phase 1.5 (atmosphere_state_correction.jl:133-146) and the fluxes downstream of the corrected state."""
    g_kw = dict(nx=96, ny=40, latitude=(-70.0, 70.0))
    rng = np.random.default_rng(11)
    zs = rng.uniform(0.0, 1500.0, (40, 96))
    kw = dict(FT=FT, atm_FT=FT, atmosphere_correction=ne_b200.ElevationCorrection(zs, 120.0))
    ref, dev = build_pair(g_kw, oracle_lib, cuda_backend, **kw)
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP); dev.fused_interface_step(T_STEP)
    cuda_backend.synchronize()
    tol = 1e-14 if FT == "f64" else 2e-6
    res = compare_fields(ref.atmos_state, dev.atmos_state, ref.grid, cuda_backend)
    for n, (r, fr, exact) in res.items():
        assert exact or (n in ("T", "p") and r <= tol), f"{n}: {r}"
    res = compare_fields(ref.ao_fluxes, dev.ao_fluxes, ref.grid, cuda_backend)
    for n, (r, fr, _) in res.items():
        assert fr <= (F64_TOL if FT == "f64" else F32_TOL), f"{n}: {fr}"


def _ao_outputs(ci, backend):
    out = {n: backend.to_numpy(getattr(ci.ao_fluxes, n)).copy() for n in ci.ao_fluxes.names()}
    out["interface_temperature"] = backend.to_numpy(ci.ao_temperature).copy()
    out["iterations"] = backend.to_numpy(ci.ao_iterations).copy()
    return out


@pytest.mark.parametrize("setting", ["NE_B200_TAB2_ORDER=0", "NE_B200_TAB2_ORDER=1", "NE_B200_TAB2_ORDER=2",
                                     "NE_B200_TAB2_ORDER=2,NE_B200_TAB2_DESCENDING=1", "NE_B200_TAB2_STAGING=1", "NE_B200_TAB2_SHAPE=0",
                                     "NE_B200_TAB2_SHAPE=3"])
def test_trip_ordering_passes_and_launch_shapes_do_not_change_a_bit(cuda_backend, cuda_lib, monkeypatch, setting):
    """Which lane of which warp computes a point must not change its arithmetic: the three ordering passes of the round-2 solve
    (bitonic network on (trips, record), stable counting sort on trips, histogram on (trips, record / 4) — whose permutation is
    not even reproducible from run to run), the descending draw order, the cp.async staging and the other CTA shapes against
    the launch in memory order, every output and the trip counts bit for bit.  A permutation that lost or doubled a point would
    leave a point unwritten (NaN-filled beforehand) or fail the comparison."""
    dev = synthetic.build_case("C2", cuda_backend, FT="f64", atm_FT="f32", with_iterations=True)
    dev.initialize()
    dev.interpolate_state(T_STEP)
    monkeypatch.setenv("NE_B200_TAB2_NO_ORDER", "1")
    dev.compute_atmosphere_ocean_fluxes()
    cuda_backend.synchronize()
    plain = _ao_outputs(dev, cuda_backend)
    monkeypatch.delenv("NE_B200_TAB2_NO_ORDER")
    for kv in setting.split(","):
        k, v = kv.split("=")
        monkeypatch.setenv(k, v)
    for step in range(3):      # the first ordered launch sorts on the hints the launches before it left, the later ones on their own
        for n in dev.ao_fluxes.names():
            getattr(dev.ao_fluxes, n).fill_(float("nan"))
        dev.ao_iterations.fill_(-1)
        dev.compute_atmosphere_ocean_fluxes()
        cuda_backend.synchronize()
        got = _ao_outputs(dev, cuda_backend)
        g = dev.grid
        for n, a in plain.items():
            x, y = g.interior(a), g.interior(got[n])
            assert np.array_equal(x, y, equal_nan=True), f"{setting}, launch {step}: {n} differs at {int((x != y).sum())} points"
        assert not np.isnan(g.interior(got["friction_velocity"])).any()


@pytest.mark.parametrize("atm_FT", ["f64", "f32"])
@pytest.mark.parametrize("theta", ["0", "8", "20"])
def test_work_queue_kernel_equals_one_thread_per_point_kernel_bitwise(cuda_backend, cuda_lib, monkeypatch, atm_FT, theta):
    """ne_flux_queue.cuh carries every point's iterate across rounds bit for bit: outputs AND trip counts equal those
    of the one-thread-per-point table kernel (NE_B200_TAB_CLASSIC=1), whatever the deferral threshold."""
    dev = synthetic.build_case("C2", cuda_backend, FT="f64", atm_FT=atm_FT, with_iterations=True)
    dev.initialize()
    dev.interpolate_state(T_STEP)
    monkeypatch.setenv("NE_B200_TAB_CLASSIC", "1")
    monkeypatch.setenv("NE_B200_TAB_V1", "1")      # the round-1 one-thread-per-point kernel (the queue kernel embeds ITS iteration)
    dev.compute_atmosphere_ocean_fluxes()
    cuda_backend.synchronize()
    classic = _ao_outputs(dev, cuda_backend)
    monkeypatch.delenv("NE_B200_TAB_CLASSIC")
    monkeypatch.delenv("NE_B200_TAB_V1")
    monkeypatch.setenv("NE_B200_QUEUE", "1")
    monkeypatch.setenv("NE_B200_QUEUE_THETA", theta)
    for n in dev.ao_fluxes.names():
        getattr(dev.ao_fluxes, n).fill_(float("nan"))
    dev.ao_iterations.fill_(-1)
    dev.compute_atmosphere_ocean_fluxes()
    cuda_backend.synchronize()
    queue = _ao_outputs(dev, cuda_backend)
    g = dev.grid
    # trip counts and the converged iterate: bit for bit; the flux epilogue is compiled once per kernel (FMA
    # contraction may differ between the two inlining contexts): 1e-13
    for n, a in classic.items():
        x, y = g.interior(a), g.interior(queue[n])
        if n in ("iterations", "friction_velocity", "temperature_scale", "water_vapor_scale", "interface_temperature"):
            assert np.array_equal(x, y, equal_nan=True), f"{n} differs at {int((x != y).sum())} points (theta={theta})"
        else:
            s = float(np.abs(x).max()) or 1.0
            assert float(np.abs(x - y).max()) / s <= 1e-13, f"{n} (theta={theta}): {float(np.abs(x - y).max()) / s}"


def test_float32_model_fast_path_against_generic_kernel_and_oracle(oracle_lib, cuda_backend, cuda_lib, monkeypatch):
    """Float32 model, default tree: the mixed-precision table iteration on the work-queue kernel against (i) the
    generic kernel (same promotion rules, libdevice closed forms) and (ii) the oracle.  Bar: 1e-5 relative to the
    field scale; the trip counts of points that converge agree except where a last-bit difference moves the
    Float32 rounding of an iterate (tol = 1e-8 < eps(Float32): the loop stops on exact stationarity)."""
    ref, dev = build_pair("C2", oracle_lib, cuda_backend, FT="f32", atm_FT="f32")
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP)
    dev.interpolate_state(T_STEP)
    dev.compute_atmosphere_ocean_fluxes()
    cuda_backend.synchronize()
    fast = _ao_outputs(dev, cuda_backend)
    monkeypatch.setenv("NE_B200_FORCE_GENERIC", "1")
    dev.compute_atmosphere_ocean_fluxes()
    cuda_backend.synchronize()
    generic = _ao_outputs(dev, cuda_backend)
    monkeypatch.delenv("NE_B200_FORCE_GENERIC")
    g = dev.grid
    for n in dev.ao_fluxes.names():
        a, b, r = g.interior(fast[n]).astype(np.float64), g.interior(generic[n]).astype(np.float64), \
            g.interior(getattr(ref.ao_fluxes, n)).astype(np.float64)
        s = float(np.abs(r).max()) or 1.0
        assert np.isfinite(a).all(), n
        assert np.abs(a - b).max() / s <= F32_TOL, f"{n}: fast vs generic {np.abs(a - b).max() / s}"
        assert np.abs(a - r).max() / s <= F32_TOL, f"{n}: fast vs oracle {np.abs(a - r).max() / s}"
    assert any((fast[n] != generic[n]).any() for n in dev.ao_fluxes.names()), "the Float32 fast path did not run"
    fi, gi, oi = (g.interior(x).astype(int) for x in (fast["iterations"], generic["iterations"], ref.ao_iterations))
    assert ((fi > 0) == (oi > 0)).all()
    # trip counts: statistically those of the reference algorithm (a last-bit difference in a Float64 intermediate
    # moves the Float32 rounding of an iterate, and with it the trip at which the iterate becomes stationary)
    conv = (oi < 100) & (gi < 100) & (fi < 100)
    assert float((np.abs(fi[conv] - gi[conv]) > 2).mean()) <= 0.02, float((np.abs(fi[conv] - gi[conv]) > 2).mean())
    assert abs(float(fi[oi > 0].mean()) - float(oi[oi > 0].mean())) <= 0.5
    assert abs(float((fi == 100).mean()) - float((oi == 100).mean())) <= 0.005
    assert fi.max() <= 100


@pytest.mark.parametrize("FT", ["f64", "f32"])
def test_post_solve_kernel_equals_component_kernels_bitwise(cuda_backend, cuda_lib, monkeypatch, FT):
    """Assembly + radiation + diagnostics partial sums in one kernel (NeFusedStepDesc.diag) against the three
    component kernels: every output field and the diagnostics sums bit for bit."""
    from numericalearth_jl_b200 import sharding
    outs = []
    for fused in (True, False):
        if not fused:
            monkeypatch.setenv("NE_B200_NO_POST_SOLVE_FUSION", "1")
        dev = synthetic.build_case("C2", cuda_backend, FT=FT, atm_FT="f32")
        dev.initialize()
        f = dev.ao_fluxes
        diag = sharding.FluxDiagnostics(dev, [f.latent_heat, f.sensible_heat, f.water_vapor, f.x_momentum, f.y_momentum,
                                              dev.net_ocean.T, dev.net_ocean.eta])
        dev.fused_interface_step(T_STEP, diagnostics=diag)
        cuda_backend.synchronize()
        o = {"diag": cuda_backend.to_numpy(diag.result).copy()}
        for bag in ("net_ocean", "rad_fluxes_ocean"):
            for n in getattr(dev, bag).names():
                o[bag + "." + n] = cuda_backend.to_numpy(getattr(getattr(dev, bag), n)).copy()
        outs.append(o)
    monkeypatch.delenv("NE_B200_NO_POST_SOLVE_FUSION")
    assert np.isfinite(outs[0]["diag"]).all() and (outs[0]["diag"] != 0).any()
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k], equal_nan=True), k


def test_no_kernel_variant_raises(cuda_backend, cuda_lib):
    with pytest.raises(ne_b200.NoKernelVariantError):
        ne_b200.SimilarityTheoryFluxes(momentum_roughness_length=lambda u: 1e-4).pod()
    # the library itself refuses an unknown kind as well (no CPU fallback anywhere)
    dev = synthetic.build_case("tiny", cuda_backend)
    d = dev.atmosphere_ocean_desc()
    d.flux.psi_momentum.a.kind = 99
    with pytest.raises(ne_b200.NoKernelVariantError):
        cuda_lib.call("atmosphere_ocean_fluxes", "f64", d, cuda_backend.stream())


def test_full_size_properties_C4(cuda_backend, cuda_lib):
    """BASELINE 1/12° size: size-independent properties (no oracle at this size).
    (i) masked points are exactly zero / 0 °C; (ii) u★ ≥ 0, τ opposes Δu; (iii) the fixed point is
    reproduced: re-running is idempotent bit-for-bit; (iv) iteration counts within [1, maxiter]."""
    dev = synthetic.build_case("C4", cuda_backend, FT="f64", atm_FT="f32", with_iterations=True)
    dev.initialize()
    dev.update_state(T_STEP)
    cuda_backend.synchronize()
    g = dev.grid
    to = cuda_backend.to_numpy
    inactive = g.interior(to(dev.inactive)).astype(bool)
    for n in dev.ao_fluxes.names():
        a = g.interior(to(getattr(dev.ao_fluxes, n)))
        assert np.isfinite(a).all(), n
        assert (a[inactive] == 0).all(), n
    assert (g.interior(to(dev.ao_temperature))[inactive] == 0).all()
    us = g.interior(to(dev.ao_fluxes.friction_velocity))
    assert (us >= 0).all() and (us[~inactive] > 0).all()
    it = g.interior(to(dev.ao_iterations))
    assert it[~inactive].min() >= 1 and it.max() <= 100 and (it[inactive] == 0).all()
    first = {n: to(getattr(dev.ao_fluxes, n)).copy() for n in dev.ao_fluxes.names()}
    staged = {n: to(getattr(dev.atmos_state, n)).copy() for n in dev.atmos_state.names()}
    dev.update_state(T_STEP)
    cuda_backend.synchronize()
    for n, a in first.items():
        assert np.array_equal(a, to(getattr(dev.ao_fluxes, n))), n
    # (v) the shared-memory staged interpolation (taken at this size) equals the direct-gather kernel bit for bit
    import os
    os.environ["NE_B200_INTERP_DIRECT"] = "1"
    try:
        dev.interpolate_state(T_STEP)
        cuda_backend.synchronize()
    finally:
        os.environ.pop("NE_B200_INTERP_DIRECT", None)
    for n, a in staged.items():
        assert np.array_equal(a, to(getattr(dev.atmos_state, n)), equal_nan=True), f"staged vs direct interpolation: {n}"
    # (vi) the fused step interpolates atmosphere + radiation in ONE 9-series staged launch (shared fractional indices):
    # bit-identical to the two separate launches, and so is everything downstream
    rad_sep = {n: to(getattr(dev.rad_state, n)).copy() for n in dev.rad_state.names()}
    net_sep = {n: to(getattr(dev.net_ocean, n)).copy() for n in dev.net_ocean.names()}
    assert dev.shared_frac
    for bag in (dev.atmos_state, dev.rad_state, dev.net_ocean):
        for n in bag.names():
            getattr(bag, n).fill_(float("nan"))
    dev.fused_interface_step(T_STEP)
    cuda_backend.synchronize()
    for n, a in staged.items():
        assert np.array_equal(g.interior(a), g.interior(to(getattr(dev.atmos_state, n))), equal_nan=True), f"merged interpolation: {n}"
    for n, a in rad_sep.items():
        assert np.array_equal(g.interior(a), g.interior(to(getattr(dev.rad_state, n))), equal_nan=True), f"merged interpolation: {n}"
    inner = (slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx))
    for n, a in net_sep.items():
        assert np.array_equal(a[inner], to(getattr(dev.net_ocean, n))[inner], equal_nan=True), f"fused step net flux: {n}"


@pytest.mark.parametrize("case", ["fixed_iterations", "fixed_zero_iterations", "all_inactive", "no_mask", "odd_size", "array_heights"])
def test_float32_work_queue_edge_cases(oracle_lib, cuda_backend, cuda_lib, case):
    """The Float32 default tree runs on the persistent work-queue kernel: stop criteria, masks and grid sizes that
    stress its ring logic (tiles that are all inactive, partial last tiles, FixedIterations solving masked points,
    zero trips, per-point heights), each against the oracle."""
    kw = {}
    cfg = "C1"
    if case == "fixed_iterations":
        kw["atmosphere_ocean_fluxes"] = ne_b200.SimilarityTheoryFluxes(solver_stop_criteria=ne_b200.FixedIterations(7))
    elif case == "fixed_zero_iterations":
        kw["atmosphere_ocean_fluxes"] = ne_b200.SimilarityTheoryFluxes(solver_stop_criteria=ne_b200.FixedIterations(0))
    elif case == "odd_size":
        cfg = dict(nx=37, ny=11, latitude=(-60.0, 60.0))
    ref, dev = build_pair(cfg, oracle_lib, cuda_backend, FT="f32", atm_FT="f32", **kw)
    if case == "all_inactive":
        ref.inactive[:] = 1
        dev.inactive.fill_(1)
    elif case == "no_mask":
        ref.inactive = None
        dev.inactive = None
    elif case == "array_heights":   # surface-layer / boundary-layer heights as exchange-layout arrays (non-HS kernel variant)
        rng = np.random.default_rng(5)
        zs = rng.uniform(8.0, 12.0, ref.grid.shape).astype(np.float32)
        hb = rng.uniform(400.0, 700.0, ref.grid.shape).astype(np.float32)
        ref.atmosphere.surface_layer_height, ref.atmosphere.boundary_layer_height = zs, hb
        dev.atmosphere.surface_layer_height, dev.atmosphere.boundary_layer_height = cuda_backend.from_numpy(zs), cuda_backend.from_numpy(hb)
    ref.initialize(); dev.initialize()
    ref.update_state(T_STEP); dev.update_state(T_STEP)
    cuda_backend.synchronize()
    res = compare_fields(ref.ao_fluxes, dev.ao_fluxes, ref.grid, cuda_backend)
    for n, (r, fr, _) in res.items():
        assert fr <= F32_TOL, f"{case}/{n}: {fr}"
    g = ref.grid
    ri, di = g.interior(ref.ao_iterations), g.interior(cuda_backend.to_numpy(dev.ao_iterations))
    if case.startswith("fixed"):
        assert np.array_equal(ri, di)
    else:
        assert ((ri > 0) == (di > 0)).all()
    Tr, Td = g.interior(ref.ao_temperature), g.interior(cuda_backend.to_numpy(dev.ao_temperature))
    assert np.abs(Tr.astype(np.float64) - Td).max() <= 1e-4
    # net fluxes downstream
    for n, (r, fr, _) in compare_fields(ref.net_ocean, dev.net_ocean, ref.grid, cuda_backend, with_halo_ring=False).items():
        assert fr <= F32_TOL, f"{case}/net {n}: {fr}"


@pytest.mark.parametrize("case", ["fixed_iterations", "no_ice", "all_ice_no_mask"])
def test_sea_ice_work_queue_edge_cases(oracle_lib, cuda_backend, cuda_lib, case):
    """The a-si default tree on the work-queue kernel: FixedIterations (masked points ARE solved, :141-142), an ice-free
    ocean (every point takes the immediate path) and full ice cover without a mask (every tile is a full round)."""
    kw = {}
    if case == "fixed_iterations":
        kw["atmosphere_sea_ice_fluxes"] = ne_b200.SimilarityTheoryFluxes(
            stability_functions=ne_b200.atmosphere_sea_ice_stability_functions(), solver_stop_criteria=ne_b200.FixedIterations(6))
    ref, dev = build_pair("C1", oracle_lib, cuda_backend, FT="f64", atm_FT="f64", sea_ice=True, **kw)
    if case == "no_ice":
        ref.sea_ice_state.concentration[:] = 0
        dev.sea_ice_state.concentration.fill_(0)
    elif case == "all_ice_no_mask":
        for ci, one in ((ref, None), (dev, None)):
            ci.inactive = None
        ref.sea_ice_state.concentration[:] = 0.9
        dev.sea_ice_state.concentration.fill_(0.9)
        ref.sea_ice_state.hi[:] = 1.2; dev.sea_ice_state.hi.fill_(1.2)
        ref.sea_ice_state.hs[:] = 0.1; dev.sea_ice_state.hs.fill_(0.1)
    ref.initialize(); dev.initialize()
    ref.interpolate_state(T_STEP); dev.interpolate_state(T_STEP)
    ref.compute_atmosphere_sea_ice_fluxes(); dev.compute_atmosphere_sea_ice_fluxes()
    cuda_backend.synchronize()
    g = ref.grid
    ri, di = g.interior(ref.asi_iterations), g.interior(cuda_backend.to_numpy(dev.asi_iterations))
    if case == "fixed_iterations":
        assert np.array_equal(ri, di) and ri.max() == 6
    else:
        assert float((ri != di).mean()) <= 2e-3
    mask = (ri < 100) & (di < 100)
    for n in ref.asi_fluxes.names():
        a, b = g.interior(getattr(ref.asi_fluxes, n)), g.interior(cuda_backend.to_numpy(getattr(dev.asi_fluxes, n)))
        s = float(np.abs(a).max()) or 1.0
        assert float(np.abs(a - b)[mask].max()) / s <= (F64_TOL if case != "fixed_iterations" else 1e-9), f"{case}/{n}"
    Tr, Td = g.interior(ref.sea_ice_state.top_temperature), g.interior(cuda_backend.to_numpy(dev.sea_ice_state.top_temperature))
    assert np.abs(Tr - Td)[mask].max() <= 1e-8


def test_float32_sea_ice_fast_path_against_generic_kernel_and_oracle(oracle_lib, cuda_backend, cuda_lib, monkeypatch):
    """Float32 OceanSeaIce model: the a-si default tree on the work-queue kernel (Float32 skin temperature and q_sat,
    mixed-precision similarity step, tables fitted to the Float32-rounded SHEBA / Paulson parameters) against the generic
    kernel and the oracle.  Bar: 1e-5 of the field scale on points that converge in all three."""
    ref, dev = build_pair("C2", oracle_lib, cuda_backend, FT="f32", atm_FT="f32", sea_ice=True)
    ref.initialize(); dev.initialize()
    ref.interpolate_state(T_STEP); dev.interpolate_state(T_STEP)
    g = dev.grid
    Ts0 = dev.sea_ice_state.top_temperature.clone()
    ref.compute_atmosphere_sea_ice_fluxes()

    def run():
        dev.sea_ice_state.top_temperature.copy_(Ts0)
        dev.compute_atmosphere_sea_ice_fluxes()
        cuda_backend.synchronize()
        out = {n: g.interior(cuda_backend.to_numpy(getattr(dev.asi_fluxes, n))).astype(np.float64) for n in dev.asi_fluxes.names()}
        out["Ts"] = g.interior(cuda_backend.to_numpy(dev.sea_ice_state.top_temperature)).astype(np.float64)
        return out, g.interior(cuda_backend.to_numpy(dev.asi_iterations)).astype(int)

    fast, fit = run()
    monkeypatch.setenv("NE_B200_FORCE_GENERIC", "1")
    gen, git = run()
    monkeypatch.delenv("NE_B200_FORCE_GENERIC")
    oit = g.interior(ref.asi_iterations).astype(int)
    assert any((fast[n] != gen[n]).any() for n in fast), "the Float32 a-si fast path did not run (results identical to the generic kernel's)"
    assert ((fit > 0) == (oit > 0)).all() and (fit > 0).sum() > 10000
    conv = (fit < 100) & (git < 100) & (oit < 100) & (fit > 0)
    assert conv.mean() > 0.02
    for n in dev.asi_fluxes.names():
        o = g.interior(getattr(ref.asi_fluxes, n)).astype(np.float64)
        s = float(np.abs(o).max()) or 1.0
        assert np.isfinite(fast[n]).all(), n
        assert np.abs(fast[n] - gen[n])[conv].max() / s <= F32_TOL, f"{n}: fast vs generic {np.abs(fast[n] - gen[n])[conv].max() / s}"
        assert np.abs(fast[n] - o)[conv].max() / s <= F32_TOL, f"{n}: fast vs oracle {np.abs(fast[n] - o)[conv].max() / s}"
        assert np.abs(fast[n] - o).max() / s <= 2e-2, f"{n}: limit-cycle points {np.abs(fast[n] - o).max() / s}"
    To = g.interior(ref.sea_ice_state.top_temperature).astype(np.float64)
    assert np.abs(fast["Ts"] - To)[conv].max() <= 2e-3
    ice = oit > 0
    assert abs(float(fit[ice].mean()) - float(oit[ice].mean())) <= 3.0
    assert abs(float((fit[ice] == 100).mean()) - float((oit[ice] == 100).mean())) <= 0.05


@pytest.mark.parametrize("FT", ["f64", "f32"])
def test_interface_step_replays_from_a_cuda_graph(cuda_backend, cuda_lib, FT):
    """The whole step (merged interpolation, solve — the persistent work-queue kernel for Float32 —, post-solve kernel with
    diagnostics, and the a-si work-queue kernel) is capturable: no allocation, memset or synchronisation after the first
    call, the queue kernels re-arm their own tile counters.  Replaying the graph reproduces the eager results bit for bit."""
    import torch
    from numericalearth_jl_b200 import sharding
    dev = synthetic.build_case("C1", cuda_backend, FT=FT, atm_FT="f32", sea_ice=True)
    dev.initialize()
    f = dev.ao_fluxes
    diag = sharding.FluxDiagnostics(dev, [f.latent_heat, f.sensible_heat, f.x_momentum, dev.net_ocean.T])
    Ts0 = dev.sea_ice_state.top_temperature.clone()

    def step():
        dev.sea_ice_state.top_temperature.copy_(Ts0)
        dev.fused_interface_step(T_STEP, diagnostics=diag)
        dev.compute_atmosphere_sea_ice_fluxes()

    def snapshot():
        out = {"diag": diag.result.clone()}
        for bag in ("ao_fluxes", "net_ocean", "asi_fluxes"):
            for n in getattr(dev, bag).names():
                out[bag + "." + n] = getattr(getattr(dev, bag), n).clone()
        out["Ts"] = dev.sea_ice_state.top_temperature.clone()
        return out

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):       # warm-up on the capture stream: builds the solver tables, allocates the counter pool
        step(); step()
        torch.cuda.synchronize()
        eager = snapshot()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            step()
        for bag in ("ao_fluxes", "net_ocean", "asi_fluxes"):
            for n in getattr(dev, bag).names():
                getattr(getattr(dev, bag), n).fill_(float("nan"))
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        replayed = snapshot()
    g = dev.grid
    for k, v in eager.items():
        if k == "diag":
            assert torch.equal(v, replayed[k]), k
            continue
        # the kernels write their launch range only: (0:N+1) for the flux kernels, (1:N) for the net fluxes
        win = (slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)) if k.startswith("net_ocean") else \
            (slice(g.hy - 1, g.hy + g.ny + 1), slice(g.hx - 1, g.hx + g.nx + 1))
        a, b = v[win], replayed[k][win]
        assert not torch.isnan(b).any(), f"{k}: the replay did not write its launch range"
        assert torch.equal(a, b), k
