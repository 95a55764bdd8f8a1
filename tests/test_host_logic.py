"""Host-side mirror of the reference's plugin types: defaults, POD translation, error behaviour,
time-interpolation indices, exchange-grid layout."""
import math

import numpy as np
import pytest

import ne_b200
from ne_b200 import abi as A

F = ne_b200


def test_similarity_theory_defaults_match_reference():
    """similarity_theory_turbulent_fluxes.jl:174-214 + Appendix A of SURVEY.md."""
    f = F.SimilarityTheoryFluxes()
    p = f.pod()
    assert p.kind == A.NE_FLUX_SIMILARITY_THEORY and p.similarity_form == A.NE_PROFILE_LOGARITHMIC
    assert p.von_karman_constant == 0.4
    assert p.stop.kind == A.NE_STOP_CONVERGENCE and p.stop.tolerance == 1e-8 and p.stop.maxiter == 100
    assert p.subgrid_velocities.convective_kind == A.NE_SGS_CONVECTIVE
    assert (p.subgrid_velocities.gustiness_parameter, p.subgrid_velocities.minimum_gustiness) == (1.2, 0.01)
    m = p.ell_momentum
    assert m.kind == A.NE_ROUGH_MOMENTUM and m.wave_kind == A.NE_WAVE_CONSTANT
    assert (m.wave_constant, m.smooth_wall_parameter, m.maximum_roughness_length, m.nu) == (0.02, 0.11, 1.0, 1.5e-5)
    assert m.gravitational_acceleration == 9.80665 and m.visc_dtype == A.NE_F64
    s = p.ell_temperature
    assert s.kind == A.NE_ROUGH_SCALAR and (s.reynolds_A, s.reynolds_b, s.maximum_roughness_length) == (5.85e-5, 0.72, 1.6e-4)
    assert bytes(p.ell_temperature) == bytes(p.ell_water_vapor)
    em = p.psi_momentum.a
    assert em.kind == A.NE_PSI_EDSON_MOMENTUM
    assert list(em.p)[:11] == [50.0, 0.35, 0.7, 0.75, 5 / 0.35, 15.0, 2.0, math.pi / 2, 10.15, 3.0, math.pi / math.sqrt(3)]
    es = p.psi_temperature.a
    assert es.kind == A.NE_PSI_EDSON_SCALAR
    assert list(es.p) == [50.0, 0.35, 2 / 3, 1.5, 14.28, 8.525, 15.0, 2.0, 0.0, 34.15, 3.0, math.pi / math.sqrt(3)]


def test_sea_ice_and_large_yeager_presets():
    p = F.atmosphere_sea_ice_similarity_theory().pod()      # :779-794
    assert p.psi_momentum.split == 1 and p.psi_momentum.a.kind == A.NE_PSI_SHEBA_MOMENTUM
    assert p.psi_momentum.b.kind == A.NE_PSI_PAULSON_MOMENTUM
    assert p.psi_temperature.a.kind == A.NE_PSI_SHEBA_SCALAR and p.psi_temperature.b.kind == A.NE_PSI_PAULSON_SCALAR
    assert list(p.psi_momentum.a.p)[:2] == [6.5, 1.3] and list(p.psi_temperature.a.p)[:3] == [5.0, 5.0, 3.0]
    ly = F.CoefficientBasedFluxes(transfer_coefficients=F.LargeYeagerTransferCoefficients()).pod()
    assert ly.kind == A.NE_FLUX_LARGE_YEAGER
    assert (ly.large_yeager.reference_height, ly.large_yeager.stable_heat, ly.large_yeager.unstable_heat,
            ly.large_yeager.moisture) == (10, 18, 32.7, 34.6)
    assert ly.large_yeager.neutral_drag.c == 1 / 13.09 and ly.stop.maxiter == 20


def test_stability_functions_none_maps_to_zero():
    p = F.SimilarityTheoryFluxes(stability_functions=None).pod()
    assert p.psi_momentum.a.kind == A.NE_PSI_ZERO and p.psi_temperature.a.kind == A.NE_PSI_ZERO


@pytest.mark.parametrize("bad", [
    lambda: F.SimilarityTheoryFluxes(momentum_roughness_length=lambda u, *a: 1e-4).pod(),
    lambda: F.SimilarityTheoryFluxes(stability_functions=F.SimilarityScales(lambda z: 0.0, None, None)).pod(),
    lambda: F.SimilarityTheoryFluxes(subgrid_velocities=object()).pod(),
    lambda: F.SimilarityTheoryFluxes(solver_stop_criteria="never").pod(),
    lambda: F.CoefficientBasedFluxes(transfer_coefficients=(lambda *a: 1e-3, 1e-3, 1e-3)).pod(),
    lambda: F.InterfaceProperties(temperature_formulation=object()).pod(),
    lambda: F.InterfaceProperties(specific_humidity_formulation=object()).pod(),
    lambda: F.MomentumRoughnessLength(air_kinematic_viscosity=lambda T: 1e-5).pod(),
])
def test_user_closures_raise_no_kernel_variant(bad):
    """A user-defined closure with no kernel variant raises; there is no CPU fallback (north_star)."""
    with pytest.raises(F.NoKernelVariantError):
        bad()


def test_interface_properties_pod():
    p = F.InterfaceProperties().pod()
    assert p.phase == A.NE_PHASE_LIQUID and p.x_h2o_kind == A.NE_XH2O_CONSTANT and p.x_h2o == 0.98
    assert p.velocity_formulation == A.NE_VEL_RELATIVE and p.temperature_formulation == A.NE_TEMP_BULK
    p = F.InterfaceProperties(F.ImpureSaturationSpecificHumidity(F.Ice(), None),
                              F.SkinTemperature(F.IceSnowConductiveFlux(0.31, 2.0)), F.WindVelocity()).pod()
    assert p.phase == A.NE_PHASE_ICE and p.x_h2o_kind == A.NE_XH2O_ONE and p.velocity_formulation == A.NE_VEL_WIND
    assert p.temperature_formulation == A.NE_TEMP_SKIN_ICE_SNOW and p.max_dT == 5
    assert (p.ice_conductivity, p.snow_conductivity) == (2.0, 0.31)
    p = F.InterfaceProperties(temperature_formulation=F.SkinTemperature(F.DiffusiveFlux(F.InteriorDiffusivity(), 0.5))).pod()
    assert p.temperature_formulation == A.NE_TEMP_SKIN_DIFFUSIVE_INTERIOR and p.kappa == 1.4e-7 and p.delta == 0.5


def test_thermodynamics_defaults():
    """src/Atmospheres/thermodynamic_parameters.jl:45-200."""
    t = F.AtmosphereThermodynamicsParameters().pod()
    assert (t.gas_constant, t.dry_air_molar_mass, t.water_molar_mass) == (8.3144598, 0.02897, 0.018015)
    assert (t.kappa_d, t.cp_v, t.cp_l, t.cp_i) == (2 / 7, 1859, 4181, 2100)
    assert (t.LH_v0, t.LH_s0, t.T_0, t.T_triple, t.press_triple, t.T_freeze, t.T_icenuc) == \
        (2500800, 2834400, 273.16, 273.16, 611.657, 273.15, 233)
    assert F.AtmosphereThermodynamicsParameters(FT="f32").pod().dtype == A.NE_F32


def test_three_equation_defaults():
    f = F.ThreeEquationHeatFlux()       # sea_ice_ocean_heat_flux_formulations.jl:150-159
    assert f.heat_transfer_coefficient == 0.0095 and f.salt_transfer_coefficient == 0.0095 / 35
    assert f.friction_velocity == 0.002
    b = F.IceBathHeatFlux()
    assert (b.heat_transfer_coefficient, b.friction_velocity) == (0.006, 0.02)


def test_interpolating_time_indices():
    times = np.arange(4) * 10800.0
    f = ne_b200.interpolating_time_indices
    assert f(times, 0.0) == (0.0, 1, 2)
    nt, n1, n2 = f(times, 0.37 * 10800.0)
    assert (n1, n2) == (1, 2) and nt == pytest.approx(0.37)
    assert f(times, 10800.0) == (0.0, 2, 3)
    nt, n1, n2 = f(times, 3.5 * 10800.0)           # cyclical wrap: between the last and the first snapshot
    assert (n1, n2) == (4, 1) and nt == pytest.approx(0.5)
    nt, n1, n2 = f(times, 4 * 10800.0 + 100.0)     # period = (tN - t1) + Δt
    assert (n1, n2) == (1, 2) and nt == pytest.approx(100.0 / 10800.0)
    assert f(times, -5.0, "clamp") == (0.0, 1, 1)
    assert f(times, 1e9, "clamp") == (0.0, 4, 4)
    assert f(np.array([0.0]), 123.0) == (0.0, 1, 1)


def test_exchange_grid_layout_and_launch_ranges():
    g = ne_b200.ExchangeGrid(nx=360, ny=150, hx=7, hy=7, latitude=(-75.0, 75.0))
    assert g.shape == (164, 374) and g.launch_points() == 55024 and g.launch_points(False) == 54000
    p = g.pod(True)
    assert (p.i_lo, p.i_hi, p.j_lo, p.j_hi) == (0, 361, 0, 151)
    p = g.pod(False)
    assert (p.i_lo, p.i_hi, p.j_lo, p.j_hi) == (1, 360, 1, 150)
    assert g.lam[g.hx] == pytest.approx(0.5) and g.phi[g.hy] == pytest.approx(-74.5)
    assert g.interior(np.zeros(g.shape)).shape == (152, 362)
    assert ne_b200.ExchangeGrid(nx=1440, ny=560).launch_points() == 810404
    assert ne_b200.ExchangeGrid(nx=4320, ny=1680).launch_points() == 7269604


def test_cuda_library_refuses_host_arrays(host_backend):
    """The product library only takes device arrays: pairing it with host arrays fails loudly."""
    g = ne_b200.ExchangeGrid(nx=4, ny=4, hx=2, hy=2)
    with pytest.raises(RuntimeError):
        ne_b200.ComponentInterfaces(g, host_backend, None, None, lib=ne_b200.Library())


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to the CUDA arm) needs no GPU: one JSON line on
    stdout carrying the contract keys, the oracle as `cpu_baseline` and an e2e block with zero copies."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "air-sea flux points/s" and d["unit"] == "points/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C4")


def test_prescribed_radiation_defaults_match_reference():
    """test/test_radiations.jl:20-24: ocean albedo 0.05 / emissivity 0.97, sea ice 0.7 / 1.0, σ ≈ 5.67e-8 (atol 1e-10);
    :63-79: a land surface given as numbers is read back unchanged."""
    src = ne_b200.LatLonSourceGrid(nx=8, ny=4)
    rad = ne_b200.PrescribedRadiation(grid=src, times=[0.0, 1.0])
    assert rad.surface_properties["ocean"].albedo == 0.05 and rad.surface_properties["ocean"].emissivity == 0.97
    assert rad.surface_properties["sea_ice"].albedo == 0.7 and rad.surface_properties["sea_ice"].emissivity == 1.0
    assert abs(rad.stefan_boltzmann_constant - 5.67e-8) < 1e-10
    assert set(rad.surface_properties) == {"ocean", "sea_ice"}
    land = ne_b200.SurfaceRadiationProperties(0.15, 0.93)
    rad2 = ne_b200.PrescribedRadiation(grid=src, times=[0.0, 1.0], surface_properties={"land": land})
    assert rad2.surface_properties["land"].albedo == 0.15 and rad2.surface_properties["land"].emissivity == 0.93
    assert rad.window is None and rad.time_indexing == "cyclical"
