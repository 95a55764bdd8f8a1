"""Differential test of the C++ oracle against the INDEPENDENT arbitrary-precision restatement oracle/pin/reference_mp.py
(mpmath, 50 digits, written from the reference's documentation and docstring mathematics, sharing no text with
oracle/ne_oracle.cpp or the CUDA headers).  A misread constant, sign or formula in the oracle would show up here as an
O(1) difference; what remains must be the oracle's Float64 rounding: a few ulp per function, 1e-12 on whole fixed points,
equal trip counts.

Default sizes keep the CPU suite fast (NE_PIN_POINTS=100000 NE_PIN_SOLVES=5000 runs the full differential test the
VERDICT asks for; its result is recorded in profiles/r02_oracle_pin.json)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "pin"))
import mpmath as mp  # noqa: E402
import reference_mp as R  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import formulations as F  # noqa: E402

N_POINTS = int(os.environ.get("NE_PIN_POINTS", "4000"))
N_SOLVES = int(os.environ.get("NE_PIN_SOLVES", "150"))
RECORD = os.environ.get("NE_PIN_RECORD")      # path of a JSON file to append the measured errors to


def _ulps(got, exact):
    """|got - exact| in units of the spacing of doubles at |exact| (0 when both are 0)."""
    exact_f = float(exact)
    if exact_f == 0.0:
        return abs(got) / 5e-324 if got != 0 else 0.0
    return float(abs(mp.mpf(got) - exact) / mp.mpf(np.spacing(abs(exact_f))))


def _record(name, **kw):
    row = dict(check=name, **kw)
    print("PIN", json.dumps(row))
    if RECORD:
        with open(RECORD, "a") as f:
            f.write(json.dumps(row) + "\n")


def _psi_pod(kind):
    fn = F.EdsonMomentumStabilityFunction() if kind == "momentum" else F.EdsonScalarStabilityFunction()
    return F.stability_profile_pod(fn)


@pytest.mark.parametrize("kind", ["momentum", "scalar"])
def test_edson_stability_functions(oracle_lib, kind):
    """ψ(ζ) over the whole range the solve visits: ζ from −1e5 (second trip of an unstable point) to +60 (beyond ζmax/A⁺)."""
    rng = np.random.default_rng(101 if kind == "momentum" else 102)
    n = N_POINTS
    mag = 10.0 ** rng.uniform(-9, 5, n)
    z = np.where(rng.random(n) < 0.5, -mag, np.minimum(mag, 10.0 ** rng.uniform(-9, np.log10(200.0), n)))
    z[:4] = [0.0, -0.0, 1e-300, -1e-300]
    pod = _psi_pod(kind)
    exact_fn = R.psi_momentum if kind == "momentum" else R.psi_scalar
    worst_ulp, worst_abs = 0.0, 0.0
    for zz in z:
        got = oracle_lib.dll.neo_stability_f64(C.byref(pod), float(zz))
        ex = exact_fn(float(zz))
        err = abs(mp.mpf(got) - ex)
        # ψ passes through 0 near ζ = 0 and has cancellation between its terms: bound the error in units of the largest
        # term's spacing, i.e. relative to max(1, |ψ|)
        rel = float(err / max(mp.mpf(1), abs(ex)))
        worst_abs = max(worst_abs, rel)
        worst_ulp = max(worst_ulp, rel / 2.220446049250313e-16)
    _record(f"psi_{kind}", points=n, worst_error_rel_to_max1=worst_abs, worst_in_eps=worst_ulp)
    assert worst_ulp <= 32.0, worst_ulp     # |ψ| and its terms reach ~10 and each term carries the rounding of its argument: a few tens of eps(1)


OTHER_PSI = {
    # name: (formulation object handed to the oracle, stable φ, unstable φ, range of ζ sampled)
    "sheba_momentum": (lambda: F.ShebaMomentumStabilityFunction(), R.phi_sheba_momentum, None, (1e-6, 150.0)),
    "sheba_scalar": (lambda: F.ShebaScalarStabilityFunction(), R.phi_sheba_scalar, None, (1e-6, 150.0)),
    "paulson_momentum": (lambda: F.PaulsonMomentumStabilityFunction(), None, R.phi_businger_dyer_momentum, (1e-6, 1e5)),
    "paulson_scalar": (lambda: F.PaulsonScalarStabilityFunction(), None, R.phi_businger_dyer_scalar, (1e-6, 1e5)),
    "linear_stable": (lambda: F.LinearStableStabilityFunction(), R.phi_linear_stable, None, (1e-6, 40.0)),
    "sea_ice_momentum": (lambda: F.atmosphere_sea_ice_stability_functions().momentum, R.phi_sheba_momentum, R.phi_businger_dyer_momentum, (1e-6, 1e3)),
    "sea_ice_scalar": (lambda: F.atmosphere_sea_ice_stability_functions().temperature, R.phi_sheba_scalar, R.phi_businger_dyer_scalar, (1e-6, 1e3)),
    "large_yeager_momentum": (lambda: F.large_yeager_stability_functions().momentum, R.phi_linear_stable, R.phi_businger_dyer_momentum, (1e-6, 1e3)),
    "large_yeager_scalar": (lambda: F.large_yeager_stability_functions().temperature, R.phi_linear_stable, R.phi_businger_dyer_scalar, (1e-6, 1e3)),
}


@pytest.mark.parametrize("name", sorted(OTHER_PSI))
def test_other_stability_functions_against_their_flux_profile_integrals(oracle_lib, name):
    """SHEBA (Grachev et al. 2007), Paulson (1970), the linear stable function of Large & Yeager (2004) and the two shipped
    SplitStabilityFunction pairs (sea ice, land / NCAR): the oracle's closed forms against ψ(ζ) = ∫₀^ζ (1 − φ(x))/x dx evaluated
    numerically at 50 digits from the papers' φ — an antiderivative mis-copied from the reference cannot pass this."""
    make, stable, unstable, (lo, hi) = OTHER_PSI[name]
    pod = F.stability_profile_pod(make())
    rng = np.random.default_rng(abs(hash(name)) % 1000 + 7)
    n = max(60, N_POINTS // 400)
    mag = 10.0 ** rng.uniform(np.log10(lo), np.log10(hi), n)
    sign = np.where(rng.random(n) < 0.5, 1.0, -1.0)
    if stable is None:
        sign[: n * 3 // 4] = -1.0        # mostly the side that is not identically zero
    if unstable is None:
        sign[: n * 3 // 4] = 1.0
    worst = 0.0
    for zz in (sign * mag).tolist() + [0.0]:
        got = oracle_lib.dll.neo_stability_f64(C.byref(pod), float(zz))
        ex = R.psi_split(stable, unstable, float(zz))
        if name == "linear_stable" and zz < 0:
            ex = mp.mpf(0)
        rel = float(abs(mp.mpf(got) - ex) / max(mp.mpf(1), abs(ex)))
        worst = max(worst, rel / 2.220446049250313e-16)
    _record(f"psi_{name}_vs_integral", points=n + 1, worst_in_eps=worst)
    assert worst <= 64.0, worst       # closed forms with logs / arctangents of O(1..10) arguments: a few tens of eps(1)


def test_saturation_vapor_pressure_and_surface_humidity(oracle_lib):
    rng = np.random.default_rng(103)
    th = F.AtmosphereThermodynamicsParameters(FT="f64").pod()
    ip = F.InterfaceProperties(F.ImpureSaturationSpecificHumidity(F.Liquid(), 0.98), F.BulkTemperature(), F.RelativeVelocity()).pod()
    ip_ice = F.InterfaceProperties(F.ImpureSaturationSpecificHumidity(F.Ice(), None), F.BulkTemperature(), F.RelativeVelocity()).pod()
    w_p, w_q, w_qi = 0.0, 0.0, 0.0
    for _ in range(N_POINTS):
        T = float(rng.uniform(230.0, 320.0)); p = float(rng.uniform(5e4, 1.1e5))
        for ice in (False, True):
            got = oracle_lib.dll.neo_saturation_vapor_pressure_f64(C.byref(th), T, 1 if ice else 0)
            w_p = max(w_p, _ulps(got, R.p_sat(T, ice=ice)))
        got = oracle_lib.dll.neo_surface_specific_humidity_f64(C.byref(ip), C.byref(th), p, T, 35.0)
        w_q = max(w_q, _ulps(got, R.q_surface(p, T, 0.98)))
        got = oracle_lib.dll.neo_surface_specific_humidity_f64(C.byref(ip_ice), C.byref(th), p, T, 0.0)
        w_qi = max(w_qi, _ulps(got, R.q_surface(p, T, 1.0, ice=True)))
    _record("p_sat / q_surface", points=N_POINTS, p_sat_ulp=w_p, q_liquid_ulp=w_q, q_ice_ulp=w_qi)
    # pow(T/T_tr, Δcp/R_v) and exp(…) each carry the rounding of their argument times |argument| ≲ 25
    assert w_p <= 64 and w_q <= 64 and w_qi <= 64, (w_p, w_q, w_qi)


def test_roughness_lengths_and_gustiness(oracle_lib):
    rng = np.random.default_rng(104)
    P = R.SolverParams()
    ff = F.SimilarityTheoryFluxes()
    pod = F.flux_formulation_pod(ff)
    w_m, w_s, w_g = 0.0, 0.0, 0.0
    for _ in range(N_POINTS):
        us = float(10.0 ** rng.uniform(-6, 0.7)); U = float(rng.uniform(0.0, 40.0))
        got = oracle_lib.dll.neo_momentum_roughness_f64(C.byref(pod.ell_momentum), us, U)
        ex = R.ell_momentum(us, P)
        w_m = max(w_m, _ulps(got, ex))
        lu = float(ex)
        got = oracle_lib.dll.neo_scalar_roughness_f64(C.byref(pod.ell_temperature), lu, us)
        w_s = max(w_s, _ulps(got, R.ell_scalar(lu, us, P)))
        bs = float(rng.normal(0, 0.01)); hbl = float(rng.uniform(100, 2000))
        got = oracle_lib.dll.neo_vsgs2_f64(C.byref(pod.subgrid_velocities), us, bs, hbl)
        w_g = max(w_g, _ulps(got, R.gustiness_squared(us, bs, hbl, P)))
    _record("roughness / gustiness", points=N_POINTS, ell_momentum_ulp=w_m, ell_scalar_ulp=w_s, gustiness_squared_ulp=w_g)
    assert w_m <= 4 and w_s <= 8 and w_g <= 16, (w_m, w_s, w_g)   # U_G²: cbrt of a product, times β, squared


def test_interpolator(oracle_lib):
    rng = np.random.default_rng(105)
    im, ip, xi = C.c_int64(), C.c_int64(), C.c_double()
    worst = 0.0
    for _ in range(N_POINTS):
        f = float(rng.uniform(-2.0, 700.0)) if rng.random() < 0.9 else float(rng.integers(-2, 700))
        oracle_lib.dll.neo_interpolator_f64(f, C.byref(im), C.byref(ip), C.byref(xi))
        a, b, x = R.interpolator(f)
        if f >= 0:   # (the reference never sees a negative fractional index: the source grid's halo covers f ≥ −0.5 → clamped by the wrap)
            assert (im.value, ip.value) == (a, b), f
            worst = max(worst, abs(xi.value - float(x)))
    _record("interpolator", points=N_POINTS, xi_max_abs_err=worst)
    assert worst == 0.0     # f mod 1 is exact in binary floating point


def _random_point(rng):
    """One air-sea pair spanning calm / gale, stable / unstable, tropical / polar."""
    To = float(rng.uniform(271.5, 303.0))
    Ta = To + float(rng.normal(0.0, 4.0))
    p = float(rng.uniform(9.6e4, 1.04e5))
    wind = float(10.0 ** rng.uniform(-1.2, 1.5))
    ang = float(rng.uniform(0, 2 * np.pi))
    qsat_a = float(R.q_surface(p, Ta, 1.0))
    qa = float(rng.uniform(0.4, 0.98)) * qsat_a
    atm = (wind * np.cos(ang), wind * np.sin(ang), Ta, p, qa)
    ocean = (float(rng.normal(0, 0.1)), float(rng.normal(0, 0.1)), To, float(rng.uniform(30, 38)))
    return atm, ocean


def test_whole_fixed_points_and_fluxes(oracle_lib, host_backend):
    """N_SOLVES random air-sea pairs through the oracle's kernel (one grid point each) and through the independent
    restatement with the same stopping rule: trip counts equal, (u★, θ★, q★) and the five fluxes to 1e-12."""
    rng = np.random.default_rng(106)
    n = N_SOLVES
    pts = [_random_point(rng) for _ in range(n)]
    g = ne_b200.ExchangeGrid(nx=n, ny=1, hx=1, hy=1, latitude=(-1.0, 1.0), FT="f64")
    ci = ne_b200.ComponentInterfaces(g, host_backend, None, None, lib=oracle_lib, with_iterations=True)
    for k, (atm, ocean) in enumerate(pts):
        for name, v in zip(("u", "v", "T", "p", "q"), atm):
            getattr(ci.atmos_state, name)[:, :] = getattr(ci.atmos_state, name)      # keep dtype/shape
            getattr(ci.atmos_state, name)[1, 1 + k] = v
        # face-located ocean velocities: both faces of the cell equal → the cell-centre average is the value itself
        ci.ocean_state.u[1, 1 + k] = ocean[0]; ci.ocean_state.u[1, 2 + k] = ocean[0]
    # the x-average reads u[i] and u[i+1]: give every cell its own value on both faces by using zero ocean velocity instead
    ci.ocean_state.u[...] = 0.0
    ci.ocean_state.v[...] = 0.0
    for k, (atm, ocean) in enumerate(pts):
        ci.ocean_state.T[1, 1 + k] = ocean[2] - 273.15      # ocean temperature units: degrees Celsius (components.jl:7-15)
        ci.ocean_state.S[1, 1 + k] = ocean[3]
    d = ci.atmosphere_ocean_desc()
    d.grid = g.pod(False)
    oracle_lib.call("atmosphere_ocean_fluxes", "f64", d, 0)
    worst_scale, worst_flux, trip_mismatch = 0.0, 0.0, 0
    for k, (atm, ocean) in enumerate(pts):
        To_K = float(np.float64(ocean[2] - 273.15) + 273.15)   # the Kelvin value the kernel forms
        us, ts, qs, trips, fl = R.solve_point(atm, (0.0, 0.0, To_K, ocean[3]))
        got_trips = int(ci.ao_iterations[1, 1 + k])
        if got_trips != trips:
            trip_mismatch += 1
            continue
        # θ★ ∝ Δθ and q★ ∝ Δq are differences of O(300 K) / O(0.02) numbers: their rounding does not shrink with them, so the
        # denominators are floored at 1e-6 of the scales the criterion of the GPU tests uses (Δθ ~ 1 K → θ★ ~ 0.03, …)
        floors = {"friction_velocity": 1e-9, "temperature_scale": 3e-8, "water_vapor_scale": 3e-10,
                  "latent_heat": 1e-4, "sensible_heat": 1e-4, "water_vapor": 1e-10, "x_momentum": 1e-7, "y_momentum": 1e-7}
        for name, ex in (("friction_velocity", us), ("temperature_scale", ts), ("water_vapor_scale", qs)):
            got = float(getattr(ci.ao_fluxes, name)[1, 1 + k])
            worst_scale = max(worst_scale, float(abs(mp.mpf(got) - ex) / max(abs(ex), mp.mpf(floors[name]))))
        for name, ex in fl.items():
            got = float(getattr(ci.ao_fluxes, name)[1, 1 + k])
            worst_flux = max(worst_flux, float(abs(mp.mpf(got) - ex) / max(abs(ex), mp.mpf(floors[name]))))
    _record("whole fixed points", solves=n, trip_count_mismatches=trip_mismatch, worst_rel_scales=worst_scale, worst_rel_fluxes=worst_flux)
    assert trip_mismatch <= max(1, n // 500), trip_mismatch     # a drift within rounding of tol may flip the last trip
    assert worst_scale <= 1e-10 and worst_flux <= 1e-10, (worst_scale, worst_flux)


def test_closed_forms_of_the_pin_equal_their_integrals():
    """The antiderivatives the sea-ice whole-solve pin uses (reference_mp.psi_closed) against the quadrature of the same φ."""
    pairs = {"sheba_momentum": R.phi_sheba_momentum, "sheba_scalar": R.phi_sheba_scalar,
             "paulson_momentum": R.phi_businger_dyer_momentum, "paulson_scalar": R.phi_businger_dyer_scalar}
    for kind, phi in pairs.items():
        sgn = 1 if kind.startswith("sheba") else -1
        for mag in (1e-5, 3e-3, 0.2, 1.7, 30.0, 400.0):
            z = sgn * mag
            a, b = R.psi_closed(kind, z), R.psi_from_phi(phi, z)
            assert abs(a - b) <= mp.mpf("1e-25") * max(1, abs(b)), (kind, z, a, b)
        assert R.psi_closed(kind, -sgn * 2.0) == 0


def test_sea_ice_whole_fixed_points(oracle_lib, host_backend):
    """Ice-covered points of a synthetic OceanSeaIce case through the oracle's a–si kernel and through the independent
    restatement (skin-temperature balance with linearised long wave, ice-phase humidity, SHEBA / Paulson profiles, the same
    stopping rule): trip counts, the skin temperature and the five fluxes."""
    from numericalearth_jl_b200 import synthetic
    cfg = dict(nx=72, ny=40, latitude=(-80.0, 80.0), src_nx=64, src_ny=32)
    ci = synthetic.build_case(cfg, host_backend, FT="f64", atm_FT="f64", sea_ice=True, lib=oracle_lib, with_iterations=True)
    ci.initialize()
    ci.interpolate_state(0.37 * 10800.0)
    g = ci.grid
    T0 = np.array(ci.sea_ice_state.top_temperature, copy=True)
    ci.compute_atmosphere_sea_ice_fluxes()
    conc, inactive = np.asarray(ci.sea_ice_state.concentration), np.asarray(ci.inactive)
    it = np.asarray(ci.asi_iterations)
    cand = [(j, i) for j in range(g.hy, g.hy + g.ny) for i in range(g.hx, g.hx + g.nx) if conc[j, i] > 0 and not inactive[j, i]]
    assert len(cand) > 50
    rng = np.random.default_rng(9)
    picks = [cand[k] for k in rng.choice(len(cand), size=min(40, len(cand)), replace=False)]
    a, r, si = ci.atmos_state, ci.rad_state, ci.sea_ice_state
    worst_T = worst_flux = 0.0
    mismatch = maxed = 0
    for (j, i) in picks:
        atm = tuple(float(x[j, i]) for x in (a.u, a.v, a.T, a.p, a.q, r.sw, r.lw))
        ice = (float(ci.ocean_state.S[j, i]), float(si.hi[j, i]), float(si.hc[j, i]))
        Ts0 = float(np.float64(T0[j, i]) + 273.15)
        us, ts, qs, Ts, trips, fl = R.solve_sea_ice_point(atm, ice, Ts0)
        if int(it[j, i]) != trips:
            mismatch += 1
            continue
        if trips >= 100:          # a limit cycle of the rounded iterate: the two orbits need not be in phase
            maxed += 1
            continue
        got_T = float(np.float64(si.top_temperature[j, i]) + 273.15)
        worst_T = max(worst_T, float(abs(mp.mpf(got_T) - Ts) / Ts))
        floors = {"latent_heat": 1e-3, "sensible_heat": 1e-3, "water_vapor": 1e-9, "x_momentum": 1e-6, "y_momentum": 1e-6}
        for name, ex in fl.items():
            got = float(getattr(ci.asi_fluxes, name)[j, i])
            worst_flux = max(worst_flux, float(abs(mp.mpf(got) - ex) / max(abs(ex), mp.mpf(floors[name]))))
    _record("sea-ice whole fixed points", solves=len(picks), trip_count_mismatches=mismatch, at_maxiter=maxed,
            worst_rel_skin_temperature=worst_T, worst_rel_fluxes=worst_flux)
    assert mismatch <= 2, mismatch
    assert len(picks) - mismatch - maxed >= 20
    assert worst_T <= 1e-12 and worst_flux <= 1e-9, (worst_T, worst_flux)
