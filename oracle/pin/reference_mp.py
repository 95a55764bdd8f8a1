"""INDEPENDENT PIN of the oracle — TEST INFRASTRUCTURE ONLY (like everything under oracle/).

A second restatement of the reference's similarity-theory flux path, written from the reference's own documentation
and docstring mathematics — not from oracle/ne_oracle.cpp and not from the CUDA headers — in arbitrary precision
(mpmath, 50 significant digits), with a different program structure (closures over a parameter record, one generic
fixed-point driver).  It exists so that the C++ oracle is checked against something its author could not have
mis-copied the same way twice: tests/test_oracle_independent_pin.py differential-tests every function below against
the oracle on random arguments and whole fixed points.

Sources (all relative to /root/reference):
  saturation vapour pressure     docs/src/interface_fluxes.md:84-96  (Clausius–Clapeyron with constant Δc_p)
  surface specific humidity      src/EarthSystemModels/InterfaceComputations/interface_states.jl:44-74
                                 (q = ε p⁺ / (p − (1 − ε) p⁺), ε = R_d/R_v, p⁺ = min(x p_sat, 0.999 p))
  mixture gas constant, b★       docs/src/interface_fluxes.md:500-545 ; similarity_theory_turbulent_fluxes.jl:389-425
  roughness lengths              docs/src/interface_fluxes.md:236-262 ; roughness_lengths.jl:197-246 (Edson 2013 eq. 28)
  gustiness                      similarity_theory_turbulent_fluxes.jl:36-48, 88-98  (U_G = max(floor, β (J_b h_bl)^{1/3}))
  Edson et al. (2013) ψ_u, ψ_θ   docstring mathematics at similarity_theory_turbulent_fluxes.jl:450-486, 534-570
  SHEBA / Paulson / linear ψ     the flux–profile relations φ of Grachev et al. (2007), Paulson (1970), Large & Yeager (2004),
                                 integrated numerically: ψ(ζ) = ∫₀^ζ (1 − φ)/x dx (no closed form shared with the reference)
                                 (= COARE 3.5 psiu_26 / psit_26) with the constants of :487-499, 571-584
  similarity profile, χ          docs/src/interface_fluxes.md:620-690 ; similarity_theory_turbulent_fluxes.jl:239-253, 375-384
  fixed point + stopping rule    compute_interface_state.jl:5-58
  flux epilogue                  atmosphere_ocean_fluxes.jl:160-196
  interpolator / bilinear / time docs of Oceananigans (third party): i⁻ = trunc(f) + 1, ξ = f mod 1; ψ₂ ñ + ψ₁ (1 − ñ)

Every input is taken as the exact value of the double it was given as; every result is exact to ~1e-45, so a
difference from the oracle IS the oracle's rounding error (or its mistake).
"""
from dataclasses import dataclass, field

import mpmath as mp

mp.mp.dps = 50
M = mp.mpf


def _m(x):
    return M(float(x)) if not isinstance(x, mp.mpf) else x


# ----------------------------------------------------------------------------------------------------------------
# parameters (Appendix A of SURVEY.md = the reference's keyword defaults; cited where each record is used)
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class Thermo:   # src/Atmospheres/thermodynamic_parameters.jl:30-258
    R: float = 8.3144598
    M_d: float = 0.02897
    M_v: float = 0.018015
    kappa_d: float = 2.0 / 7.0
    cp_v: float = 1859.0
    cp_l: float = 4181.0
    cp_i: float = 2100.0
    LH_v0: float = 2500800.0
    LH_s0: float = 2834400.0
    T_0: float = 273.16
    T_tr: float = 273.16
    p_tr: float = 611.657

    # the reference forms these in Float64 from the Float64 fields: keep the SAME doubles as inputs, exact afterwards
    @property
    def R_d(self):
        return _m(self.R / self.M_d)

    @property
    def R_v(self):
        return _m(self.R / self.M_v)

    @property
    def cp_d(self):
        return _m((self.R / self.M_d) / self.kappa_d)

    @property
    def eps_vd(self):       # R_v / R_d = M_d / M_v
        return _m(self.M_d / self.M_v)


@dataclass
class EdsonConstants:   # similarity_theory_turbulent_fluxes.jl:487-499 (momentum), 571-584 (scalar)
    zmax: float = 50.0
    Ap: float = 0.35
    Bp_m: float = 0.7
    Cp_m: float = 0.75
    Dp_m: float = 5 / 0.35
    Am: float = 15.0
    Bm: float = 2.0
    Cm_m: float = float(mp.pi / 2)
    Dm_m: float = 10.15
    Em: float = 3.0
    Fm: float = float(mp.pi / mp.sqrt(3))
    Bp_s: float = 2 / 3
    Cp_s: float = 3 / 2
    Dp_s: float = 14.28
    Ep_s: float = 8.525
    Cm_s: float = 0.0
    Dm_s: float = 34.15


@dataclass
class SolverParams:   # SimilarityTheoryFluxes defaults, similarity_theory_turbulent_fluxes.jl:174-214 ; roughness_lengths.jl:93-139
    kappa: float = 0.4
    g: float = 9.80665
    charnock: float = 0.02
    smooth_wall: float = 0.11
    nu: float = 1.5e-5
    ell_max: float = 1.0
    reynolds_A: float = 5.85e-5
    reynolds_b: float = 0.72
    scalar_ell_max: float = 1.6e-4
    gust_beta: float = 1.2
    gust_floor: float = 0.01
    tol: float = 1e-8
    maxiter: int = 100
    x_h2o: float = 0.98          # component_interfaces.jl:353-358
    thermo: Thermo = field(default_factory=Thermo)
    edson: EdsonConstants = field(default_factory=EdsonConstants)


# ----------------------------------------------------------------------------------------------------------------
# thermodynamics
# ----------------------------------------------------------------------------------------------------------------
def p_sat(T, th: Thermo = Thermo(), ice=False):
    """p_tr (T/T_tr)^(Δc_p/R_v) exp[(ℒ₀ − Δc_p T₀)/R_v (1/T_tr − 1/T)],  Δc_p = c_pv − c_pl (liquid) or c_pv − c_pi with ℒ_s0 (ice)."""
    T = _m(T)
    dcp = _m(th.cp_v) - _m(th.cp_i if ice else th.cp_l)
    L0 = _m(th.LH_s0 if ice else th.LH_v0)
    return _m(th.p_tr) * (T / _m(th.T_tr)) ** (dcp / th.R_v) * mp.e ** ((L0 - dcp * _m(th.T_0)) / th.R_v * (1 / _m(th.T_tr) - 1 / T))


def q_surface(p, T, x_h2o=0.98, th: Thermo = Thermo(), ice=False):
    p = _m(p)
    pv = _m(x_h2o) * p_sat(T, th, ice)
    pv = min(pv, _m(0.999) * p)
    e = 1 / th.eps_vd                      # R_d / R_v
    return e * pv / (p - (1 - e) * pv)


def R_mix(q, th):
    return th.R_d * (1 - q) + th.R_v * q


def cp_mix(q, th):
    return th.cp_d * (1 - q) + _m(th.cp_v) * q


def buoyancy_scale(theta_star, q_star, Ts, qs, g, th):
    """b★ = g/𝒯ₛ [θ★ (1 + δ qₛ) + δ 𝒯ₛ q★], 𝒯ₛ = Tₛ R_m(qₛ)/R_d, δ = R_v/R_d − 1."""
    Tv = Ts * R_mix(qs, th) / th.R_d
    delta = th.eps_vd - 1
    return g / Tv * (theta_star * (1 + delta * qs) + delta * Tv * q_star)


# ----------------------------------------------------------------------------------------------------------------
# Edson et al. (2013) stability functions
# ----------------------------------------------------------------------------------------------------------------
def _convective_branch(z, D, E, F):
    y = mp.cbrt(1 - D * z)
    rE = mp.sqrt(E)
    return E / 2 * mp.log((1 + y + y * y) / E) - rE * mp.atan((1 + 2 * y) / rE) + F


def psi_momentum(zeta, c: EdsonConstants = EdsonConstants()):
    z = _m(zeta)
    if z < 0:
        x = (1 - _m(c.Am) * z) ** M("0.25")
        B = _m(c.Bm)
        kansas = B * mp.log((1 + x) / B) + mp.log((1 + x * x) / B) - B * mp.atan(x) + _m(c.Cm_m)
        conv = _convective_branch(z, _m(c.Dm_m), _m(c.Em), _m(c.Fm))
        f = z * z / (1 + z * z)
        return (1 - f) * kansas + f * conv
    dz = min(_m(c.zmax), _m(c.Ap) * z)
    return -_m(c.Bp_m) * z - _m(c.Cp_m) * (z - _m(c.Dp_m)) * mp.exp(-dz) - _m(c.Cp_m) * _m(c.Dp_m)


def psi_scalar(zeta, c: EdsonConstants = EdsonConstants()):
    z = _m(zeta)
    if z < 0:
        x = mp.sqrt(1 - _m(c.Am) * z)
        B = _m(c.Bm)
        kansas = B * mp.log((1 + x) / B) + _m(c.Cm_s)
        conv = _convective_branch(z, _m(c.Dm_s), _m(c.Em), _m(c.Fm))
        f = z * z / (1 + z * z)
        return (1 - f) * kansas + f * conv
    dz = min(_m(c.zmax), _m(c.Ap) * z)
    return -(1 + _m(c.Bp_s) * z) ** _m(c.Cp_s) - _m(c.Bp_s) * (z - _m(c.Dp_s)) * mp.exp(-dz) - _m(c.Ep_s)


# ----------------------------------------------------------------------------------------------------------------
# The other shipped stability functions, from the FLUX–PROFILE RELATIONS of the papers the reference cites — not from
# the closed forms it codes (similarity_theory_turbulent_fluxes.jl:636-745).  A stability function is the integral
#     ψ(ζ) = ∫₀^ζ (1 − φ(x)) / x dx                                   (Paulson 1970, eq. 3; Grachev et al. 2007, eq. 10)
# of the non-dimensional gradient φ; evaluating that integral numerically at 50 digits shares no algebra with the
# antiderivatives the reference (and the oracle) use.
# ----------------------------------------------------------------------------------------------------------------
def phi_sheba_momentum(x, a=6.5, b=1.3):
    """Grachev et al. (2007) eq. 9a is φ_m = 1 + a_m ζ (1 + ζ)^{1/3} / (1 + b_m ζ) with a_m = 5, b_m = a_m / 6.5, and their eq. 12
    is its integral.  The reference evaluates eq. 12 with its fields a = 6.5, b = 1.3 in the places of a_m, b_m
    (similarity_theory_turbulent_fluxes.jl:636-657), i.e. it integrates φ_m = 1 + 6.5 ζ (1 + ζ)^{1/3} / (1 + 1.3 ζ) — not the
    φ of its own comment (:642: 1 + a ζ ∛(1 + ζ) / (b + ζ), whose integral is 23 % smaller at small ζ; this pin found the
    difference).  A drop-in reproduces the reference as coded, so THAT function is the one pinned here."""
    return 1 + _m(a) * x * mp.cbrt(1 + x) / (1 + _m(b) * x)


def phi_sheba_scalar(x, a=5.0, b=5.0, c=3.0):     # eq. 9b: φ_h = 1 + (a ζ + b ζ²) / (1 + c ζ + ζ²)
    return 1 + (_m(a) * x + _m(b) * x * x) / (1 + _m(c) * x + x * x)


def phi_businger_dyer_momentum(x, gamma=16.0):    # Paulson (1970) / Businger–Dyer: φ_m = (1 − γ ζ)^{-1/4}, ζ < 0
    return (1 - _m(gamma) * x) ** M("-0.25")


def phi_businger_dyer_scalar(x, gamma=16.0):      # φ_h = (1 − γ ζ)^{-1/2}, ζ < 0
    return (1 - _m(gamma) * x) ** M("-0.5")


def phi_linear_stable(x, c=5.0, zmax=10.0):       # Large & Yeager (2004): φ = 1 + c ζ, ψ held at its ζ_max value beyond
    return 1 + _m(c) * x if x <= _m(zmax) else M(1)


def psi_from_phi(phi, zeta):
    """∫₀^ζ (1 − φ(x))/x dx, decade by decade (the integrand is smooth but varies over many scales)."""
    z = _m(zeta)
    if z == 0:
        return M(0)
    sgn = 1 if z > 0 else -1
    a = abs(z)
    edges = [M(0)]
    e = M("1e-8")
    while e < a:
        edges.append(e)
        e *= 10
    edges.append(a)
    f = lambda t: (1 - phi(sgn * t)) / t        # noqa: E731   (dx/x is invariant under x -> -x)
    return mp.quad(f, edges)


def psi_split(stable_phi, unstable_phi, zeta):
    """SplitStabilityFunction(stable, unstable): the stable function sees max(0, ζ), the unstable one min(0, ζ)."""
    z = _m(zeta)
    if z > 0:
        return psi_from_phi(stable_phi, z) if stable_phi is not None else M(0)
    return psi_from_phi(unstable_phi, z) if unstable_phi is not None else M(0)


# ----------------------------------------------------------------------------------------------------------------
# roughness, gustiness
# ----------------------------------------------------------------------------------------------------------------
def ell_momentum(ustar, P: SolverParams):
    u = _m(ustar)
    wave = _m(P.charnock) * u * u / _m(P.g)
    visc = _m(P.smooth_wall) * _m(P.nu) / u if P.smooth_wall != 0 else M(0)
    return min(wave + visc, _m(P.ell_max))


def ell_scalar(ell_u, ustar, P: SolverParams):
    Re = _m(ell_u) * _m(ustar) / _m(P.nu)
    val = M(0) if Re == 0 else _m(P.reynolds_A) / Re ** _m(P.reynolds_b)
    return min(val, _m(P.scalar_ell_max))


def gustiness_squared(ustar, bstar, h_bl, P: SolverParams):
    Jb = max(M(0), -_m(ustar) * _m(bstar))
    ug = max(_m(P.gust_floor), _m(P.gust_beta) * mp.cbrt(Jb * _m(h_bl)))
    return ug * ug


# ----------------------------------------------------------------------------------------------------------------
# one trip and the fixed point
# ----------------------------------------------------------------------------------------------------------------
def most_map(state, inv, P: SolverParams):
    """(u★, θ★, q★) ↦ (u★, θ★, q★): one iterate_interface_state of the default a–o tree with BulkTemperature.
    inv = dict(Ts, qs, dtheta, dq, du, dv, h, h_bl)."""
    us, ts, qs_ = state
    th = P.thermo
    b = buoyancy_scale(ts, qs_, inv["Ts"], inv["qs"], _m(P.g), th)
    U = mp.sqrt(inv["du"] ** 2 + inv["dv"] ** 2 + gustiness_squared(us, b, inv["h_bl"], P))
    lu = ell_momentum(us, P)
    ls = ell_scalar(lu, us, P)
    h = max(inv["h"], 2 * lu)                                  # zero-plane displacement 0
    kap = _m(P.kappa)

    def profile(psi, ell):
        if b == 0:                                            # L★ = Inf: ψ(0) terms
            return mp.log(h / ell) - psi(0) + psi(0)
        L = us * us / (kap * b)
        return mp.log(h / ell) - psi(h / L) + psi(ell / L)

    cu = kap / profile(lambda z: psi_momentum(z, P.edson), lu)
    cs = kap / profile(lambda z: psi_scalar(z, P.edson), ls)
    return cu * U, cs * inv["dtheta"], cs * inv["dq"]


def invariants(atm, ocean, P: SolverParams):
    """atm = (u, v, T, p, q), ocean = (u, v, T_kelvin, S) as doubles; h = 10 m, h_bl = 512 m defaults passed in P?"""
    ua, va, Ta, pa, qa = (_m(x) for x in atm)
    uo, vo, To, So = (_m(x) for x in ocean)
    th = P.thermo
    qs = q_surface(pa, To, P.x_h2o, th)
    h = _m(P_h(P))
    theta_a = Ta + _m(P.g) * h / cp_mix(qa, th)
    return dict(Ts=To, qs=qs, dtheta=theta_a - To, dq=qa - qs, du=ua - uo, dv=va - vo, h=h, h_bl=_m(P_hbl(P)))


def P_h(P):
    return getattr(P, "surface_layer_height", 10.0)


def P_hbl(P):
    return getattr(P, "boundary_layer_height", 512.0)


def solve_point(atm, ocean, P: SolverParams = SolverParams(), round_iterate=True):
    """compute_interface_state: do-while until |Δu★| + |Δθ★| + |Δq★| < tol or maxiter trips.  With round_iterate the iterate
    is rounded to Float64 after every trip, as the reference stores it (compute_interface_state.jl:116-121) — otherwise the
    exact orbit.  Returns (u★, θ★, q★, trips, fluxes dict)."""
    inv = invariants(atm, ocean, P)
    s = (M("1e-4"),) * 3 if not round_iterate else (_m(1e-4),) * 3
    trips = 0
    while True:
        new = most_map(s, inv, P)
        if round_iterate:
            new = tuple(_m(float(x)) for x in new)
        trips += 1
        drift = abs(new[0] - s[0]) + abs(new[1] - s[1]) + abs(new[2] - s[2])
        s = new
        if drift < _m(P.tol) or trips >= P.maxiter:
            break
    us, ts, qs_ = s
    ua, va, Ta, pa, qa = (_m(x) for x in atm)
    th = P.thermo
    dU = mp.sqrt(inv["du"] ** 2 + inv["dv"] ** 2)
    taux = M(0) if dU == 0 else -us * us * inv["du"] / dU
    tauy = M(0) if dU == 0 else -us * us * inv["dv"] / dU
    rho = pa / (R_mix(qa, th) * Ta)
    Lv = _m(th.LH_v0) + (_m(th.cp_v) - _m(th.cp_l)) * (Ta - _m(th.T_0))
    fluxes = dict(latent_heat=-rho * Lv * us * qs_, sensible_heat=-rho * cp_mix(qa, th) * us * ts, water_vapor=-rho * us * qs_,
                  x_momentum=rho * taux, y_momentum=rho * tauy)
    return us, ts, qs_, trips, fluxes


# ----------------------------------------------------------------------------------------------------------------
# atmosphere – sea ice: the same similarity step with the sea-ice stability functions, ice-phase surface humidity and a
# skin temperature re-solved every trip (docs/src/interface_fluxes.md:679-700: "a four-variable system"; the balance and its
# linearisation are the comment block of interface_states.jl:459-467)
# ----------------------------------------------------------------------------------------------------------------
def psi_closed(kind, zeta):
    """Antiderivatives of the flux–profile relations above, derived by hand and CHECKED against psi_from_phi to 1e-25 by
    tests/test_oracle_independent_pin.py (so the whole-solve pin below does not pay a quadrature per trip)."""
    z = _m(zeta)
    if kind == "sheba_momentum":                     # ∫ of 1 − φ, φ = 1 + a ζ (1+ζ)^{1/3} / (1 + b ζ): substitute x = (1+ζ)^{1/3}
        if z <= 0:
            return M(0)
        a, b = _m(6.5), _m(1.3)                      # the doubles the reference holds
        x = mp.cbrt(1 + z)
        B = -mp.cbrt((b - 1) / b)                    # real cube root of (1 − b)/b < 0
        r3 = mp.sqrt(3)
        poly = -3 * a * (x - 1) / b
        rest = a * B / (2 * b) * (2 * mp.log((x + B) / (1 + B)) - mp.log((x * x - B * x + B * B) / (1 - B + B * B))
                                  + 2 * r3 * (mp.atan((2 * x - B) / (r3 * B)) - mp.atan((2 - B) / (r3 * B))))
        return poly + rest
    if kind == "sheba_scalar":                       # partial fractions of (a ζ + b ζ²)/(ζ (1 + c ζ + ζ²))
        if z <= 0:
            return M(0)
        a, b, c = M(5), M(5), M(3)
        B = mp.sqrt(c * c - 4)
        return -b / 2 * mp.log(1 + c * z + z * z) + (b * c / (2 * B) - a / B) * (
            mp.log((2 * z + c - B) / (2 * z + c + B)) - mp.log((c - B) / (c + B)))
    if kind == "paulson_momentum":                   # Paulson (1970) eq. 8
        if z >= 0:
            return M(0)
        x = (1 - 16 * z) ** M("0.25")
        return 2 * mp.log((1 + x) / 2) + mp.log((1 + x * x) / 2) - 2 * mp.atan(x) + mp.pi / 2
    if kind == "paulson_scalar":                     # Paulson (1970) eq. 9
        if z >= 0:
            return M(0)
        return 2 * mp.log((1 + mp.sqrt(1 - 16 * z)) / 2)
    raise ValueError(kind)


@dataclass
class SeaIceParams:   # component_interfaces.jl:262-290, 396-418 (defaults of an OceanSeaIceModel); ClimaSeaIce defaults (third party)
    conductivity: float = 2.0                 # SkinTemperature(ConductiveFlux(k))
    max_dT: float = 5.0                       # interface_states.jl:360
    liquidus_slope: float = 0.054             # LinearLiquidus
    freshwater_melting_celsius: float = 0.0
    albedo: float = 0.7
    emissivity: float = 1.0
    sigma: float = 5.670374419e-8


def sea_ice_trip(state, atm, ice, P: SolverParams, I: SeaIceParams, rnd):
    """state = (u★, θ★, q★, Tₛ, qₛ); atm = (u, v, T, p, q, sw, lw); ice = (S_ocean, hi, hc).  `rnd` rounds to the model's
    element type where the reference converts (the new Tₛ, qₛ and scales)."""
    us, ts, qs_star, Ts_prev, q_prev = state
    ua, va, Ta, pa, qa, sw, lw = atm
    S_oc, hi, hc = ice
    th = P.thermo
    h = _m(P_h(P))
    rho = pa / (R_mix(qa, th) * Ta)
    cpm = cp_mix(qa, th)
    L_sub = _m(th.LH_s0) + (_m(th.cp_v) - _m(th.cp_i)) * (Ta - _m(th.T_0))
    # flux balance with the upwelling long wave linearised about the previous skin temperature
    up = _m(I.sigma) * _m(I.emissivity) * Ts_prev ** 4
    Q_down = -(1 - _m(I.albedo)) * sw - _m(I.emissivity) * lw
    Q_sensible = -rho * cpm * us * ts
    Q_latent = -rho * L_sub * us * qs_star
    Q_a = Q_latent + up + Q_down
    theta_a = Ta + _m(P.g) * h / cpm
    dT = theta_a - Ts_prev
    omega = M(0) if dT == 0 else Q_sensible / dT
    beta = 4 * up / Ts_prev
    Rth = hi / _m(I.conductivity)
    T_bottom = _m(I.freshwater_melting_celsius) - _m(I.liquidus_slope) * S_oc + M("273.15")
    D = 1 + beta * Rth - omega * Rth
    T_new = Ts_prev if D == 0 else (T_bottom + beta * Rth * Ts_prev - omega * Rth * theta_a - Q_a * Rth) / D
    step = max(-_m(I.max_dT), min(_m(I.max_dT), T_new - Ts_prev))
    T_new = min(Ts_prev + step, _m(I.freshwater_melting_celsius) + M("273.15"))
    if not hi >= hc:
        T_new = T_bottom
    T_new = rnd(T_new)
    q_new = rnd(q_surface(pa, T_new, 1.0, th, ice=True))
    # similarity step at the new surface state
    b = buoyancy_scale(ts, qs_star, T_new, q_new, _m(P.g), th)
    U = mp.sqrt(ua * ua + va * va + gustiness_squared(us, b, _m(P_hbl(P)), P))     # sea ice at rest (atmosphere_sea_ice_fluxes.jl:96-97)
    lu = ell_momentum(us, P)
    ls = ell_scalar(lu, us, P)
    hh = max(h, 2 * lu)
    kap = _m(P.kappa)

    def profile(stable, unstable, ell):
        if b == 0:
            return mp.log(hh / ell)
        L = us * us / (kap * b)
        f = lambda zz: psi_closed(stable, zz) if zz > 0 else psi_closed(unstable, zz)   # noqa: E731
        return mp.log(hh / ell) - f(hh / L) + f(ell / L)

    cu = kap / profile("sheba_momentum", "paulson_momentum", lu)
    cs = kap / profile("sheba_scalar", "paulson_scalar", ls)
    return rnd(cu * U), rnd(cs * (theta_a - T_new)), rnd(cs * (qa - q_new)), T_new, q_new


def solve_sea_ice_point(atm, ice, Ts0_kelvin, P: SolverParams = SolverParams(), I: SeaIceParams = SeaIceParams()):
    """compute_interface_state for one consolidated-or-not ice point; returns (u★, θ★, q★, Tₛ, trips, fluxes)."""
    rnd = lambda x: _m(float(x))   # noqa: E731
    atm = tuple(_m(x) for x in atm)
    ice = tuple(_m(x) for x in ice)
    th = P.thermo
    Ts = _m(Ts0_kelvin)
    s = (_m(9.999999747378752e-05),) * 3 + (Ts, rnd(q_surface(atm[3], Ts, 1.0, th, ice=True)))   # convert(FT, 1f-4): the Float32 literal, widened (atmosphere_sea_ice_fluxes.jl:127)
    trips = 0
    while True:
        new = sea_ice_trip(s, atm, ice, P, I, rnd)
        trips += 1
        drift = abs(new[0] - s[0]) + abs(new[1] - s[1]) + abs(new[2] - s[2])
        s = new
        if drift < _m(P.tol) or trips >= P.maxiter:
            break
    us, ts, qs_, Ts, _ = s
    ua, va, Ta, pa, qa = atm[:5]
    rho = pa / (R_mix(qa, th) * Ta)
    L_sub = _m(th.LH_s0) + (_m(th.cp_v) - _m(th.cp_i)) * (Ta - _m(th.T_0))
    dU = mp.sqrt(ua * ua + va * va)
    fluxes = dict(latent_heat=-rho * us * qs_ * L_sub, sensible_heat=-rho * cp_mix(qa, th) * us * ts, water_vapor=-rho * us * qs_,
                  x_momentum=M(0) if dU == 0 else rho * (-us * us * ua / dU), y_momentum=M(0) if dU == 0 else rho * (-us * us * va / dU))
    return us, ts, qs_, Ts, trips, fluxes


# ----------------------------------------------------------------------------------------------------------------
# interpolation (Oceananigans, third party: restated from its documented behaviour)
# ----------------------------------------------------------------------------------------------------------------
def interpolator(f):
    """fractional index (0-based) → (i⁻, i⁺, ξ), 1-based indices: i⁻ = trunc(f) + 1, i⁺ = i⁻ + sign(f), ξ = f mod 1."""
    f = _m(f)
    t = int(mp.floor(f)) if f >= 0 else -int(mp.floor(-f))
    sgn = 0 if f == 0 else (1 if f > 0 else -1)
    xi = f - mp.floor(f)
    return t + 1, t + 1 + sgn, xi


def bilinear_time(corners1, corners2, xi, eta, nt, same):
    """corners = (d(i⁻,j⁻), d(i⁻,j⁺), d(i⁺,j⁻), d(i⁺,j⁺)) at the two time levels; weights (1−ξ)(1−η) …; ψ₂ ñ + ψ₁ (1 − ñ)."""
    xi, eta, nt = _m(xi), _m(eta), _m(nt)

    def space(c):
        a, b, c_, d = (_m(x) for x in c)
        return (1 - xi) * (1 - eta) * a + (1 - xi) * eta * b + xi * (1 - eta) * c_ + xi * eta * d

    p1 = space(corners1)
    if same:
        return p1
    return space(corners2) * nt + p1 * (1 - nt)
