// ne_physics.cuh — device-side physics of the turbulent flux solve (sm_100a).
//
// Kernel variants of the reference's plugin types (all citations relative to
// /root/reference/src/EarthSystemModels/InterfaceComputations/):
//   stability functions      similarity_theory_turbulent_fluxes.jl:445-752
//   roughness lengths        roughness_lengths.jl:182-246
//   subgrid velocities       similarity_theory_turbulent_fluxes.jl:88-98
//   saturation humidity      interface_states.jl:44-74, 236-277
//   temperature formulations interface_states.jl:330-577
//   coefficient-based fluxes coefficient_based_turbulent_fluxes.jl:265-371
//
// Element types follow the reference's promotion rules through C++'s usual arithmetic
// conversions: FT = exchange grid, CT = thermodynamics parameters, VT = kinematic viscosity
// (double for the Float64 literal 1.5e-5, roughness_lengths.jl:94,126).  Only the ψ branch the
// reference's ifelse selects is evaluated (the discarded branch has no side effects).
#pragma once

#include "ne_common.cuh"

namespace ne {

// ---- thermodynamics (parameters: src/Atmospheres/thermodynamic_parameters.jl:30-258;
// formulas: Thermodynamics.jl, docs/src/interface_fluxes.md:94-96,507) -----------------------------
template <class CT>
struct Thermo {
  CT R_d, R_v, eps, eps_inv, delta, cp_d, cp_v, cp_l, cp_i, LH_v0, LH_s0, T_0, T_triple, press_triple;
  // constants of saturation_vapor_pressure per phase (0 liquid, 1 ice), formed once on the host in CT arithmetic in the
  // order the formula applies them: Δcp/R_v, (ℒ₀ − Δcp T₀)/R_v, 1/T_triple
  CT psat_pow[2], psat_exp[2], inv_T_triple;

  static Thermo make(const NeThermoParams& p) {  // host: derived constants in CT arithmetic (:73-78, 256)
    Thermo t;
    CT R = (CT)p.gas_constant, Md = (CT)p.dry_air_molar_mass, Mv = (CT)p.water_molar_mass;
    t.R_d = R / Md;
    t.R_v = R / Mv;
    t.eps = Md / Mv;
    t.eps_inv = 1 / t.eps;
    t.delta = t.eps - 1;
    t.cp_d = t.R_d / (CT)p.kappa_d;
    t.cp_v = (CT)p.cp_v; t.cp_l = (CT)p.cp_l; t.cp_i = (CT)p.cp_i;
    t.LH_v0 = (CT)p.LH_v0; t.LH_s0 = (CT)p.LH_s0; t.T_0 = (CT)p.T_0;
    t.T_triple = (CT)p.T_triple; t.press_triple = (CT)p.press_triple;
    const CT dcp[2] = {(CT)(t.cp_v - t.cp_l), (CT)(t.cp_v - t.cp_i)}, LH[2] = {t.LH_v0, t.LH_s0};
    for (int ph = 0; ph < 2; ++ph) {
      volatile CT prod = dcp[ph] * t.T_0;       // volatile: separately rounded whatever the host compiler's contraction mode
      volatile CT diff = LH[ph] - prod;
      t.psat_pow[ph] = dcp[ph] / t.R_v;
      t.psat_exp[ph] = diff / t.R_v;
    }
    t.inv_T_triple = 1 / t.T_triple;
    return t;
  }

  __device__ __forceinline__ CT saturation_vapor_pressure(CT T, int phase) const {
    CT LH_0 = phase == NE_PHASE_LIQUID ? LH_v0 : LH_s0;
    CT dcp = phase == NE_PHASE_LIQUID ? (cp_v - cp_l) : (cp_v - cp_i);
    return press_triple * m_pow(T / T_triple, dcp / R_v) * m_exp((LH_0 - dcp * T_0) / R_v * (1 / T_triple - 1 / T));
  }
  template <class Q> __device__ __forceinline__ auto gas_constant_air(Q q) const { return R_d * (1 - q) + R_v * q; }
  template <class T, class P, class Q> __device__ __forceinline__ auto air_density(T temp, P p, Q q) const {
    return p / (gas_constant_air(q) * temp);
  }
  template <class Q> __device__ __forceinline__ auto cp_m(Q q) const { return cp_d * (1 - q) + cp_v * q; }
  template <class T> __device__ __forceinline__ auto latent_heat_vapor(T temp) const { return LH_v0 + (cp_v - cp_l) * (temp - T_0); }
  template <class T> __device__ __forceinline__ auto latent_heat_sublim(T temp) const { return LH_s0 + (cp_v - cp_i) * (temp - T_0); }
  template <class T, class Q> __device__ __forceinline__ auto virtual_temperature(T temp, Q q) const {
    return temp * gas_constant_air(q) / R_d;
  }
};

// ---- surface specific humidity (interface_states.jl:55-74) ----------------------------------------
template <class FT>
__device__ __forceinline__ FT water_mole_fraction(const NeInterfaceProperties& ip, FT S) {  // :255-277
  FT s = S / 1000;
  FT alpha = (FT)ip.water_molar_mass * ((FT)ip.constituent_mass_fraction[0] / (FT)ip.constituent_molar_mass[0] +
                                        (FT)ip.constituent_mass_fraction[1] / (FT)ip.constituent_molar_mass[1] +
                                        (FT)ip.constituent_mass_fraction[2] / (FT)ip.constituent_molar_mass[2] +
                                        (FT)ip.constituent_mass_fraction[3] / (FT)ip.constituent_molar_mass[3]);
  return (1 - s) / (1 - s + alpha * s);
}

template <class FT, class CT>
__device__ __forceinline__ FT surface_specific_humidity(const NeInterfaceProperties& ip, const Thermo<CT>& th,
                                                        FT p_at, FT Ts, FT Ss) {
  CT T = (CT)Ts, p = (CT)p_at;
  CT psat = th.saturation_vapor_pressure(T, ip.phase);
  using W = decltype(FT() * CT());
  W pv;
  if (ip.x_h2o_kind == NE_XH2O_ONE) pv = psat;
  else if (ip.x_h2o_kind == NE_XH2O_CONSTANT) pv = (FT)ip.x_h2o * psat;
  else pv = water_mole_fraction<FT>(ip, Ss) * psat;
  pv = mn(pv, (CT)0.999 * p);
  auto q = th.eps_inv * pv / (p - (1 - th.eps_inv) * pv);
  return (FT)q;
}

// ---- stability functions -----------------------------------------------------------------------
// Edson et al. (2013) momentum :501-532
template <class FT, class Z>
__device__ __forceinline__ auto psi_edson_momentum(const double* p, Z zeta) -> decltype(FT() * Z()) {
  using W = decltype(FT() * Z());
  if (zeta < 0) {
    FT Am = (FT)p[5], Bm = (FT)p[6], Cm = (FT)p[7], Dm = (FT)p[8], Em = (FT)p[9], Fm = (FT)p[10];
    auto f1 = m_sqrt(m_sqrt(1 - Am * zeta));
    auto psi1 = Bm * m_log((1 + f1) / Bm) + m_log((1 + sq(f1)) / Bm) - Bm * m_atan(f1) + Cm;
    auto f2 = m_cbrt(1 - Dm * zeta);
    FT rE = m_sqrt(Em);
    auto psi2 = Em / 2 * m_log((1 + f2 + sq(f2)) / Em) - rE * m_atan((1 + 2 * f2) / rE) + Fm;
    auto z2 = sq(zeta);
    auto fw = z2 / (1 + z2);
    return (W)((1 - fw) * psi1 + fw * psi2);
  } else {
    FT zmax = (FT)p[0], Ap = (FT)p[1], Bp = (FT)p[2], Cp = (FT)p[3], Dp = (FT)p[4];
    Z zp = mx((Z)0, zeta);
    auto dz = mn(zmax, Ap * zp);
    return (W)(-Bp * zp - Cp * (zp - Dp) * m_exp(-dz) - Cp * Dp);
  }
}

// Edson et al. (2013) scalar :586-618
template <class FT, class Z>
__device__ __forceinline__ auto psi_edson_scalar(const double* p, Z zeta) -> decltype(FT() * Z()) {
  using W = decltype(FT() * Z());
  if (zeta < 0) {
    FT Am = (FT)p[6], Bm = (FT)p[7], Cm = (FT)p[8], Dm = (FT)p[9], Em = (FT)p[10], Fm = (FT)p[11];
    auto f1 = m_sqrt(1 - Am * zeta);
    auto psi1 = Bm * m_log((1 + f1) / Bm) + Cm;
    auto f2 = m_cbrt(1 - Dm * zeta);
    FT rE = m_sqrt(Em);
    auto psi2 = Em / 2 * m_log((1 + f2 + sq(f2)) / Em) - rE * m_atan((1 + 2 * f2) / rE) + Fm;
    auto z2 = sq(zeta);
    auto fw = z2 / (1 + z2);
    return (W)((1 - fw) * psi1 + fw * psi2);
  } else {
    FT zmax = (FT)p[0], Ap = (FT)p[1], Bp = (FT)p[2], Cp = (FT)p[3], Dp = (FT)p[4], Ep = (FT)p[5];
    Z zp = mx((Z)0, zeta);
    auto dz = mn(zmax, Ap * zp);
    return (W)(-m_pow(1 + Bp * zp, Cp) - Bp * (zp - Dp) * m_exp(-dz) - Ep);
  }
}

template <class FT, class Z>
__device__ auto stability_fn(const NeStabilityFn& f, Z zeta) -> decltype(FT() * Z()) {
  using W = decltype(FT() * Z());
  const double* p = f.p;
  switch (f.kind) {
    case NE_PSI_ZERO: return (W)0;
    case NE_PSI_EDSON_MOMENTUM: return psi_edson_momentum<FT, Z>(p, zeta);
    case NE_PSI_EDSON_SCALAR: return psi_edson_scalar<FT, Z>(p, zeta);
    case NE_PSI_SHEBA_MOMENTUM: {  // :643-657 (rt3 = sqrt(3) is a Float64 in the reference)
      FT a = (FT)p[0], b = (FT)p[1];
      Z zp = mx((Z)0, zeta);
      auto z = m_cbrt(1 + zp);
      FT B = m_cbrt((1 - b) / b);
      const double rt3 = 1.7320508075688772;
      auto P1 = -3 * a * (z - 1) / b;
      auto P2 = a * B / (2 * b) *
                (2 * m_log((z + B) / (1 + B)) - m_log((sq(z) - B * z + sq(B)) / (1 - B + sq(B))) +
                 2 * rt3 * (m_atan((2 * z - B) / (rt3 * B)) - m_atan((2 - B) / (rt3 * B))));
      return (W)(P1 + P2);
    }
    case NE_PSI_SHEBA_SCALAR: {  // :665-677
      FT a = (FT)p[0], b = (FT)p[1], c = (FT)p[2];
      FT B = m_sqrt(sq(c) - 4);
      Z zp = mx((Z)0, zeta);
      auto P1 = -b / 2 * m_log(1 + c * zp + sq(zp));
      auto P2 = (b * c / (2 * B) - a / B) * (m_log((2 * zp + c - B) / (2 * zp + c + B)) - m_log((c - B) / (c + B)));
      return (W)(P1 + P2);
    }
    case NE_PSI_PAULSON_MOMENTUM: {  // :688-699
      FT a = (FT)p[0], b = (FT)p[1];
      Z zm = mn((Z)0, zeta);
      auto z = m_sqrt(m_sqrt(1 - a * zm));
      auto P1 = 2 * m_log((1 + z) / 2);
      auto P2 = m_log((1 + sq(z)) / 2);
      auto P3 = -2 * m_atan(z);
      return (W)(P1 + P2 + P3 + b);
    }
    case NE_PSI_PAULSON_SCALAR: {  // :705-710
      FT a = (FT)p[0];
      Z zm = mn((Z)0, zeta);
      auto z = m_sqrt(m_sqrt(1 - a * zm));
      return (W)(2 * m_log((1 + sq(z)) / 2));
    }
    case NE_PSI_LINEAR_STABLE: {  // :747-752
      FT c = (FT)p[0], zmax = (FT)p[1];
      Z zp = mx((Z)0, zeta);
      return (W)(-c * mn(zp, zmax));
    }
  }
  return (W)0;
}

// SplitStabilityFunction :720-725 — only the selected side is evaluated
template <class FT, class Z>
__device__ __forceinline__ auto stability_profile(const NeStabilityProfile& s, Z zeta) -> decltype(FT() * Z()) {
  if (!s.split) return stability_fn<FT, Z>(s.a, zeta);
  return (zeta > 0) ? stability_fn<FT, Z>(s.a, zeta) : stability_fn<FT, Z>(s.b, zeta);
}

// similarity_profile :242-253
template <class FT, class H, class L, class LS>
__device__ __forceinline__ auto similarity_profile(int form, const NeStabilityProfile& psi, H h, L ell, LS Lstar) {
  auto psi_h = stability_profile<FT>(psi, h / Lstar);
  if (form == NE_PROFILE_COARE) return m_log(h / ell) - psi_h;
  auto psi_l = stability_profile<FT>(psi, ell / Lstar);
  return m_log(h / ell) - psi_h + psi_l;
}

// ---- roughness lengths ---------------------------------------------------------------------------
template <class FT, class T>
__device__ __forceinline__ FT temperature_dependent_viscosity(const NeRoughnessLength& r, T Tk) {  // :185-189
  FT Tp = (FT)(Tk - 273.15);
  return (FT)r.nu_C[0] + (FT)r.nu_C[1] * Tp + (FT)r.nu_C[2] * sq(Tp) + (FT)r.nu_C[3] * cube(Tp);
}
template <class FT, class VT, class T>
__device__ __forceinline__ VT air_viscosity(const NeRoughnessLength& r, T Ts) {
  if (r.visc_kind == NE_VISC_CONSTANT) return (VT)r.nu;
  return (VT)temperature_dependent_viscosity<FT>(r, Ts);
}

template <class FT, class VT, class US, class UU>
__device__ __forceinline__ auto momentum_roughness(const NeRoughnessLength& r, VT nu, US ustar, UU U) {  // :197-210
  FT g = (FT)r.gravitational_acceleration, Cnu = (FT)r.smooth_wall_parameter, lmax = (FT)r.maximum_roughness_length;
  using WG = decltype(FT() * UU());
  WG Cg;
  if (r.wave_kind == NE_WAVE_CONSTANT) Cg = (FT)r.wave_constant;
  else Cg = mx((UU)0, (FT)r.wave_C1 * mn(U, (FT)r.wave_Umax) + (FT)r.wave_C2);  // :75
  auto lW = Cg * sq(ustar) / g;
  using WR = decltype(Cnu * nu / ustar);
  WR lR = (Cnu == 0) ? (WR)0 : Cnu * nu / ustar;
  return mn(lW + lR, lmax);
}

template <class FT, class VT, class LU, class US>
__device__ __forceinline__ auto scalar_roughness(const NeRoughnessLength& r, VT nu, LU ell_u, US ustar) {  // :234-246
  auto Rstar = ell_u * ustar / nu;
  FT A = (FT)r.reynolds_A, b = (FT)r.reynolds_b;
  using WR = decltype(A / m_pow(Rstar, b));
  WR ls = (Rstar == 0) ? (WR)0 : A / m_pow(Rstar, b);
  return mn(ls, (FT)r.maximum_roughness_length);
}

// ---- subgrid velocities :88-98 -------------------------------------------------------------------
template <class FT, class US, class BS, class HB>
__device__ __forceinline__ auto vsgs2_one(int kind, const NeSubgridVelocity& s, double constant, US ustar, BS bstar,
                                          HB h_bl) -> decltype(FT() * US() * BS() * HB()) {
  using W = decltype(FT() * US() * BS() * HB());
  if (kind == NE_SGS_NONE) return (W)0;
  if (kind == NE_SGS_CONSTANT) { FT v = (FT)constant; return (W)sq(v); }
  auto Jb = -ustar * bstar;
  using J = decltype(Jb);
  auto UG = mx((FT)s.minimum_gustiness, (FT)s.gustiness_parameter * m_cbrt(mx((J)0, Jb) * h_bl));
  return (W)sq(UG);
}
template <class FT, class US, class BS, class HB>
__device__ __forceinline__ auto vsgs2(const NeSubgridVelocity& s, US ustar, BS bstar, HB h_bl) {
  auto c = vsgs2_one<FT>(s.convective_kind, s, s.convective_constant, ustar, bstar, h_bl);
  if (!s.composite) return c;
  return c + vsgs2_one<FT>(s.mesoscale_kind, s, s.mesoscale_constant, ustar, bstar, h_bl);
}

// PolynomialNeutralDragCoefficient :46-52 (coefficient_based_turbulent_fluxes.jl)
template <class FT, class UU>
__device__ __forceinline__ auto polynomial_drag(const NePolynomialDrag& p, UU U) {
  auto Um = mx(U, (FT)p.minimum_wind_speed);
  using W = decltype(Um);
  W poly = ((FT)p.a / Um + (FT)p.b + (FT)p.c * Um - (FT)p.d * pow6(Um)) / 1000;
  return (Um < (FT)p.high_wind_speed_threshold) ? poly : (W)(FT)p.high_wind_drag_coefficient;
}

// ---- per-point states ---------------------------------------------------------------------------
template <class FT> struct AtmosState { FT z, u, v, T, p, q, h_bl; };
template <class FT> struct Interior { FT u, v, T, S, kappa, hi, hs, hc; };
template <class FT> struct RadState { FT sigma, alpha, eps, sw, lw; };
template <class FT> struct Scales { FT ustar, theta_star, q_star; };

// SeaIceAlbedo stateindex (src/Radiations/sea_ice_albedo.jl:106-133)
template <class FT>
__device__ __forceinline__ FT sea_ice_albedo(const NeSeaIceAlbedo& a, int64_t idx) {
  const FT hi = __ldg((const FT*)a.ice_thickness + idx);
  const FT Ts = __ldg((const FT*)a.surface_temperature + idx);
  const FT hs = a.snow_thickness ? __ldg((const FT*)a.snow_thickness + idx) : (FT)0;
  const FT Tm = (FT)a.melting_temperature, dT = (FT)a.temperature_range;
  const FT fT = clampv((Ts - Tm + dT) / dT, (FT)0, (FT)1);
  FT alpha_i = (FT)a.ice_albedo - (FT)a.ice_melt_reduction * fT;
  const FT alpha_s = (FT)a.snow_albedo - (FT)a.snow_melt_reduction * fT;
  const FT alpha_o = (FT)a.ocean_albedo;
  const FT fh = clampv(hi / (FT)a.minimum_ice_thickness, (FT)0, (FT)1);
  alpha_i = alpha_o + (alpha_i - alpha_o) * fh;
  const FT fs = clampv(hs / (FT)a.minimum_snow_depth, (FT)0, (FT)1);
  return fs * alpha_s + (1 - fs) * alpha_i;
}

__device__ __forceinline__ float m_sin(float x) { return sinf(x); }
__device__ __forceinline__ double m_sin(double x) { return sin(x); }

// TabulatedAlbedo stateindex (src/Radiations/tabulated_albedo.jl:109-160); interpolator semantics as in
// ne_interp_device.cuh (Oceananigans interpolator, third party)
template <class FT>
__device__ __forceinline__ FT tabulated_albedo(const NeTabulatedAlbedo& a, FT lam_deg, FT phi_deg, FT sw) {
  const FT deg = (FT)3.141592653589793 / 180;       // deg2rad(x) = x * (π / 180) rounded to FT
  const FT phi = phi_deg * deg, lam = lam_deg * deg;
  const double h = (a.seconds_in_day - a.noon_in_seconds) * (FT)a.day_to_radians + lam;   // Float64 clock (:122)
  const FT delta = (FT)a.declination;
  auto cosz_raw = m_sin(phi) * m_sin(delta) + m_cos(h) * m_cos(delta) * m_cos(phi);
  using W = decltype(cosz_raw);
  const W cosz = cosz_raw > 0 ? cosz_raw : (W)0;
  const W Qmax = (FT)a.solar_constant * cosz;
  W tr = 0;
  if (Qmax > 0) { tr = sw / Qmax; if (tr > 1) tr = 1; }
  const FT t1 = (FT)a.t_values[0], dt = (FT)a.t_values[1] - t1;
  const FT p1 = (FT)a.phi_values[0], dp = (FT)a.phi_values[1] - p1;
  const W fi = (tr - t1) / dt;
  const FT fj = (m_abs(phi) - p1) / dp;
  auto interp1 = [](auto f, int32_t& im, int32_t& ip) {
    using T = decltype(f);
    im = (int32_t)f + 1;
    ip = im + ((f > 0) ? 1 : ((f < 0) ? -1 : 0));
    T r = f - trunc(f);
    if (r == 0) r = (T)0;
    else if (!(r > 0)) r = r + (T)1;
    return r;
  };
  int32_t im, ip, jm, jp;
  const W xi = interp1(fi, im, ip);
  const FT eta = interp1(fj, jm, jp);
  const FT* T = (const FT*)a.table;
  const int64_t nt = a.n_t;
  // i⁺/j⁺ one past the table edge (𝓉 = 1, |φ| = 90°) carry weight 0 in the reference, which reads them unchecked
  // (@inbounds): clamp the index so that the zero-weighted read stays inside the table
  auto at = [&](int32_t i, int32_t j) {
    i = i > a.n_t ? a.n_t : (i < 1 ? 1 : i);
    j = j > a.n_phi ? a.n_phi : (j < 1 ? 1 : j);
    return __ldg(T + (i - 1) + (int64_t)(j - 1) * nt);
  };
  return (FT)((1 - xi) * (1 - eta) * at(im, jm) + (1 - xi) * eta * at(im, jp) + xi * (1 - eta) * at(ip, jm) + xi * eta * at(ip, jp));
}

// radiation state of one surface (src/Radiations/air_sea_interface_radiation_state.jl:4-39)
template <class FT>
__device__ __forceinline__ RadState<FT> radiation_state(const NeSurfaceRadiation& r, const Layout& L, int64_t idx, int32_t j) {
  RadState<FT> s = {0, 0, 0, 0, 0};
  if (!r.enabled) return s;
  s.sigma = (FT)r.stefan_boltzmann_constant;
  s.sw = __ldg((const FT*)r.downwelling_shortwave + idx);
  s.lw = __ldg((const FT*)r.downwelling_longwave + idx);
  if (r.albedo_kind == NE_ALBEDO_CONSTANT) s.alpha = (FT)r.albedo;
  else if (r.albedo_kind == NE_ALBEDO_FIELD) s.alpha = __ldg((const FT*)r.albedo_field + idx);
  else if (r.albedo_kind == NE_ALBEDO_SEA_ICE) s.alpha = sea_ice_albedo<FT>(r.sea_ice_albedo, idx);
  else {
    const FT phi = r.nodes_2d ? __ldg((const FT*)r.latitude + idx) : __ldg((const FT*)r.latitude + (j + L.hy - 1));
    if (r.albedo_kind == NE_ALBEDO_TABULATED) {
      const int64_t col = idx - (int64_t)(j + L.hy - 1) * L.sx;   // i + hx - 1
      const FT lam = r.nodes_2d ? __ldg((const FT*)r.tabulated_albedo.longitude + idx)
                                : __ldg((const FT*)r.tabulated_albedo.longitude + col);
      s.alpha = tabulated_albedo<FT>(r.tabulated_albedo, lam, phi, s.sw);
    } else {  // latitude_dependent_albedo.jl:48-53, hack_cosd radiation_kernels.jl:1
      s.alpha = (FT)r.albedo - (FT)r.albedo_direct * m_cos((FT)3.141592653589793 * (2 * phi) / 180);
    }
  }
  s.eps = (FT)r.emissivity;
  return s;
}

template <class FT, class CT>
struct FluxEpilogue {
  FT Qv, Qc, Jv, tx, ty;
  __device__ __forceinline__ FluxEpilogue(const Thermo<CT>& th, const AtmosState<FT>& a, FT ustar, FT theta_star,
                                          FT q_star, FT du, FT dv, bool ice) {
    FT dU = m_sqrt(sq(du) + sq(dv));
    FT taux = (dU == 0) ? (FT)0 : -sq(ustar) * du / dU;
    FT tauy = (dU == 0) ? (FT)0 : -sq(ustar) * dv / dU;
    auto rho_a = th.air_density(a.T, a.p, a.q);
    auto cpm = th.cp_m(a.q);
    if (ice) Qv = (FT)(-rho_a * ustar * q_star * th.latent_heat_sublim(a.T));  // atmosphere_sea_ice_fluxes.jl:178
    else Qv = (FT)(-rho_a * th.latent_heat_vapor(a.T) * ustar * q_star);       // atmosphere_ocean_fluxes.jl:186
    Qc = (FT)(-rho_a * cpm * ustar * theta_star);
    Jv = (FT)(-rho_a * ustar * q_star);
    tx = (FT)(rho_a * taux);
    ty = (FT)(rho_a * tauy);
  }
};

}  // namespace ne
