// ne_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A scalar, line-by-line C++ restatement of the reference's atmosphere–surface interface
// path, used only as the checker by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs.  Nothing in the product package may call it.
//
// PARITY STATUS: pinned by (a) every known-answer / analytic test the reference's own
// test-suite holds for this path (tests/test_oracle_reference_kats.py cites them one by
// one; the reference has no golden vectors for converged fluxes: test/test_surface_fluxes.jl:
// 425-499 is commented out) and (b) an INDEPENDENT restatement in 50-digit arithmetic written
// from the reference's documentation (oracle/pin/reference_mp.py, no text shared with this
// file; tests/test_oracle_independent_pin.py: every function within a few ulp, whole fixed
// points to 1e-12 with equal trip counts; profiles/r02_oracle_pin.jsonl).  NOT pinned against
// outputs of the reference itself: Julia is absent from this container, so it cannot be run
// here ("parity unpinned" in that sense); julia/dump_reference.jl produces those outputs from
// the same raw inputs on any machine that has Julia + NumericalEarth, and
// tests/golden/make_golden.py diffs them.  The arithmetic that lives in third-party packages that are
// NOT under /root/reference is restated from their published algorithms and flagged
// [3rd-party] below:
//   Thermodynamics.jl  (compat "0.15.3, 1", Project.toml:113)  — saturation_vapor_pressure,
//       air_density, cp_m, latent_heat_vapor/sublim, virtual_temperature;
//   Oceananigans.jl    (compat "0.110.15, 0.111", Project.toml:102) — FractionalIndices,
//       interpolator, _interpolate, ℑ operators;
//   ClimaSeaIce.jl     (compat "0.5, 0.6", Project.toml:78) — LinearLiquidus melting_temperature,
//       SemiImplicitStress.
//
// Type flow: the reference is generic Julia; mixed Float32/Float64 expressions promote
// exactly like C++'s usual arithmetic conversions (float∘double→double, int∘T→T), so each
// expression below is written with variables of the element type they have in the reference
// (FT exchange grid, CT thermodynamics, double for the Float64 literal viscosity,
// roughness_lengths.jl:94,126) and `auto` results.  Compile with -ffp-contract=off: Julia
// does not contract a*b+c.
//
// All citations are relative to /root/reference/src unless stated otherwise.

#include "../include/ne_b200.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <type_traits>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---- op census: counts transcendental calls when enabled ------------------
struct OpCounts { uint64_t exp_, log_, pow_, atan_, cbrt_, sqrt_, iters, points; };
thread_local OpCounts g_ops = {0, 0, 0, 0, 0, 0, 0, 0};
bool g_count_ops = false;
#define NEO_COUNT(field) do { if (g_count_ops) ++g_ops.field; } while (0)

// Float32 transcendentals: evaluated in Float64 and rounded once = the correctly rounded Float32 function (up to double
// rounding).  Julia's Float32 `^` IS that (Base.Math.pow_body: T(exp2(log2(abs(widen(x))) * y))), its Float32 exp / log /
// cbrt are within half an ulp of it; glibc's powf / logf are not (0.82 ulp), so they are not used.
inline float  m_exp(float x)  { NEO_COUNT(exp_);  return (float)std::exp((double)x); }
inline double m_exp(double x) { NEO_COUNT(exp_);  return std::exp(x); }
inline float  m_log(float x)  { NEO_COUNT(log_);  return (float)std::log((double)x); }
inline double m_log(double x) { NEO_COUNT(log_);  return std::log(x); }
inline float  m_atan(float x) { NEO_COUNT(atan_); return (float)std::atan((double)x); }
inline double m_atan(double x){ NEO_COUNT(atan_); return std::atan(x); }
inline float  m_cbrt(float x) { NEO_COUNT(cbrt_); return (float)std::cbrt((double)x); }
inline double m_cbrt(double x){ NEO_COUNT(cbrt_); return std::cbrt(x); }
inline float  m_sqrt(float x) { NEO_COUNT(sqrt_); return std::sqrt(x); }
inline double m_sqrt(double x){ NEO_COUNT(sqrt_); return std::sqrt(x); }
inline float  m_pow(float x, float y)   { NEO_COUNT(pow_); return (float)std::pow((double)x, (double)y); }
inline double m_pow(double x, double y) { NEO_COUNT(pow_); return std::pow(x, y); }
inline double m_pow(float x, double y)  { NEO_COUNT(pow_); return std::pow((double)x, y); }
inline double m_pow(double x, float y)  { NEO_COUNT(pow_); return std::pow(x, (double)y); }

template <class T> inline T m_min(T a, T b) { return (b < a) ? b : a; }   // Julia min for non-NaN
template <class T> inline T m_max(T a, T b) { return (a < b) ? b : a; }
template <class A, class B> inline auto mn(A a, B b) { using W = decltype(a + b); return m_min<W>((W)a, (W)b); }
template <class A, class B> inline auto mx(A a, B b) { using W = decltype(a + b); return m_max<W>((W)a, (W)b); }
template <class T> inline T m_clamp(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }

// Julia x^2 / x^3 are literal_pow => x*x, x*x*x.  Higher integer powers go through
// Base.power_by_squaring-like pow_body (Float32: evaluated in Float64 and rounded).
template <class T> inline T sq(T x) { return x * x; }
template <class T> inline T cube(T x) { return x * x * x; }
inline double ipow(double x, int n) { double r = 1, b = x; while (n) { if (n & 1) r *= b; b *= b; n >>= 1; } return r; }
inline float ipow(float x, int n) { return (float)ipow((double)x, n); }

// ---- exchange layout -----------------------------------------------------------------------
struct Layout {
  int64_t sx, hx, hy;
  explicit Layout(const NeExchangeGrid& g) : sx(g.nx + 2 * g.hx), hx(g.hx), hy(g.hy) {}
  inline int64_t at(int64_t i, int64_t j) const { return (i + hx - 1) + (j + hy - 1) * sx; }
};

template <class FT> inline FT slot_at(const NeSlot& s, int64_t idx) {
  return s.ptr ? static_cast<const FT*>(s.ptr)[idx] : static_cast<FT>(s.value);
}

// ---- thermodynamics [3rd-party: Thermodynamics.jl], parameters pinned by
// Atmospheres/thermodynamic_parameters.jl:30-258 ------------------------------------------------
template <class CT> struct Thermo {
  CT R, Md, Mv, kappa_d, cp_v, cp_l, cp_i, LH_v0, LH_s0, T_0, T_triple, press_triple;
  CT R_d, R_v, eps, cp_d;  // derived in CT arithmetic: thermodynamic_parameters.jl:73-78, 256
  explicit Thermo(const NeThermoParams& p) {
    R = (CT)p.gas_constant; Md = (CT)p.dry_air_molar_mass; Mv = (CT)p.water_molar_mass;
    kappa_d = (CT)p.kappa_d; cp_v = (CT)p.cp_v; cp_l = (CT)p.cp_l; cp_i = (CT)p.cp_i;
    LH_v0 = (CT)p.LH_v0; LH_s0 = (CT)p.LH_s0; T_0 = (CT)p.T_0; T_triple = (CT)p.T_triple;
    press_triple = (CT)p.press_triple;
    R_d = R / Md; R_v = R / Mv; eps = Md / Mv; cp_d = R_d / kappa_d;
  }
  // Clausius–Clapeyron with constant Δcp (docs/src/interface_fluxes.md:94-96)
  CT saturation_vapor_pressure(CT T, int phase) const {
    CT LH_0 = phase == NE_PHASE_LIQUID ? LH_v0 : LH_s0;
    CT dcp = phase == NE_PHASE_LIQUID ? (cp_v - cp_l) : (cp_v - cp_i);
    return press_triple * m_pow(T / T_triple, dcp / R_v) *
           m_exp((LH_0 - dcp * T_0) / R_v * (1 / T_triple - 1 / T));
  }
  template <class Q> auto gas_constant_air(Q q) const { return R_d * (1 - q) + R_v * q; }  // docs/src/interface_fluxes.md:507
  template <class T, class P, class Q> auto air_density(T temp, P p, Q q) const { return p / (gas_constant_air(q) * temp); }
  template <class Q> auto cp_m(Q q) const { return cp_d * (1 - q) + cp_v * q; }
  template <class T> auto latent_heat_vapor(T temp) const { return LH_v0 + (cp_v - cp_l) * (temp - T_0); }
  template <class T> auto latent_heat_sublim(T temp) const { return LH_s0 + (cp_v - cp_i) * (temp - T_0); }
  template <class T, class Q> auto virtual_temperature(T temp, Q q) const { return temp * gas_constant_air(q) / R_d; }
};

// ---- interface humidity: EarthSystemModels/InterfaceComputations/interface_states.jl:44-74, 236-277
template <class FT> inline FT water_mole_fraction(const NeInterfaceProperties& ip, FT S) {
  FT s = S / 1000;
  FT mu_w = (FT)ip.water_molar_mass;
  FT mu0 = (FT)ip.constituent_molar_mass[0], mu1 = (FT)ip.constituent_molar_mass[1];
  FT mu2 = (FT)ip.constituent_molar_mass[2], mu3 = (FT)ip.constituent_molar_mass[3];
  FT e0 = (FT)ip.constituent_mass_fraction[0], e1 = (FT)ip.constituent_mass_fraction[1];
  FT e2 = (FT)ip.constituent_mass_fraction[2], e3 = (FT)ip.constituent_mass_fraction[3];
  FT alpha = mu_w * (e0 / mu0 + e1 / mu1 + e2 / mu2 + e3 / mu3);
  return (1 - s) / (1 - s + alpha * s);
}

template <class FT, class CT>
inline FT surface_specific_humidity(const NeInterfaceProperties& ip, const Thermo<CT>& th, FT p_at, FT Ts, FT Ss) {
  CT T = (CT)Ts;
  CT p = (CT)p_at;
  CT psat = th.saturation_vapor_pressure(T, ip.phase);
  // χ·p_sat: χ is the Int 1, an FT constant, or an FT function of salinity (:46-47, 255-277)
  using W = decltype(FT() * CT());
  W pv;
  if (ip.x_h2o_kind == NE_XH2O_ONE) pv = psat;
  else if (ip.x_h2o_kind == NE_XH2O_CONSTANT) pv = (FT)ip.x_h2o * psat;
  else pv = water_mole_fraction<FT>(ip, Ss) * psat;
  CT eps_inv = 1 / th.eps;
  pv = mn(pv, (CT)0.999 * p);
  auto q = eps_inv * pv / (p - (1 - eps_inv) * pv);
  return (FT)q;
}

// ---- stability functions: similarity_theory_turbulent_fluxes.jl:445-752 ---------------------
template <class FT, class Z> auto stability_fn(const NeStabilityFn& f, Z zeta) -> decltype(FT() * Z()) {
  using W = decltype(FT() * Z());
  const double* p = f.p;
  switch (f.kind) {
    case NE_PSI_ZERO: return (W)0;
    case NE_PSI_EDSON_MOMENTUM: {  // :501-532
      FT zmax = (FT)p[0], Ap = (FT)p[1], Bp = (FT)p[2], Cp = (FT)p[3], Dp = (FT)p[4];
      FT Am = (FT)p[5], Bm = (FT)p[6], Cm = (FT)p[7], Dm = (FT)p[8], Em = (FT)p[9], Fm = (FT)p[10];
      Z zm = m_min<Z>(0, zeta), zp = m_max<Z>(0, zeta);
      auto dz = mn(zmax, Ap * zp);
      auto psi_p = -Bp * zp - Cp * (zp - Dp) * m_exp(-dz) - Cp * Dp;
      auto f1 = m_sqrt(m_sqrt(1 - Am * zm));
      auto psi1 = Bm * m_log((1 + f1) / Bm) + m_log((1 + sq(f1)) / Bm) - Bm * m_atan(f1) + Cm;
      auto f2 = m_cbrt(1 - Dm * zm);
      auto psi2 = Em / 2 * m_log((1 + f2 + sq(f2)) / Em) - m_sqrt(Em) * m_atan((1 + 2 * f2) / m_sqrt(Em)) + Fm;
      auto fw = sq(zm) / (1 + sq(zm));
      auto psi_m = (1 - fw) * psi1 + fw * psi2;
      return zeta < 0 ? (W)psi_m : (W)psi_p;
    }
    case NE_PSI_EDSON_SCALAR: {  // :586-618
      FT zmax = (FT)p[0], Ap = (FT)p[1], Bp = (FT)p[2], Cp = (FT)p[3], Dp = (FT)p[4], Ep = (FT)p[5];
      FT Am = (FT)p[6], Bm = (FT)p[7], Cm = (FT)p[8], Dm = (FT)p[9], Em = (FT)p[10], Fm = (FT)p[11];
      Z zm = m_min<Z>(0, zeta), zp = m_max<Z>(0, zeta);
      auto dz = mn(zmax, Ap * zp);
      auto psi_p = -m_pow(1 + Bp * zp, Cp) - Bp * (zp - Dp) * m_exp(-dz) - Ep;
      auto f1 = m_sqrt(1 - Am * zm);
      auto psi1 = Bm * m_log((1 + f1) / Bm) + Cm;
      auto f2 = m_cbrt(1 - Dm * zm);
      auto psi2 = Em / 2 * m_log((1 + f2 + sq(f2)) / Em) - m_sqrt(Em) * m_atan((1 + 2 * f2) / m_sqrt(Em)) + Fm;
      auto fw = sq(zm) / (1 + sq(zm));
      auto psi_m = (1 - fw) * psi1 + fw * psi2;
      return zeta < 0 ? (W)psi_m : (W)psi_p;
    }
    case NE_PSI_SHEBA_MOMENTUM: {  // :643-657 ; rt3 = sqrt(3) is Float64 in the reference
      FT a = (FT)p[0], b = (FT)p[1];
      Z zp = m_max<Z>(0, zeta);
      auto z = m_cbrt(1 + zp);
      FT B = m_cbrt((1 - b) / b);
      double rt3 = std::sqrt(3.0);
      auto P1 = -3 * a * (z - 1) / b;
      auto P2 = a * B / (2 * b) * (2 * m_log((z + B) / (1 + B)) - m_log((sq(z) - B * z + sq(B)) / (1 - B + sq(B))) +
                                 2 * rt3 * (m_atan((2 * z - B) / (rt3 * B)) - m_atan((2 - B) / (rt3 * B))));
      return (W)(P1 + P2);
    }
    case NE_PSI_SHEBA_SCALAR: {  // :665-677
      FT a = (FT)p[0], b = (FT)p[1], c = (FT)p[2];
      FT B = m_sqrt(sq(c) - 4);
      Z zp = m_max<Z>(0, zeta);
      auto P1 = -b / 2 * m_log(1 + c * zp + sq(zp));
      auto P2 = (b * c / (2 * B) - a / B) * (m_log((2 * zp + c - B) / (2 * zp + c + B)) - m_log((c - B) / (c + B)));
      return (W)(P1 + P2);
    }
    case NE_PSI_PAULSON_MOMENTUM: {  // :688-699
      FT a = (FT)p[0], b = (FT)p[1];
      Z zm = m_min<Z>(0, zeta);
      auto z = m_sqrt(m_sqrt(1 - a * zm));
      auto P1 = 2 * m_log((1 + z) / 2);
      auto P2 = m_log((1 + sq(z)) / 2);
      auto P3 = -2 * m_atan(z);
      return (W)(P1 + P2 + P3 + b);
    }
    case NE_PSI_PAULSON_SCALAR: {  // :705-710
      FT a = (FT)p[0];
      Z zm = m_min<Z>(0, zeta);
      auto z = m_sqrt(m_sqrt(1 - a * zm));
      return (W)(2 * m_log((1 + sq(z)) / 2));
    }
    case NE_PSI_LINEAR_STABLE: {  // :747-752
      FT c = (FT)p[0], zmax = (FT)p[1];
      Z zp = m_max<Z>(0, zeta);
      return (W)(-c * mn(zp, zmax));
    }
  }
  return std::numeric_limits<W>::quiet_NaN();
}

template <class FT, class Z> auto stability_profile(const NeStabilityProfile& s, Z zeta) -> decltype(FT() * Z()) {
  if (!s.split) return stability_fn<FT, Z>(s.a, zeta);
  auto st = stability_fn<FT, Z>(s.a, zeta);   // :720-725 evaluates both
  auto un = stability_fn<FT, Z>(s.b, zeta);
  return zeta > 0 ? st : un;
}

// similarity_profile :242-253
template <class FT, class H, class L, class LS>
auto similarity_profile(int form, const NeStabilityProfile& psi, H h, L ell, LS Lstar) {
  auto zeta = h / Lstar;
  auto psi_h = stability_profile<FT>(psi, zeta);
  if (form == NE_PROFILE_COARE) return m_log(h / ell) - psi_h;
  auto psi_l = stability_profile<FT>(psi, ell / Lstar);
  return m_log(h / ell) - psi_h + psi_l;
}

// ---- roughness lengths: roughness_lengths.jl:182-246 -----------------------------------------
// VT = type of the kinematic viscosity: double for the constant Float64 literal (:94,126),
// FT for TemperatureDependentAirViscosity{FT} (:185-189).
template <class FT, class T> inline FT temperature_dependent_viscosity(const NeRoughnessLength& r, T Tk) {
  FT Tp = (FT)(Tk - 273.15);  // celsius_to_kelvin is a Float64 const (components.jl:10)
  FT C0 = (FT)r.nu_C[0], C1 = (FT)r.nu_C[1], C2 = (FT)r.nu_C[2], C3 = (FT)r.nu_C[3];
  return C0 + C1 * Tp + C2 * sq(Tp) + C3 * cube(Tp);
}

template <class FT, class VT, class US, class UU>
auto momentum_roughness(const NeRoughnessLength& r, VT nu, US ustar, UU U) {
  FT g = (FT)r.gravitational_acceleration;
  FT Cnu = (FT)r.smooth_wall_parameter;
  FT lmax = (FT)r.maximum_roughness_length;
  using WG = decltype(FT() * UU());
  WG Cg;
  if (r.wave_kind == NE_WAVE_CONSTANT) Cg = (FT)r.wave_constant;
  else Cg = mx((UU)0, (FT)r.wave_C1 * mn(U, (FT)r.wave_Umax) + (FT)r.wave_C2);  // :75
  auto lW = Cg * sq(ustar) / g;
  using WR = decltype(Cnu * nu / ustar);
  WR lR = (Cnu == 0) ? (WR)0 : Cnu * nu / ustar;
  auto lstar = lW + lR;
  return mn(lstar, lmax);
}

template <class FT, class VT, class LU, class US>
auto scalar_roughness(const NeRoughnessLength& r, VT nu, LU ell_u, US ustar) {
  auto Rstar = ell_u * ustar / nu;
  FT A = (FT)r.reynolds_A, b = (FT)r.reynolds_b;
  using WR = decltype(A / m_pow(Rstar, b));
  WR ls = (Rstar == 0) ? (WR)0 : A / m_pow(Rstar, b);   // :231
  FT lmax = (FT)r.maximum_roughness_length;
  return mn(ls, lmax);
}

// ---- subgrid velocities :88-98 ------------------------------------------------------------
template <class FT, class US, class BS, class HB>
auto vsgs2_one(int kind, const NeSubgridVelocity& s, double constant, US ustar, BS bstar, HB h_bl)
    -> decltype(FT() * US() * BS() * HB()) {
  using W = decltype(FT() * US() * BS() * HB());
  if (kind == NE_SGS_NONE) return (W)0;
  if (kind == NE_SGS_CONSTANT) { FT v = (FT)constant; return (W)sq(v); }
  auto Jb = -ustar * bstar;
  using J = decltype(Jb);
  auto UG = mx((FT)s.minimum_gustiness, (FT)s.gustiness_parameter * m_cbrt(m_max<J>(0, Jb) * h_bl));
  return (W)sq(UG);
}
template <class FT, class US, class BS, class HB>
auto vsgs2(const NeSubgridVelocity& s, US ustar, BS bstar, HB h_bl) {
  auto c = vsgs2_one<FT>(s.convective_kind, s, s.convective_constant, ustar, bstar, h_bl);
  if (!s.composite) return c;
  return c + vsgs2_one<FT>(s.mesoscale_kind, s, s.mesoscale_constant, ustar, bstar, h_bl);
}

// buoyancy_scale :417-425
template <class CT, class A, class B, class C, class D, class G>
auto buoyancy_scale(A theta_star, B q_star, const Thermo<CT>& th, C Ts, D qs, G g) {
  auto Tv = th.virtual_temperature(Ts, qs);
  CT delta = th.eps - 1;
  return g / Tv * (theta_star * (1 + delta * qs) + delta * Tv * q_star);
}

// PolynomialNeutralDragCoefficient functor: coefficient_based_turbulent_fluxes.jl:46-52
template <class FT, class UU> auto polynomial_drag(const NePolynomialDrag& p, UU U) {
  auto Um = mx(U, (FT)p.minimum_wind_speed);
  using W = decltype(Um);
  W poly = ((FT)p.a / Um + (FT)p.b + (FT)p.c * Um - (FT)p.d * ipow(Um, 6)) / 1000;
  return (Um < (FT)p.high_wind_speed_threshold) ? poly : (W)(FT)p.high_wind_drag_coefficient;
}

// ---- the interface state carried through the iteration: interface_states.jl:664-803 ----------
template <class FT> struct State { FT ustar, theta_star, q_star, u, v, T, q, S; };

template <class FT> struct AtmosState { FT z, u, v, T, p, q, h_bl; };
template <class FT> struct Interior { FT u, v, T, S, kappa, hi, hs, hc; };
template <class FT> struct RadState { FT sigma, alpha, eps, sw, lw; };

template <class FT, class CT>
auto surface_atmosphere_temperature(const AtmosState<FT>& a, const Thermo<CT>& th, FT g) {  // interface_states.jl:308-317
  auto c = th.cp_m(a.q);
  return a.T + g * a.z / c;
}

// ---- temperature formulations: interface_states.jl:330-577 -----------------------------------
template <class FT, class CT>
FT interface_temperature(const NeInterfaceProperties& ip, const State<FT>& s, const AtmosState<FT>& a,
                         const Interior<FT>& in, const RadState<FT>& rad, const Thermo<CT>& th, FT g,
                         const NeMediumProperties& medium) {
  if (ip.temperature_formulation == NE_TEMP_BULK) return s.T;  // :333
  // compute_interface_temperature(::SkinTemperature, ...) :526-577
  auto rho_a = th.air_density(a.T, a.p, a.q);
  auto c_a = th.cp_m(a.q);
  auto Li = th.latent_heat_sublim(a.T);   // sublimation for EVERY surface (:542-544)
  FT Tsm = s.T;
  FT lw_up = rad.sigma * rad.eps * ipow(Tsm, 4);
  FT Qd = -(1 - rad.alpha) * rad.sw - rad.eps * rad.lw;
  auto QT = -rho_a * c_a * s.ustar * s.theta_star;
  auto Qv = -rho_a * Li * s.ustar * s.q_star;
  auto Tat = surface_atmosphere_temperature(a, th, g);

  if (ip.temperature_formulation == NE_TEMP_SKIN_DIFFUSIVE || ip.temperature_formulation == NE_TEMP_SKIN_DIFFUSIVE_INTERIOR) {
    // flux_balance_temperature(::SkinTemperature{<:DiffusiveFlux}) :434-457
    FT kappa = ip.temperature_formulation == NE_TEMP_SKIN_DIFFUSIVE ? (FT)ip.kappa : m_max<FT>(in.kappa, (FT)ip.kappa);
    FT delta = (FT)ip.delta;
    FT rho = (FT)medium.reference_density, c = (FT)medium.heat_capacity;
    auto Qa = Qv + lw_up + Qd;
    FT lambda = 1 / (rho * c);
    auto JT = Qa * lambda;
    auto dT = Tat - s.T;
    auto Om = QT * lambda;
    auto D = kappa * dT - Om * delta;
    auto Tstar = (in.T * kappa * dT - (JT * dT + Om * Tat) * delta) / D;
    using W = decltype(Tstar);
    Tstar = (D == 0) ? (W)s.T : Tstar;
    FT maxdT = (FT)ip.max_dT;
    return (FT)(in.T + m_clamp<W>(Tstar - in.T, -maxdT, maxdT));
  }
  // conductive_flux_balance_temperature :468-508 (sea ice); medium = sea_ice_properties
  FT R;
  if (ip.temperature_formulation == NE_TEMP_SKIN_CONDUCTIVE) R = in.hi / (FT)ip.ice_conductivity;          // :511-516
  else R = in.hs / (FT)ip.snow_conductivity + in.hi / (FT)ip.ice_conductivity;                             // :519-524
  // [3rd-party: ClimaSeaIce LinearLiquidus] Tm = T_fresh - slope * S
  FT Tb = (FT)medium.liquidus_freshwater_melting_temperature - (FT)medium.liquidus_slope * in.S;
  if (medium.temperature_units == NE_DEGREES_CELSIUS) Tb = Tb + (FT)273.15;
  auto dT = Tat - Tsm;
  auto Qa = Qv + lw_up + Qd;
  using W = decltype(QT / dT);
  W Oc = (dT == 0) ? (W)0 : QT / dT;
  FT beta = 4 * lw_up / Tsm;
  auto D = 1 + beta * R - Oc * R;
  auto Tstar = (Tb + beta * R * Tsm - Oc * R * Tat - Qa * R) / D;
  using WT = decltype(Tstar);
  Tstar = (D == 0) ? (WT)Tsm : Tstar;
  Tstar = std::isnan(Tstar) ? (WT)Tsm : Tstar;
  auto dTs = Tstar - Tsm;
  WT maxdT = (WT)ip.max_dT;
  auto Tsp = Tsm + m_clamp<WT>(dTs, -maxdT, maxdT);
  FT Tm = (FT)medium.liquidus_freshwater_melting_temperature;
  if (medium.temperature_units == NE_DEGREES_CELSIUS) Tm = Tm + (FT)273.15;
  Tsp = mn(Tsp, Tm);
  Tsp = (in.hi >= in.hc) ? Tsp : (WT)Tb;
  return (FT)Tsp;
}

// ---- iterate_interface_fluxes(::SimilarityTheoryFluxes) :315-385 -------------------------------
template <class FT, class CT, class VT>
void iterate_similarity(const NeFluxFormulation& ff, const NeInterfaceProperties& ip, const Thermo<CT>& th, FT g,
                        FT Ts, FT qs, decltype(FT() + CT()) dtheta, FT dq, FT dh0,
                        const State<FT>& s, const AtmosState<FT>& a, FT& us_out, FT& ts_out, FT& qs_out) {
  FT ustar = s.ustar, theta_star = s.theta_star, q_star = s.q_star;
  auto bstar = buoyancy_scale(theta_star, q_star, th, Ts, qs, g);
  auto Usg2 = vsgs2<FT>(ff.subgrid_velocities, ustar, bstar, a.h_bl);
  FT du, dv;
  if (ip.velocity_formulation == NE_VEL_RELATIVE) { du = a.u - s.u; dv = a.v - s.v; } else { du = a.u; dv = a.v; }
  auto U = m_sqrt(sq(du) + sq(dv) + Usg2);

  // air viscosity per roughness-length struct (roughness_lengths.jl:182-189)
  auto visc = [&](const NeRoughnessLength& r) -> VT {
    if (r.visc_kind == NE_VISC_CONSTANT) return (VT)r.nu;
    return (VT)temperature_dependent_viscosity<FT>(r, Ts);
  };
  using LW = decltype(FT() * VT() * U);
  LW lu, lq, lt;
  if (ff.ell_momentum.kind == NE_ROUGH_CONSTANT) lu = (FT)ff.ell_momentum.constant;
  else lu = momentum_roughness<FT, VT>(ff.ell_momentum, visc(ff.ell_momentum), ustar, U);
  if (ff.ell_water_vapor.kind == NE_ROUGH_CONSTANT) lq = (FT)ff.ell_water_vapor.constant;
  else lq = scalar_roughness<FT, VT>(ff.ell_water_vapor, visc(ff.ell_water_vapor), lu, ustar);
  if (ff.ell_temperature.kind == NE_ROUGH_CONSTANT) lt = (FT)ff.ell_temperature.constant;
  else lt = scalar_roughness<FT, VT>(ff.ell_temperature, visc(ff.ell_temperature), lu, ustar);

  FT d = (FT)ff.zero_plane_displacement;
  auto dh = mx(dh0 - d, 2 * lu);  // displaced_profile_height :313
  FT kappa = (FT)ff.von_karman_constant;
  using BW = decltype(sq(ustar) / (kappa * bstar));
  BW Lstar = (bstar == 0) ? std::numeric_limits<BW>::infinity() : sq(ustar) / (kappa * bstar);

  auto chi_u = kappa / similarity_profile<FT>(ff.similarity_form, ff.psi_momentum, dh, lu, Lstar);
  auto chi_t = kappa / similarity_profile<FT>(ff.similarity_form, ff.psi_temperature, dh, lt, Lstar);
  auto chi_q = kappa / similarity_profile<FT>(ff.similarity_form, ff.psi_water_vapor, dh, lq, Lstar);

  us_out = (FT)(chi_u * U);
  ts_out = (FT)(chi_t * dtheta);
  qs_out = (FT)(chi_q * dq);
}

// ---- evaluate_coefficients(::LargeYeagerTransferCoefficients) + CoefficientBased iterate
// coefficient_based_turbulent_fluxes.jl:265-371 ---------------------------------------------------
template <class FT, class CT>
void iterate_coefficient(const NeFluxFormulation& ff, const NeInterfaceProperties& ip, const Thermo<CT>& th, FT g,
                         FT Ts, FT qs, decltype(FT() + CT()) dtheta, FT dq, FT dh,
                         const State<FT>& s, const AtmosState<FT>& a, FT& us_out, FT& ts_out, FT& qs_out) {
  FT du, dv;
  if (ip.velocity_formulation == NE_VEL_RELATIVE) { du = a.u - s.u; dv = a.v - s.v; } else { du = a.u; dv = a.v; }
  using W = decltype(FT() + CT());
  W Cd, Ch, Cq;
  W dU;
  if (ff.kind == NE_FLUX_LARGE_YEAGER) {
    const NeLargeYeager& ly = ff.large_yeager;
    FT Umin = (FT)ly.neutral_drag.minimum_wind_speed;
    dU = m_max<FT>(m_sqrt(sq(du) + sq(dv)), Umin);
    FT kap = (FT)ly.von_karman_constant;
    FT ustar = s.ustar, theta_star = s.theta_star, q_star = s.q_star;
    FT h0 = (FT)ly.reference_height;
    W dUm = m_max<W>(dU, Umin);
    auto bstar = buoyancy_scale(theta_star, q_star, th, Ts, qs, g);
    using BW = decltype(sq(ustar) / (kap * bstar));
    BW Lstar = (bstar == 0) ? (BW)std::numeric_limits<FT>::infinity() : sq(ustar) / (kap * bstar);
    auto zeta = dh / Lstar;
    auto psi_m = stability_profile<FT>(ly.psi_momentum, zeta);
    auto psi_h = stability_profile<FT>(ly.psi_temperature, zeta);
    W Cdp = (ustar == 0) ? (W)polynomial_drag<FT>(ly.neutral_drag, dUm) : (W)(sq(ustar) / sq(dUm));
    W UN10 = dUm / (1 + m_sqrt(Cdp) / kap * (m_log(dh / h0) - psi_m));
    UN10 = m_max<W>(UN10, Umin);
    W CdN = polynomial_drag<FT>(ly.neutral_drag, UN10);
    bool stable = zeta > 0;
    W ChN = m_sqrt(CdN) / 1000 * (stable ? (FT)ly.stable_heat : (FT)ly.unstable_heat);
    W CqN = m_sqrt(CdN) / 1000 * (FT)ly.moisture;
    W xi_m = m_sqrt(CdN) / kap * (m_log(dh / h0) - psi_m);
    Cd = CdN / sq(1 + xi_m);
    W xi_h = m_sqrt(CdN) / kap * (m_log(dh / h0) - psi_h);
    W ratio = m_sqrt(Cd) / m_sqrt(CdN);
    Ch = ChN * ratio / (1 + ChN * xi_h);
    Cq = CqN * ratio / (1 + CqN * xi_h);
  } else {
    dU = m_max<FT>(m_sqrt(sq(du) + sq(dv)), 0);   // minimum_wind_speed(::Tuple) = 0 :260-262
    W c[3];
    for (int k = 0; k < 3; ++k) {
      const NeTransferCoefficient& tc = ff.coefficients[k];
      c[k] = tc.kind == NE_COEFF_CONSTANT ? (W)(FT)tc.constant : (W)polynomial_drag<FT>(tc.polynomial, dU);
    }
    Cd = c[0]; Ch = c[1]; Cq = c[2];
  }
  us_out = (FT)(m_sqrt(Cd) * dU);
  ts_out = (Cd == 0) ? (FT)0 : (FT)(Ch / m_sqrt(Cd) * dtheta);
  qs_out = (Cd == 0) ? (FT)0 : (FT)(Cq / m_sqrt(Cd) * dq);
}

// ---- land surface humidity closures (interface_states.jl:76-229, 585-651) ------------------------------
// saturation_specific_humidity (:79-90): pressure-based, in the thermodynamics' element type
template <class CT, class T, class P>
inline CT saturation_specific_humidity(const Thermo<CT>& th, T Ts, P p_at, int phase) {
  CT Tc = (CT)Ts, p = (CT)p_at;
  CT pv = th.saturation_vapor_pressure(Tc, phase);
  pv = m_min<CT>(pv, (CT)0.999 * p);
  const CT eps_inv = 1 / th.eps;
  return eps_inv * pv / (p - (1 - eps_inv) * pv);
}

template <class FT> struct LandSurface {
  const NeLandHumidity* h;   // nullptr: not a land interface
  FT saturation;             // Ψ.hydrology.saturation
  FT T_bulk;                 // Ψ.energy.temperature (the SkinHumidity reservoir)
};

// compute_interface_humidity for AirLandInterfaceState: BulkHumidity :120-126 (via :587-588), FractionalHumidity
// :592-598 with evaporation_efficiency :151-157, SkinHumidity :625-651 (previous iterate's u★, q★, qˢ)
template <class FT, class CT>
inline FT land_interface_humidity(const LandSurface<FT>& land, const Thermo<CT>& th, const State<FT>& s,
                                  const AtmosState<FT>& a, FT Ts) {
  const NeLandHumidity& h = *land.h;
  if (h.kind == NE_LANDQ_BULK) {
    CT qv = saturation_specific_humidity<CT>(th, Ts, a.p, h.phase);
    return (FT)((land.saturation > 0) ? qv : (CT)0);
  }
  if (h.kind == NE_LANDQ_FRACTIONAL_CRITICAL) {
    FT beta = m_min<FT>(land.saturation / (FT)h.critical_saturation, (FT)1);
    CT qv = saturation_specific_humidity<CT>(th, Ts, a.p, h.phase);
    return (FT)(beta * qv);
  }
  if (h.kind == NE_LANDQ_FRACTIONAL_CONSTANT) {
    CT qv = saturation_specific_humidity<CT>(th, Ts, a.p, h.phase);
    return (FT)(h.efficiency * qv);   // β::Number keeps its own (Float64) type
  }
  if (h.kind == NE_LANDQ_DRY_LAYER) {   // compute_interface_humidity(::DryLayerHumidity) dry_layer_humidity.jl
    auto rho_a = th.air_density(a.T, a.p, a.q);
    const FT S = land.saturation, Tin = Ts, Tla = land.T_bulk;
    const FT sc = m_min<FT>(S / (FT)h.dry_layer_onset_saturation, (FT)1);
    const FT dv = (FT)h.maximum_dry_layer_depth * m_pow(m_max<FT>((FT)1 - sc, (FT)0), (FT)h.dry_layer_exponent);
    const FT dvmin = (FT)h.minimum_dry_layer_depth, lT = (FT)h.thermal_exchange_depth;
    const FT chi = m_clamp<FT>(dv / lT, (FT)0, (FT)1);
    const FT Te = Tin + chi * (Tla - Tin);
    CT qe = saturation_specific_humidity<CT>(th, Te, a.p, h.phase);
    const FT theta_l = S * (FT)h.porosity;
    FT Dv;
    if (h.tortuosity == NE_TORTUOSITY_CONSTANT) Dv = (FT)h.molecular_diffusivity;
    else {
      const FT nu = (FT)h.porosity, tg = m_max<FT>(nu - theta_l, (FT)0);
      Dv = (FT)h.molecular_diffusivity * m_pow(tg, (FT)10 / (FT)3) / (nu * nu);
    }
    auto Ge = rho_a * Dv / m_max<FT>(dv, dvmin);
    auto Ja = -rho_a * s.ustar * s.q_star;
    FT dq = s.q - a.q;
    auto D = Ge * dq + Ja;
    auto qbal = (D == 0) ? (decltype((Ge * qe * dq + Ja * a.q) / D))s.q : (Ge * qe * dq + Ja * a.q) / D;
    CT qinp = saturation_specific_humidity<CT>(th, Tin, a.p, h.phase);
    const FT dvw = (FT)h.wet_transition_width;
    const FT z = 10 * (dv - dvmin - dvw / 2) / m_max<FT>(dvw, std::numeric_limits<FT>::epsilon());
    const FT sigma = 1 / (1 + m_exp(-z));
    return (FT)(qinp + sigma * (qbal - qinp));
  }
  // SkinHumidity
  auto rho_a = th.air_density(a.T, a.p, a.q);
  CT qv = saturation_specific_humidity<CT>(th, land.T_bulk, a.p, h.phase);
  double gs = h.vapor_diffusivity / h.surface_thickness;   // κ / d, the user's numbers
  auto Ja = -rho_a * s.ustar * s.q_star;
  FT dq = s.q - a.q;
  auto D = gs * dq + Ja;
  auto qs = (gs * qv * dq + Ja * a.q) / D;
  return (FT)((D == 0) ? (decltype(qs))s.q : qs);
}

// ---- iterate_interface_state: compute_interface_state.jl:69-122 ---------------------------------
template <class FT, class CT, class VT>
State<FT> iterate_interface_state(const NeFluxFormulation& ff, const NeInterfaceProperties& ip, const Thermo<CT>& th,
                                  FT g, const State<FT>& s, const AtmosState<FT>& a, const Interior<FT>& in,
                                  const RadState<FT>& rad, const NeMediumProperties& medium, bool ice_state,
                                  const LandSurface<FT>* land = nullptr) {
  FT Ts = interface_temperature<FT, CT>(ip, s, a, in, rad, th, g, medium);
  // humidity_surface_scalar: salinity for AirSea, 0 for AirIce (interface_states.jl:717,737); land closures for AirLand
  FT qs = land ? land_interface_humidity<FT, CT>(*land, th, s, a, Ts)
               : surface_specific_humidity<FT, CT>(ip, th, a.p, Ts, ice_state ? (FT)0 : s.S);
  FT dq = a.q - qs;
  auto theta_a = surface_atmosphere_temperature(a, th, g);
  auto dtheta = theta_a - Ts;
  FT dh = a.z;
  State<FT> n = s;
  if (ff.kind == NE_FLUX_SIMILARITY_THEORY)
    iterate_similarity<FT, CT, VT>(ff, ip, th, g, Ts, qs, dtheta, dq, dh, s, a, n.ustar, n.theta_star, n.q_star);
  else
    iterate_coefficient<FT, CT>(ff, ip, th, g, Ts, qs, dtheta, dq, dh, s, a, n.ustar, n.theta_star, n.q_star);
  n.T = Ts;
  n.q = qs;
  return n;
}

// compute_interface_state + iterating: compute_interface_state.jl:5-58
template <class FT, class CT, class VT>
State<FT> compute_interface_state(const NeFluxFormulation& ff, const NeInterfaceProperties& ip, const Thermo<CT>& th,
                                  FT g, const State<FT>& init, const AtmosState<FT>& a, const Interior<FT>& in,
                                  const RadState<FT>& rad, const NeMediumProperties& medium, bool ice_state,
                                  int& iterations, const LandSurface<FT>* land = nullptr) {
  State<FT> cur = init, prev = init;
  int it = 0;
  for (;;) {
    bool go;
    if (ff.stop.kind == NE_STOP_FIXED_ITERATIONS) {
      go = it < ff.stop.maxiter;
    } else {
      bool hasnt_started = it == 0;
      bool reached = it >= ff.stop.maxiter;
      FT drift = std::fabs(cur.ustar - prev.ustar) + std::fabs(cur.theta_star - prev.theta_star) +
                 std::fabs(cur.q_star - prev.q_star);
      bool converged = drift < (FT)ff.stop.tolerance;
      go = !(converged | reached) | hasnt_started;
    }
    if (!go) break;
    prev = cur;
    cur = iterate_interface_state<FT, CT, VT>(ff, ip, th, g, prev, a, in, rad, medium, ice_state, land);
    ++it;
  }
  iterations = it;
  if (g_count_ops) { g_ops.iters += it; g_ops.points += 1; }
  return cur;
}

template <class T> inline T clampv(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }   // Base.clamp

// Oceananigans interpolator(fractional_idx) = (unsafe_trunc(Int, f) + 1, i⁻ + Int(sign(f)), mod(f, 1)) [3rd party]
template <class T> inline void interpolator1(T f, int64_t& im, int64_t& ip, T& xi) {
  im = (int64_t)f + 1;
  ip = im + ((f > 0) ? 1 : ((f < 0) ? -1 : 0));
  T r = f - std::trunc(f);
  if (r == 0) r = (T)0;
  else if (!(r > 0)) r = r + (T)1;
  xi = r;
}

// SeaIceAlbedo stateindex: Radiations/sea_ice_albedo.jl:106-133
template <class FT> FT sea_ice_albedo(const NeSeaIceAlbedo& a, int64_t idx) {
  FT hi = static_cast<const FT*>(a.ice_thickness)[idx];
  FT Ts = static_cast<const FT*>(a.surface_temperature)[idx];
  FT hs = a.snow_thickness ? static_cast<const FT*>(a.snow_thickness)[idx] : (FT)0;   // get_snow_thickness(::Nothing) :132
  FT Tm = (FT)a.melting_temperature, dT = (FT)a.temperature_range;
  FT fT = clampv((Ts - Tm + dT) / dT, (FT)0, (FT)1);                                   // :115
  FT ai = (FT)a.ice_albedo - (FT)a.ice_melt_reduction * fT;                            // :117
  FT as = (FT)a.snow_albedo - (FT)a.snow_melt_reduction * fT;                          // :118
  FT ao = (FT)a.ocean_albedo;
  FT fh = clampv(hi / (FT)a.minimum_ice_thickness, (FT)0, (FT)1);                      // :122
  ai = ao + (ai - ao) * fh;                                                            // :123
  FT fs = clampv(hs / (FT)a.minimum_snow_depth, (FT)0, (FT)1);                         // :126
  return fs * as + (1 - fs) * ai;                                                      // :127
}

// TabulatedAlbedo stateindex: Radiations/tabulated_albedo.jl:109-160.  day / seconds-in-day / declination
// (:113-131) are clock-only scalars evaluated by the caller.
template <class FT> FT tabulated_albedo(const NeTabulatedAlbedo& a, FT lam_deg, FT phi_deg, FT sw) {
  const FT deg = (FT)M_PI / 180;
  FT phi = phi_deg * deg, lam = lam_deg * deg;                                          // :113-114
  double h = (a.seconds_in_day - a.noon_in_seconds) * (FT)a.day_to_radians + lam;      // :122
  FT delta = (FT)a.declination;                                                        // :127
  auto cosz_raw = std::sin(phi) * std::sin(delta) + std::cos(h) * std::cos(delta) * std::cos(phi);   // :130
  using W = decltype(cosz_raw);
  W cosz = cosz_raw > 0 ? cosz_raw : (W)0;
  W Qmax = (FT)a.solar_constant * cosz;                                                // :134
  W tr = 0;
  if (Qmax > 0) { tr = sw / Qmax; if (tr > 1) tr = 1; }                                // :137
  FT t1 = (FT)a.t_values[0], dt = (FT)a.t_values[1] - t1;                              // :141-142
  FT p1 = (FT)a.phi_values[0], dp = (FT)a.phi_values[1] - p1;                          // :147-148
  W fi = (tr - t1) / dt;
  FT fj = (std::abs(phi) - p1) / dp;
  int64_t im, ip, jm, jp;
  W xi; FT eta;
  interpolator1<W>(fi, im, ip, xi);                                                    // :143
  interpolator1<FT>(fj, jm, jp, eta);                                                  // :149
  const FT* T = static_cast<const FT*>(a.table);
  const int64_t nt = a.n_t;
  // i⁺/j⁺ one past the table edge carry weight 0 in the reference (@inbounds read): keep the read inside the table
  auto at = [&](int64_t i, int64_t j) {
    i = i > a.n_t ? a.n_t : (i < 1 ? 1 : i);
    j = j > a.n_phi ? a.n_phi : (j < 1 ? 1 : j);
    return T[(i - 1) + (j - 1) * nt];
  };
  return (FT)((1 - xi) * (1 - eta) * at(im, jm) + (1 - xi) * eta * at(im, jp) +        // :152-155
              xi * (1 - eta) * at(ip, jm) + xi * eta * at(ip, jp));
}

// radiation state of one surface: Radiations/air_sea_interface_radiation_state.jl:4-39
template <class FT> RadState<FT> radiation_state(const NeSurfaceRadiation& r, const Layout& L, int64_t i, int64_t j) {
  RadState<FT> s = {0, 0, 0, 0, 0};
  if (!r.enabled) return s;
  int64_t idx = L.at(i, j);
  s.sigma = (FT)r.stefan_boltzmann_constant;
  s.sw = static_cast<const FT*>(r.downwelling_shortwave)[idx];
  s.lw = static_cast<const FT*>(r.downwelling_longwave)[idx];
  if (r.albedo_kind == NE_ALBEDO_CONSTANT) s.alpha = (FT)r.albedo;
  else if (r.albedo_kind == NE_ALBEDO_FIELD) s.alpha = static_cast<const FT*>(r.albedo_field)[idx];
  else if (r.albedo_kind == NE_ALBEDO_SEA_ICE) s.alpha = sea_ice_albedo<FT>(r.sea_ice_albedo, idx);
  else {
    FT phi = r.nodes_2d ? static_cast<const FT*>(r.latitude)[idx] : static_cast<const FT*>(r.latitude)[j + L.hy - 1];
    if (r.albedo_kind == NE_ALBEDO_TABULATED) {
      FT lam = r.nodes_2d ? static_cast<const FT*>(r.tabulated_albedo.longitude)[idx]
                          : static_cast<const FT*>(r.tabulated_albedo.longitude)[i + L.hx - 1];
      s.alpha = tabulated_albedo<FT>(r.tabulated_albedo, lam, phi, s.sw);
    } else {  // latitude_dependent_albedo.jl:48-53 ; hack_cosd(φ) = cos(π φ / 180) radiation_kernels.jl:1
      FT x = 2 * phi;
      s.alpha = (FT)r.albedo - (FT)r.albedo_direct * std::cos((FT)M_PI * x / 180);
    }
  }
  s.eps = (FT)r.emissivity;
  return s;
}

// ---- _compute_atmosphere_ocean_interface_state!: atmosphere_ocean_fluxes.jl:80-197 ---------------
template <class FT, class CT, class VT>
void atmosphere_ocean(const NeAtmosOceanDesc& d) {
  const Layout L(d.grid);
  const Thermo<CT> th(d.thermo);
  const FT g = (FT)d.gravitational_acceleration;
  const FT* ua = (const FT*)d.ua; const FT* va = (const FT*)d.va; const FT* Ta = (const FT*)d.Ta;
  const FT* pa = (const FT*)d.pa; const FT* qa = (const FT*)d.qa;
  const bool celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  const bool needs_to_converge = d.flux.stop.kind == NE_STOP_CONVERGENCE;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t j = d.grid.j_lo; j <= d.grid.j_hi; ++j) {
    for (int64_t i = d.grid.i_lo; i <= d.grid.i_hi; ++i) {
      const int64_t idx = L.at(i, j);
      AtmosState<FT> a;
      a.u = ua[idx]; a.v = va[idx]; a.T = Ta[idx]; a.p = pa[idx]; a.q = qa[idx];
      a.z = slot_at<FT>(d.surface_layer_height, idx);
      a.h_bl = slot_at<FT>(d.boundary_layer_height, idx);
      Interior<FT> in = {};
      // ℑxᶜᵃᵃ u, ℑyᵃᶜᵃ v [3rd-party: Oceananigans operators] :62-71
      in.u = d.uo.ptr ? (slot_at<FT>(d.uo, idx) + slot_at<FT>(d.uo, idx + 1)) / 2 : (FT)d.uo.value;
      in.v = d.vo.ptr ? (slot_at<FT>(d.vo, idx) + slot_at<FT>(d.vo, idx + L.sx)) / 2 : (FT)d.vo.value;
      FT To = slot_at<FT>(d.To, idx);
      if (celsius) To = To + (FT)273.15;
      in.T = To;
      in.S = slot_at<FT>(d.So, idx);
      if (d.properties.temperature_formulation == NE_TEMP_SKIN_DIFFUSIVE_INTERIOR) in.kappa = ((const FT*)d.kappa)[idx];
      RadState<FT> rad = radiation_state<FT>(d.radiation, L, i, j);

      FT us0 = (FT)1e-4;   // convert(FT, 1e-4) :132
      FT qs0 = surface_specific_humidity<FT, CT>(d.properties, th, a.p, in.T, in.S);  // :136
      State<FT> init = {us0, us0, us0, in.u, in.v, in.T, qs0, in.S};
      const bool not_water = d.inactive ? d.inactive[idx] != 0 : false;
      State<FT> st;
      int iters = 0;
      const State<FT> zero_state = {0, 0, 0, 0, 0, (FT)273.15, 0, 0};  // interface_states.jl:800-803
      if (needs_to_converge && not_water) st = zero_state;
      else st = compute_interface_state<FT, CT, VT>(d.flux, d.properties, th, g, init, a, in, rad, d.ocean, false, iters);
      if (not_water) st = zero_state;   // :158

      FT ustar = st.ustar, theta_star = st.theta_star, q_star = st.q_star;
      FT du, dv;
      if (d.properties.velocity_formulation == NE_VEL_RELATIVE) { du = a.u - st.u; dv = a.v - st.v; } else { du = a.u; dv = a.v; }
      FT dU = m_sqrt(sq(du) + sq(dv));
      FT taux = (dU == 0) ? (FT)0 : -sq(ustar) * du / dU;
      FT tauy = (dU == 0) ? (FT)0 : -sq(ustar) * dv / dU;
      auto rho_a = th.air_density(a.T, a.p, a.q);
      auto cpm = th.cp_m(a.q);
      auto Ll = th.latent_heat_vapor(a.T);
      ((FT*)d.latent_heat)[idx] = (FT)(-rho_a * Ll * ustar * q_star);
      ((FT*)d.sensible_heat)[idx] = (FT)(-rho_a * cpm * ustar * theta_star);
      ((FT*)d.water_vapor)[idx] = (FT)(-rho_a * ustar * q_star);
      ((FT*)d.x_momentum)[idx] = (FT)(rho_a * taux);
      ((FT*)d.y_momentum)[idx] = (FT)(rho_a * tauy);
      ((FT*)d.interface_temperature)[idx] = celsius ? st.T - (FT)273.15 : st.T;
      ((FT*)d.friction_velocity)[idx] = ustar;
      ((FT*)d.temperature_scale)[idx] = theta_star;
      ((FT*)d.water_vapor_scale)[idx] = q_star;
      if (d.iterations) d.iterations[idx] = iters;
    }
  }
}

// ---- _compute_atmosphere_sea_ice_interface_state!: atmosphere_sea_ice_fluxes.jl:65-185 -------------
template <class FT, class CT, class VT>
void atmosphere_sea_ice(const NeAtmosSeaIceDesc& d) {
  const Layout L(d.grid);
  const Thermo<CT> th(d.thermo);
  const FT g = (FT)d.gravitational_acceleration;
  const FT* ua = (const FT*)d.ua; const FT* va = (const FT*)d.va; const FT* Ta = (const FT*)d.Ta;
  const FT* pa = (const FT*)d.pa; const FT* qa = (const FT*)d.qa;
  const bool ocean_celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  const bool ice_celsius = d.sea_ice.temperature_units == NE_DEGREES_CELSIUS;
  const bool needs_to_converge = d.flux.stop.kind == NE_STOP_CONVERGENCE;
  FT* Tsurf = (FT*)d.interface_temperature;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t j = d.grid.j_lo; j <= d.grid.j_hi; ++j) {
    for (int64_t i = d.grid.i_lo; i <= d.grid.i_hi; ++i) {
      const int64_t idx = L.at(i, j);
      AtmosState<FT> a;
      a.u = ua[idx]; a.v = va[idx]; a.T = Ta[idx]; a.p = pa[idx]; a.q = qa[idx];
      a.z = slot_at<FT>(d.surface_layer_height, idx);
      a.h_bl = slot_at<FT>(d.boundary_layer_height, idx);
      FT To = slot_at<FT>(d.To, idx);
      if (ocean_celsius) To = To + (FT)273.15;
      FT So = slot_at<FT>(d.So, idx);
      Interior<FT> in = {};
      in.u = 0; in.v = 0; in.T = To; in.S = So;
      in.hi = slot_at<FT>(d.hi, idx); in.hs = slot_at<FT>(d.hs, idx); in.hc = slot_at<FT>(d.hc, idx);
      FT conc = slot_at<FT>(d.concentration, idx);
      FT Ts = Tsurf[idx];
      if (ice_celsius) Ts = Ts + (FT)273.15;
      RadState<FT> rad = radiation_state<FT>(d.radiation, L, i, j);
      FT us0 = (FT)1e-4f;   // convert(FT, 1f-4) :127
      FT qs0 = surface_specific_humidity<FT, CT>(d.properties, th, a.p, Ts, So);  // :131
      State<FT> init = {us0, us0, us0, (FT)0, (FT)0, Ts, qs0, (FT)0};
      const bool not_water = d.inactive ? d.inactive[idx] != 0 : false;
      const bool ice_free = conc == 0;
      State<FT> st;
      int iters = 0;
      if ((needs_to_converge && not_water) || ice_free) st = State<FT>{0, 0, 0, 0, 0, To, 0, 0};  // :141-142
      else st = compute_interface_state<FT, CT, VT>(d.flux, d.properties, th, g, init, a, in, rad, d.sea_ice, true, iters);
      FT ustar = st.ustar, theta_star = st.theta_star, q_star = st.q_star;
      FT du, dv;
      if (d.properties.velocity_formulation == NE_VEL_RELATIVE) { du = a.u - st.u; dv = a.v - st.v; } else { du = a.u; dv = a.v; }
      FT dU = m_sqrt(sq(du) + sq(dv));
      FT taux = (dU == 0) ? (FT)0 : -sq(ustar) * du / dU;
      FT tauy = (dU == 0) ? (FT)0 : -sq(ustar) * dv / dU;
      auto rho_a = th.air_density(a.T, a.p, a.q);
      auto cpm = th.cp_m(a.q);
      auto Li = th.latent_heat_sublim(a.T);
      ((FT*)d.latent_heat)[idx] = (FT)(-rho_a * ustar * q_star * Li);   // :178 (note the different product order)
      ((FT*)d.sensible_heat)[idx] = (FT)(-rho_a * cpm * ustar * theta_star);
      ((FT*)d.water_vapor)[idx] = (FT)(-rho_a * ustar * q_star);
      ((FT*)d.x_momentum)[idx] = (FT)(rho_a * taux);
      ((FT*)d.y_momentum)[idx] = (FT)(rho_a * tauy);
      Tsurf[idx] = ice_celsius ? st.T - (FT)273.15 : st.T;
      if (d.iterations) d.iterations[idx] = iters;
    }
  }
}

// ---- interpolation: Atmospheres/interpolate_atmospheric_state.jl:91-182 + [3rd-party Oceananigans
// interpolator/_interpolate/time interpolation] ----------------------------------------------------
// local_roughness_length(::LandRoughnessLength, interior_properties, Val(R)) (similarity_theory_turbulent_fluxes.jl:265-278):
// the land model's per-cell field, floored, scaled and floored again; the floor alone when the land model has no such field
template <class FT>
static void resolve_land_marker(NeRoughnessLength& r, const FT* field, int64_t idx) {
  if (r.kind != NE_ROUGH_LAND) return;
  const FT lmin = (FT)r.land_minimum_roughness_length, mult = (FT)r.land_multiplier;
  const FT candidate = field ? mx(field[idx], lmin) : lmin;
  r.kind = NE_ROUGH_CONSTANT;
  r.constant = (double)mx(mult * candidate, lmin);
}

// ---- _compute_atmosphere_land_interface_state!: atmosphere_land_fluxes.jl:147-251 ------------------------
template <class FT, class CT, class VT>
void atmosphere_land(const NeAtmosLandDesc& d) {
  const Layout L(d.grid);
  const Thermo<CT> th(d.thermo);
  const FT g = (FT)d.gravitational_acceleration;
  const FT* ua = (const FT*)d.ua; const FT* va = (const FT*)d.va; const FT* Ta = (const FT*)d.Ta;
  const FT* pa = (const FT*)d.pa; const FT* qa = (const FT*)d.qa;
  NeMediumProperties medium = {};   // only read by skin temperatures, which the land interface does not use
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t j = d.grid.j_lo; j <= d.grid.j_hi; ++j) {
    for (int64_t i = d.grid.i_lo; i <= d.grid.i_hi; ++i) {
      const int64_t idx = L.at(i, j);
      AtmosState<FT> a;
      a.u = ua[idx]; a.v = va[idx]; a.T = Ta[idx]; a.p = pa[idx]; a.q = qa[idx];
      a.z = slot_at<FT>(d.surface_layer_height, idx);
      a.h_bl = slot_at<FT>(d.boundary_layer_height, idx);
      const FT Ts = slot_at<FT>(d.land_temperature, idx);   // bulk land temperature = initial (and Bulk) interface temperature
      Interior<FT> in = {};
      in.u = 0; in.v = 0; in.T = Ts;                         // surface velocities are zero for land (:189-191)
      RadState<FT> rad = {};
      LandSurface<FT> land = {&d.humidity, slot_at<FT>(d.saturation, idx), Ts};
      const FT us0 = (FT)1e-4;                                // convert(FT, 1e-4) :204
      const FT qs0 = (FT)saturation_specific_humidity<CT>(th, Ts, a.p, d.humidity.phase);   // :205
      State<FT> init = {us0, us0, us0, (FT)0, (FT)0, Ts, qs0, land.saturation};
      int iters = 0;
      // local_roughness_lengths / local_zero_plane_displacement (similarity_theory_turbulent_fluxes.jl:265-303): the land
      // markers resolve to Numbers from this cell's land properties before the iteration sees them
      NeFluxFormulation flux = d.flux;
      resolve_land_marker<FT>(flux.ell_momentum, (const FT*)d.momentum_roughness_length, idx);
      resolve_land_marker<FT>(flux.ell_temperature, (const FT*)d.scalar_roughness_length, idx);
      resolve_land_marker<FT>(flux.ell_water_vapor, (const FT*)d.scalar_roughness_length, idx);
      if (flux.zero_plane_displacement_kind == NE_DISPLACEMENT_LAND) {
        flux.zero_plane_displacement = d.zero_plane_displacement ? (double)((const FT*)d.zero_plane_displacement)[idx] : 0.0;
        flux.zero_plane_displacement_kind = NE_DISPLACEMENT_CONSTANT;
      }
      State<FT> st = compute_interface_state<FT, CT, VT>(flux, d.properties, th, g, init, a, in, rad, medium, false, iters, &land);
      FT ustar = st.ustar, theta_star = st.theta_star, q_star = st.q_star;
      FT du, dv;
      if (d.properties.velocity_formulation == NE_VEL_RELATIVE) { du = a.u - st.u; dv = a.v - st.v; } else { du = a.u; dv = a.v; }
      FT dU = m_sqrt(sq(du) + sq(dv));
      FT taux = (dU == 0) ? (FT)0 : -sq(ustar) * du / dU;
      FT tauy = (dU == 0) ? (FT)0 : -sq(ustar) * dv / dU;
      auto rho_a = th.air_density(a.T, a.p, a.q);
      auto cpm = th.cp_m(a.q);
      auto Ll = th.latent_heat_vapor(a.T);
      ((FT*)d.latent_heat)[idx] = (FT)(-rho_a * Ll * ustar * q_star);
      ((FT*)d.sensible_heat)[idx] = (FT)(-rho_a * cpm * ustar * theta_star);
      ((FT*)d.water_vapor)[idx] = (FT)(-rho_a * ustar * q_star);
      ((FT*)d.x_momentum)[idx] = (FT)(rho_a * taux);
      ((FT*)d.y_momentum)[idx] = (FT)(rho_a * tauy);
      ((FT*)d.interface_temperature)[idx] = st.T;
      ((FT*)d.friction_velocity)[idx] = ustar;
      ((FT*)d.temperature_scale)[idx] = theta_star;
      ((FT*)d.water_vapor_scale)[idx] = q_star;
      if (d.iterations) d.iterations[idx] = iters;
    }
  }
}

template <class AT> struct Interpolator { int64_t im, ip; AT xi; };

template <class AT> inline AT julia_mod1(AT x) {   // Base.mod(x, one(x)) for floats
  AT r = std::fmod(x, (AT)1);
  if (r == 0) return std::copysign(r, (AT)1);
  if ((r > 0) != true) return r + (AT)1;   // (r > 0) ⊻ (y > 0) with y = 1
  return r;
}
template <class AT> inline Interpolator<AT> interpolator(AT f) {
  Interpolator<AT> it;
  it.im = (int64_t)f + 1;                                // Base.unsafe_trunc(Int, f) + 1
  int64_t sgn = (f > 0) ? 1 : ((f < 0) ? -1 : 0);        // Int(sign(f))
  it.ip = it.im + sgn;
  it.xi = julia_mod1(f);
  return it;
}

template <class FT, class AT, class TT>
void interp_state(const NeInterpDesc& d) {
  const Layout L(d.grid);
  const int64_t ssx = d.src_nx + 2 * d.src_hx, ssy = d.src_ny + 2 * d.src_hy;
  const int64_t plane = ssx * ssy;
  const AT* fi = (const AT*)d.frac_i;
  const AT* fj = (const AT*)d.frac_j;
  const TT nt = (TT)d.time.frac;
  const int64_t o1 = (int64_t)(d.time.m1 - 1) * plane, o2 = (int64_t)(d.time.m2 - 1) * plane;
  using W = decltype(AT() * TT());
  const bool rotate = d.rotation_cos && d.rotation_sin && d.rotate_u >= 0 && d.rotate_v >= 0 && d.rotate_u < d.n_fields &&
                      d.rotate_v < d.n_fields && d.rotate_u != d.rotate_v && d.out[d.rotate_u] && d.out[d.rotate_v];
#pragma omp parallel for schedule(static)
  for (int64_t j = d.grid.j_lo; j <= d.grid.j_hi; ++j) {
    for (int64_t i = d.grid.i_lo; i <= d.grid.i_hi; ++i) {
      const int64_t idx = L.at(i, j);
      Interpolator<AT> ix = {1, 1, 0}, iy = {1, 1, 0};   // interpolator(nothing) = (1, 1, 0)
      if (fi) ix = interpolator<AT>(fi[idx]);
      if (fj) iy = interpolator<AT>(fj[idx]);
      const AT xi = ix.xi, eta = iy.xi;
      // ϕ₁..ϕ₈ with ζ = 0 (the Int 0): the k⁺ terms are exact zeros and drop out of the left-to-right sum
      const AT w1 = (1 - xi) * (1 - eta), w3 = (1 - xi) * eta, w5 = xi * (1 - eta), w7 = xi * eta;
      const int64_t a_mm = (ix.im + d.src_hx - 1) + (iy.im + d.src_hy - 1) * ssx;
      const int64_t a_mp = (ix.im + d.src_hx - 1) + (iy.ip + d.src_hy - 1) * ssx;
      const int64_t a_pm = (ix.ip + d.src_hx - 1) + (iy.im + d.src_hy - 1) * ssx;
      const int64_t a_pp = (ix.ip + d.src_hx - 1) + (iy.ip + d.src_hy - 1) * ssx;
      W vec_u = 0, vec_v = 0;
      for (int f = 0; f < d.n_fields; ++f) {
        FT* out = (FT*)d.out[f];
        if (!out) continue;
        W total = 0;
        bool first = true;
        for (int s = 0; s < d.n_summands[f]; ++s) {
          const AT* data = (const AT*)d.series[f][s].data;
          W val;
          if (!data) val = 0;   // interp_atmos_time_series(::Nothing, ...) = 0
          else {
            const AT* d1 = data + o1;
            const AT* d2 = data + o2;
            AT p1 = w1 * d1[a_mm] + w3 * d1[a_mp] + w5 * d1[a_pm] + w7 * d1[a_pp];
            AT p2 = w1 * d2[a_mm] + w3 * d2[a_mp] + w5 * d2[a_pm] + w7 * d2[a_pp];
            W pt = p2 * nt + p1 * (1 - nt);
            val = d.time.same ? (W)p1 : pt;
          }
          total = first ? val : total + val;
          first = false;
        }
        if (rotate && (f == d.rotate_u || f == d.rotate_v)) { (f == d.rotate_u ? vec_u : vec_v) = total; continue; }
        out[idx] = (FT)total;
      }
      if (rotate) {   // intrinsic_vector (interpolate_atmospheric_state.jl:123-126; Oceananigans rotation, third party)
        using P = decltype(W() * FT());
        const P c = ((const FT*)d.rotation_cos)[idx], sn = ((const FT*)d.rotation_sin)[idx];
        const P ur = (P)vec_u * c + (P)vec_v * sn;
        const P vr = -(P)vec_u * sn + (P)vec_v * c;
        ((FT*)d.out[d.rotate_u])[idx] = (FT)ur;
        ((FT*)d.out[d.rotate_v])[idx] = (FT)vr;
      }
      if (d.potential) ((FT*)d.potential)[idx] = ((FT*)d.out[d.potential_from])[idx] / (FT)d.ocean_reference_density;
    }
  }
}

// ---- fractional indices [3rd-party Oceananigans FractionalIndices on LatitudeLongitudeGrid] -------
template <class AT> inline AT convert_to_lam0_lam0_plus360(AT x, AT lam0) {
  return std::fmod(std::fmod(x - lam0, (AT)360) + (AT)360, (AT)360) + lam0;
}
template <class AT> inline AT fractional_index_search(AT x, const AT* xs, int64_t N) {   // 1-based result
  int64_t low = 0, high = N - 1;
  while (low + 1 < high) {
    int64_t mid = (low + high) >> 1;
    if (xs[mid] == x) return (AT)(mid + 1);
    else if (xs[mid] < x) low = mid;
    else high = mid;
  }
  int64_t i1, i2;
  if (xs[high] == x) { i1 = i2 = high + 1; }
  else if (xs[low] == x) { i1 = i2 = low + 1; }
  else { i1 = low + 1; i2 = high + 1; }
  if (i1 == i2) return (AT)i1;
  AT x1 = xs[i1 - 1], x2 = xs[i2 - 1];
  return (AT)(i2 - i1) / (x2 - x1) * (x - x1) + (AT)i1;
}

template <class FT, class AT>
void frac_indices(const NeFracIndexDesc& d) {
  const Layout L(d.grid);
  const AT* lam_n = (const AT*)d.src_lam_nodes;
  const AT* phi_n = (const AT*)d.src_phi_nodes;
  const AT lam0 = lam_n[0], dlam = lam_n[1] - lam_n[0];
  const AT phi0 = phi_n[0], dphi = phi_n[1] - phi_n[0];
#pragma omp parallel for schedule(static)
  for (int64_t j = d.grid.j_lo; j <= d.grid.j_hi; ++j) {
    for (int64_t i = d.grid.i_lo; i <= d.grid.i_hi; ++i) {
      const int64_t idx = L.at(i, j);
      FT lam = d.nodes_2d ? ((const FT*)d.lam)[idx] : ((const FT*)d.lam)[i + L.hx - 1];
      FT phi = d.nodes_2d ? ((const FT*)d.phi)[idx] : ((const FT*)d.phi)[j + L.hy - 1];
      using W = decltype(FT() + AT());
      W lc = convert_to_lam0_lam0_plus360<W>((W)lam, (W)(lam0 - dlam / 2));
      AT fi, fj;
      if (d.src_x_regular) fi = (AT)((lc - lam0) / dlam);
      else fi = fractional_index_search<AT>((AT)lc, lam_n, d.src_nx) - 1;   // stretched axis
      if (d.src_y_regular) fj = (AT)(((W)phi - phi0) / dphi);
      else fj = fractional_index_search<AT>((AT)phi, phi_n, d.src_ny) - 1;
      if (d.frac_i) ((AT*)d.frac_i)[idx] = fi;
      if (d.frac_j) ((AT*)d.frac_j)[idx] = fj;
    }
  }
}

// ---- sea-ice–ocean: sea_ice_ocean_fluxes.jl:106-226, sea_ice_ocean_heat_flux_formulations.jl -------
template <class FT>
void sea_ice_ocean(const NeSeaIceOceanDesc& d) {
  const Layout L(d.grid);
  const int64_t plane = L.sx * (d.grid.ny + 2 * d.grid.hy);
  FT* T = (FT*)d.T; const FT* S = (const FT*)d.S; const FT* dz = (const FT*)d.dz;
  const FT rho = (FT)d.ocean.reference_density, c = (FT)d.ocean.heat_capacity;
  const FT slope = (FT)d.ocean.liquidus_slope, Tfresh = (FT)d.ocean.liquidus_freshwater_melting_temperature;
  const FT dt = (FT)d.dt;
  const FT E = (FT)d.latent_heat;
#pragma omp parallel for schedule(static)
  for (int64_t j = d.grid.j_lo; j <= d.grid.j_hi; ++j) {
    for (int64_t i = d.grid.i_lo; i <= d.grid.i_hi; ++i) {
      const int64_t idx = L.at(i, j);
      FT dQ = 0;
      for (int64_t k = d.nz; k >= 1; --k) {   // :155-178
        const int64_t a = idx + (k + d.hz - 1) * plane;
        FT Tk = T[a], Sk = S[a];
        FT Tm = Tfresh - slope * Sk;          // [3rd-party] melting_temperature(LinearLiquidus, S)
        bool freezing = Tk < Tm;
        FT dE = (FT)freezing * rho * c * (Tm - Tk);
        T[a] = freezing ? Tm : Tk;
        dQ -= dE * dz[k - 1] / dt;
      }
      ((FT*)d.frazil_heat)[idx] = dQ;
      if (d.formulation == NE_SIO_FREEZE_ONLY) continue;   // freezing_limited_ocean_temperature.jl:95-118

      const int64_t top = idx + (d.nz + d.hz - 1) * plane;
      FT TN = T[top], SN = S[top];
      FT Si = slot_at<FT>(d.ice_salinity, idx);
      FT hi = slot_at<FT>(d.hi, idx), conc = slot_at<FT>(d.concentration, idx), hc = slot_at<FT>(d.hc, idx);
      FT Tint = d.has_conductive_flux ? ((const FT*)d.internal_temperature)[idx] : (FT)0;
      FT ustar;
      if (d.friction_velocity_kind == NE_USTAR_CONSTANT) ustar = (FT)d.friction_velocity;
      else {  // friction_velocity.jl:26-27,44 : sqrt(sqrt(ℑx τx² + ℑy τy²)/ρ)
        const FT* tx = (const FT*)d.x_momentum_in; const FT* ty = (const FT*)d.y_momentum_in;
        FT ax = (sq(tx[idx]) + sq(tx[idx + 1])) / 2;
        FT ay = (sq(ty[idx]) + sq(ty[idx + L.sx])) / 2;
        ustar = m_sqrt(m_sqrt(ax + ay) / rho);
      }
      FT Q, Tb, Sb;
      if (d.formulation == NE_SIO_ICE_BATH) {   // heat_flux_formulations.jl:176-195
        FT Tm = Tfresh - slope * SN;
        Q = rho * c * (FT)d.heat_transfer_coefficient * ustar * (TN - Tm) * conc;
        Tb = Tm; Sb = SN;
      } else {   // three-equation :224-313
        FT ah = (FT)d.heat_transfer_coefficient, as = (FT)d.salt_transfer_coefficient;
        FT kap = 0, Tsi = 0;
        if (d.has_conductive_flux) {   // :251-260
          bool consolidated = hi >= hc;
          kap = consolidated ? (FT)d.conductivity / (hi * E) : (FT)0;
          Tsi = Tint;
        }
        FT l1 = -slope, l2 = Tfresh;
        FT eta = rho * c * ah * ustar / E;
        FT gam = rho * as * ustar;
        FT th = eta + kap;
        FT qa = th * l1;
        FT qb = -gam - eta * TN - kap * Tsi + th * (l2 - l1 * Si);
        FT qc = gam * SN + (eta * TN + kap * Tsi - th * l2) * Si;
        FT xi = (qa == 0) ? (FT)0 : 1 / (2 * qa);
        FT Dl = m_max<FT>(sq(qb) - 4 * qa * qc, 0);
        FT Ss = (-qb - m_sqrt(Dl)) * xi;
        Ss = (Ss < 0) ? (-qb + m_sqrt(Dl)) * xi : Ss;
        FT Ts = Tfresh - slope * Ss;
        FT q = eta * (TN - Ts) + kap * (Tsi - Ts);
        Q = E * q * conc;
        Tb = Ts; Sb = Ss;
        ((FT*)d.interface_temperature)[idx] = Tb;
        ((FT*)d.interface_salinity)[idx] = Sb;
      }
      ((FT*)d.interface_heat)[idx] = Q;
      FT Ei = slot_at<FT>(d.ice_mass_flux, idx), Es = slot_at<FT>(d.snow_mass_flux, idx);
      ((FT*)d.freshwater)[idx] = -(Ei + Es) / rho;
      ((FT*)d.salt)[idx] = Ei * Si / rho;
    }
  }
}

// _compute_sea_ice_ocean_stress! with [3rd-party ClimaSeaIce SemiImplicitStress]:
// tau_x at (Face,Center): rho Cd |Δu| (uo - ui) with v averaged to the u point, and vice versa.
template <class FT>
void sea_ice_ocean_stress(const NeSeaIceOceanStressDesc& d) {
  const Layout L(d.grid);
  const FT* ui = (const FT*)d.ui; const FT* vi = (const FT*)d.vi;
  const FT* uo = (const FT*)d.uo; const FT* vo = (const FT*)d.vo;
  const FT rho = (FT)d.ocean_density, Cd = (FT)d.drag_coefficient;
#pragma omp parallel for schedule(static)
  for (int64_t j = d.grid.j_lo; j <= d.grid.j_hi; ++j)
    for (int64_t i = d.grid.i_lo; i <= d.grid.i_hi; ++i) {
      const int64_t idx = L.at(i, j);
      // x-stress at (f, c): Δu local, Δv = ℑxyᶠᶜᵃ(vi - vo)
      FT du = uo[idx] - ui[idx];   // Δu = uₑ − uᵢ (sign pinned by test/test_surface_fluxes.jl:334-335)
      FT dv4 = ((vo[idx] - vi[idx]) + (vo[idx - 1] - vi[idx - 1]) + (vo[idx + L.sx] - vi[idx + L.sx]) +
                (vo[idx - 1 + L.sx] - vi[idx - 1 + L.sx])) / 4;
      ((FT*)d.x_momentum)[idx] = rho * Cd * m_sqrt(sq(du) + sq(dv4)) * du;
      FT dv = vo[idx] - vi[idx];
      FT du4 = ((uo[idx] - ui[idx]) + (uo[idx + 1] - ui[idx + 1]) + (uo[idx - L.sx] - ui[idx - L.sx]) +
                (uo[idx + 1 - L.sx] - ui[idx + 1 - L.sx])) / 4;
      ((FT*)d.y_momentum)[idx] = rho * Cd * m_sqrt(sq(du4) + sq(dv)) * dv;
    }
}

// ---- net flux assembly: Oceans/assemble_net_ocean_fluxes.jl:74-153 ----------------------------------
template <class FT>
void assemble_ocean(const NeAssembleOceanDesc& d) {
  const Layout L(d.grid);
  const FT rho_inv = 1 / (FT)d.ocean.reference_density;
  const FT c_inv = 1 / (FT)d.ocean.heat_capacity;
#pragma omp parallel for schedule(static)
  for (int64_t j = d.grid.j_lo; j <= d.grid.j_hi; ++j)
    for (int64_t i = d.grid.i_lo; i <= d.grid.i_hi; ++i) {
      const int64_t idx = L.at(i, j);
      FT conc = slot_at<FT>(d.concentration, idx);
      FT To = slot_at<FT>(d.ocean_surface_temperature, idx);
      FT Jrn = slot_at<FT>(d.rainfall, idx), Jsn = slot_at<FT>(d.snowfall, idx);
      FT Psn = slot_at<FT>(d.intercepted_snowfall, idx), Jln = slot_at<FT>(d.land_freshwater, idx);
      FT QT = slot_at<FT>(d.sensible_heat, idx), Qv = slot_at<FT>(d.latent_heat, idx), Jv = slot_at<FT>(d.water_vapor, idx);
      FT SQ = (QT + Qv) * (1 - conc);
      FT SF = -(Jrn + Jln + Jsn - Psn) * rho_inv + (1 - conc) * Jv * rho_inv;
      FT Jw_ao = -SF;
      bool inactive = d.inactive ? d.inactive[idx] != 0 : false;
      FT Qin = slot_at<FT>(d.interface_heat, idx), Js_io = slot_at<FT>(d.salt_io, idx), Jw_io = slot_at<FT>(d.freshwater_io, idx);
      FT JT_ao = SQ * rho_inv * c_inv;
      FT JT_io = Qin * rho_inv * c_inv;
      auto tau_ccc = [&](const NeSlot& t, int64_t a) { return rho_inv * (1 - slot_at<FT>(d.concentration, a)) * slot_at<FT>(t, a); };
      // ℑxᶠᵃᵃ f = (f[i-1] + f[i]) / 2 ; ℑyᵃᶠᵃ f = (f[j-1] + f[j]) / 2  [3rd-party operators]
      FT tx_ao = (tau_ccc(d.x_momentum_ao, idx - 1) + tau_ccc(d.x_momentum_ao, idx)) / 2;
      FT ty_ao = (tau_ccc(d.y_momentum_ao, idx - L.sx) + tau_ccc(d.y_momentum_ao, idx)) / 2;
      FT cx = (slot_at<FT>(d.concentration, idx - 1) + slot_at<FT>(d.concentration, idx)) / 2;
      FT cy = (slot_at<FT>(d.concentration, idx - L.sx) + slot_at<FT>(d.concentration, idx)) / 2;
      FT tx_io = slot_at<FT>(d.x_momentum_io, idx) * rho_inv * cx;
      FT ty_io = slot_at<FT>(d.y_momentum_io, idx) * rho_inv * cy;
      ((FT*)d.tau_x)[idx] = inactive ? (FT)0 : tx_ao + tx_io;
      ((FT*)d.tau_y)[idx] = inactive ? (FT)0 : ty_ao + ty_io;
      ((FT*)d.JT)[idx] = inactive ? (FT)0 : JT_ao + JT_io;
      ((FT*)d.JS)[idx] = inactive ? (FT)0 : Js_io;
      ((FT*)d.Jw)[idx] = inactive ? (FT)0 : Jw_ao + Jw_io;
      ((FT*)d.JH)[idx] = inactive ? (FT)0 : To * Jw_ao;
    }
}

// SeaIces/assemble_net_sea_ice_fluxes.jl:42-81
template <class FT>
void assemble_sea_ice(const NeAssembleSeaIceDesc& d) {
  const Layout L(d.grid);
#pragma omp parallel for schedule(static)
  for (int64_t j = d.grid.j_lo; j <= d.grid.j_hi; ++j)
    for (int64_t i = d.grid.i_lo; i <= d.grid.i_hi; ++i) {
      const int64_t idx = L.at(i, j);
      FT conc = slot_at<FT>(d.concentration, idx);
      FT QT = slot_at<FT>(d.sensible_heat, idx), Qv = slot_at<FT>(d.latent_heat, idx);
      FT Qf = slot_at<FT>(d.frazil_heat, idx), Qi = slot_at<FT>(d.interface_heat, idx);
      FT Jsn = slot_at<FT>(d.snowfall, idx);
      FT SQt = (QT + Qv) * conc;
      FT SQb = Qf + Qi;
      bool inactive = d.inactive ? d.inactive[idx] != 0 : false;
      FT tu = (slot_at<FT>(d.x_momentum, idx - 1) + slot_at<FT>(d.x_momentum, idx)) / 2;
      FT tv = (slot_at<FT>(d.y_momentum, idx - L.sx) + slot_at<FT>(d.y_momentum, idx)) / 2;
      ((FT*)d.top_heat)[idx] = inactive ? (FT)0 : SQt;
      ((FT*)d.top_snowfall)[idx] = inactive ? (FT)0 : Jsn;
      ((FT*)d.top_u)[idx] = inactive ? (FT)0 : tu;
      ((FT*)d.top_v)[idx] = inactive ? (FT)0 : tv;
      ((FT*)d.bottom_heat)[idx] = inactive ? (FT)0 : SQb;
    }
}

// Radiations/apply_air_sea_radiative_fluxes.jl:62-111 and apply_air_sea_ice_radiative_fluxes.jl:55-90
template <class FT>
void apply_radiation(const NeApplyRadiationDesc& d) {
  const Layout L(d.grid);
  const bool celsius = d.medium.temperature_units == NE_DEGREES_CELSIUS;
#pragma omp parallel for schedule(static)
  for (int64_t j = d.grid.j_lo; j <= d.grid.j_hi; ++j)
    for (int64_t i = d.grid.i_lo; i <= d.grid.i_hi; ++i) {
      const int64_t idx = L.at(i, j);
      FT conc = slot_at<FT>(d.concentration, idx);
      FT Ts = ((const FT*)d.surface_temperature)[idx];
      if (celsius) Ts = Ts + (FT)273.15;
      RadState<FT> rs = radiation_state<FT>(d.radiation, L, i, j);
      FT up = rs.sigma * rs.eps * ipow(Ts, 4);      // radiation_kernels.jl:3
      FT ab = -rs.eps * rs.lw;                      // :4
      FT tr = -(1 - rs.alpha) * rs.sw;              // :5
      bool inactive = d.inactive ? d.inactive[idx] != 0 : false;
      FT* H = (FT*)d.heat_flux;
      if (!d.over_sea_ice) {
        ab *= (1 - conc);
        tr *= (1 - conc);
        FT up_o = up * (1 - conc);
        FT Qss = tr;
        if (d.two_color) {   // Oceans/radiative_forcing.jl:84-91
          FT rho = (FT)d.medium.reference_density, c = (FT)d.medium.heat_capacity;
          ((FT*)d.two_color_surface_flux)[idx] = -tr / (rho * c);
          Qss = 0;
        }
        FT SQ = up_o + ab + Qss;
        FT rho_inv = 1 / (FT)d.medium.reference_density, c_inv = 1 / (FT)d.medium.heat_capacity;
        FT JT = SQ * rho_inv * c_inv;
        H[idx] += inactive ? (FT)0 : JT;
      } else if (d.over_sea_ice == 2) {   // land: apply_air_land_radiative_fluxes.jl:79-93 (surface_energy_flux is positive upward)
        FT SQrad = -up - (ab + tr);
        H[idx] += inactive ? (FT)0 : -SQrad;
      } else {
        FT SQ = (up + ab + tr) * conc;
        H[idx] += inactive ? (FT)0 : SQ;
      }
      ((FT*)d.upwelling_longwave)[idx] = up;
      ((FT*)d.downwelling_longwave)[idx] = -ab;
      ((FT*)d.downwelling_shortwave)[idx] = -tr;
    }
}

// _correct_atmosphere_elevation! (EarthSystemModels/InterfaceComputations/atmosphere_state_correction.jl:133-146)
template <class FT>
void correct_atmosphere_elevation(const NeElevationCorrectionDesc& d) {
  const Layout L(d.grid);
#pragma omp parallel for schedule(static)
  for (int64_t j = d.grid.j_lo; j <= d.grid.j_hi; ++j)
    for (int64_t i = d.grid.i_lo; i <= d.grid.i_hi; ++i) {
      const int64_t idx = L.at(i, j);
      FT dz = ((const FT*)d.elevation_difference)[idx];          // δz = convert(FT, Δz[i, j, 1])
      FT dT = (FT)d.lapse_rate * dz;                               // ΔT = convert(FT, Γ) * δz
      FT* T = (FT*)d.T;
      FT* p = (FT*)d.p;
      FT T0 = T[idx];
      FT Tbar = T0 - dT / 2;                                       // :141
      p[idx] = p[idx] * std::exp(-(FT)d.gravitational_acceleration * dz / ((FT)d.dry_air_gas_constant * Tbar));   // :142
      T[idx] = T0 - dT;                                            // :143
    }
}

// viscosity element type: double when every constant-viscosity slot keeps the Float64 literal,
// FT when all use TemperatureDependentAirViscosity{FT}.
inline bool viscosity_is_f64_literal(const NeFluxFormulation& f) {
  const NeRoughnessLength* r[3] = {&f.ell_momentum, &f.ell_temperature, &f.ell_water_vapor};
  for (auto p : r)
    if (p->kind != NE_ROUGH_CONSTANT && p->visc_kind == NE_VISC_CONSTANT && p->visc_dtype == NE_F64) return true;
  return false;
}

}  // namespace

#define NEO_DISPATCH_FLUX(FUNC, FT, desc)                                                     \
  do {                                                                                        \
    const bool ct64 = (desc)->thermo.dtype == NE_F64;                                         \
    const bool v64 = std::is_same<FT, double>::value || viscosity_is_f64_literal((desc)->flux); \
    if (ct64) { if (v64) FUNC<FT, double, double>(*(desc)); else FUNC<FT, double, FT>(*(desc)); } \
    else      { if (v64) FUNC<FT, float, double>(*(desc));  else FUNC<FT, float, FT>(*(desc)); }  \
  } while (0)

// set!(fts) for ONE slice of ONE series (JRA55_field_time_series.jl:60-76): _set_region_kernel! over the interior
// (set_region_data.jl:200-205; whole-globe region, no mangling: missing -> NaN, then convert_units,
// metadata_field.jl:486-525), then fill_halo_regions!: periodic in x (test_jra55.jl:40-47) or zero-flux mirror,
// zero-flux mirror in y.  raw: nx*ny, x fastest; dst: (ny+2hy) rows of (nx+2hx).
template <typename T>
static void series_slot_fill(const NeSeriesRingDesc& d, int k, const T* raw, T* dst) {
  const int64_t W = d.nx + 2 * d.hx;
  const int64_t rnx = d.raw_nx ? d.raw_nx : d.nx, rny = d.raw_ny ? d.raw_ny : d.ny;
  const T a = (T)d.conv_a[k], b = (T)d.conv_b[k], mv = (T)d.missing_value[k];
  auto at = [&](int64_t i, int64_t j) -> T& { return dst[(j + d.hy) * W + (i + d.hx)]; };
  // mangle (set_region_data.jl:48-53): 1-based file indices clamped to the file extent; here 0-based
  auto file = [&](int64_t i, int64_t j) -> T {
    i = std::min(std::max<int64_t>(i, 0), rnx - 1);
    j = std::min(std::max<int64_t>(j, 0), rny - 1);
    return raw[j * rnx + i];
  };
  // mangle + nan_convert_missing of one file cell (:48-53, :162-163)
  auto read = [&](int64_t fi, int64_t fj) -> T {
    T v;
    switch (d.mangling[k]) {
      case NE_MANGLE_SHIFT_SOUTH: v = file(fi, fj - 1); break;
      case NE_MANGLE_AVERAGE_NORTH_SOUTH: { volatile T sum = file(fi, fj) + file(fi, fj + 1); v = sum / (T)2; break; }
      default: v = file(fi, fj); break;
    }
    if (d.has_missing[k] && v == mv) v = std::numeric_limits<T>::quiet_NaN();
    return v;
  };
  // blend(::Linear, …) (:168-187): NaN corners dropped, weights renormalised; blend(::Nearest, …) (:189-195)
  auto blend_linear = [&]() -> T {
    const T wx = (T)d.col_wx, wy = (T)d.col_wy;
    const T d00 = read(d.col_i_minus, d.col_j_minus), d10 = read(d.col_i_plus, d.col_j_minus);
    const T d01 = read(d.col_i_minus, d.col_j_plus), d11 = read(d.col_i_plus, d.col_j_plus);
    volatile T cx = (T)1 - wx, cy = (T)1 - wy;
    volatile T p00 = cx * cy, p10 = wx * cy, p01 = cx * wy, p11 = wx * wy;
    volatile T w00 = p00 * (T)!std::isnan(d00), w10 = p10 * (T)!std::isnan(d10);
    volatile T w01 = p01 * (T)!std::isnan(d01), w11 = p11 * (T)!std::isnan(d11);
    volatile T s1 = w00 + w10; volatile T s2 = s1 + w01; volatile T sw = s2 + w11;
    volatile T t00 = w00 * (std::isnan(d00) ? (T)0 : d00), t10 = w10 * (std::isnan(d10) ? (T)0 : d10);
    volatile T t01 = w01 * (std::isnan(d01) ? (T)0 : d01), t11 = w11 * (std::isnan(d11) ? (T)0 : d11);
    volatile T n1 = t00 + t10; volatile T n2 = n1 + t01; volatile T num = n2 + t11;
    if (sw == (T)0) return std::numeric_limits<T>::quiet_NaN();
    return num / sw;
  };
  for (int64_t j = 0; j < d.ny; ++j)
    for (int64_t i = 0; i < d.nx; ++i) {
      T v;
      if (d.region_kind == NE_REGION_COLUMN) {       // read_data(…, ::ColumnInfo, …) (:164)
        if (d.column_interpolation == NE_COLUMN_NEAREST) {
          const T wx = (T)d.col_wx, wy = (T)d.col_wy;
          v = read(wx >= (T)0.5 ? d.col_i_plus : d.col_i_minus, wy >= (T)0.5 ? d.col_j_plus : d.col_j_minus);
          if (std::isnan(v)) v = blend_linear();
        } else {
          v = blend_linear();
        }
      } else {
        v = read(i + d.di, j + d.dj);                // read_data(…, ::BoundingBoxOffset, …) (:163)
      }
      switch (d.conv_kind[k]) {
        case NE_CONV_NEGATE: v = -v; break;
        case NE_CONV_ADD: v = v + a; break;
        case NE_CONV_SUB: v = v - a; break;
        case NE_CONV_MUL: v = v * a; break;
        case NE_CONV_DIV: v = v / a; break;
        case NE_CONV_MUL_DIV: { volatile T m = v * a; v = m / b; break; }
        default: break;
      }
      at(i, j) = v;
    }
  for (int64_t j = 0; j < d.ny; ++j)          // west / east halos of the interior rows
    for (int64_t h = 1; h <= d.hx; ++h) {
      at(-h, j) = d.periodic_x ? at(d.nx - h, j) : at(h - 1, j);
      at(d.nx - 1 + h, j) = d.periodic_x ? at(h - 1, j) : at(d.nx - h, j);
    }
  for (int64_t h = 1; h <= d.hy; ++h)         // south / north halos over the whole row, corners included
    for (int64_t i = -d.hx; i < d.nx + d.hx; ++i) {
      at(i, -h) = at(i, h - 1);
      at(i, d.ny - 1 + h) = at(i, d.ny - h);
    }
}

extern "C" {

int neo_version(void) { return NE_ABI_VERSION; }

void neo_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int neo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int neo_atmosphere_ocean_fluxes_f64(const NeAtmosOceanDesc* d) { NEO_DISPATCH_FLUX(atmosphere_ocean, double, d); return 0; }
int neo_atmosphere_ocean_fluxes_f32(const NeAtmosOceanDesc* d) { NEO_DISPATCH_FLUX(atmosphere_ocean, float, d); return 0; }
int neo_atmosphere_sea_ice_fluxes_f64(const NeAtmosSeaIceDesc* d) { NEO_DISPATCH_FLUX(atmosphere_sea_ice, double, d); return 0; }
int neo_atmosphere_sea_ice_fluxes_f32(const NeAtmosSeaIceDesc* d) { NEO_DISPATCH_FLUX(atmosphere_sea_ice, float, d); return 0; }
int neo_atmosphere_land_fluxes_f64(const NeAtmosLandDesc* d) { NEO_DISPATCH_FLUX(atmosphere_land, double, d); return 0; }
int neo_atmosphere_land_fluxes_f32(const NeAtmosLandDesc* d) { NEO_DISPATCH_FLUX(atmosphere_land, float, d); return 0; }

/* compute_interface_humidity of one AirLandInterfaceState (Float64): for the reference's own known-answer tests
 * (test/test_surface_fluxes.jl:340-423) */
double neo_land_interface_humidity(const NeLandHumidity* h, const NeThermoParams* thermo, double p_at, double q_at, double T_at,
                                   double T_skin, double T_bulk, double saturation, double ustar, double q_star, double q_prev) {
  const Thermo<double> th(*thermo);
  LandSurface<double> land = {h, saturation, T_bulk};
  State<double> s = {ustar, 0.0, q_star, 0.0, 0.0, T_skin, q_prev, saturation};
  AtmosState<double> a = {10.0, 0.0, 0.0, T_at, p_at, q_at, 512.0};
  return land_interface_humidity<double, double>(land, th, s, a, T_skin);
}
double neo_saturation_specific_humidity(const NeThermoParams* thermo, double T, double p, int phase) {
  const Thermo<double> th(*thermo);
  return saturation_specific_humidity<double>(th, T, p, phase);
}

int neo_series_slot_fill(const NeSeriesRingDesc* d, int32_t series, const void* raw, void* dst) {
  if (!d || !raw || !dst || series < 0 || series >= d->n_series) return -1;
  if (d->dtype == NE_F64) series_slot_fill<double>(*d, series, (const double*)raw, (double*)dst);
  else series_slot_fill<float>(*d, series, (const float*)raw, (float*)dst);
  return 0;
}

static int interp_dispatch(const NeInterpDesc* d, bool out64) {
  const bool a64 = d->src_dtype == NE_F64, t64 = d->time.frac_dtype == NE_F64;
  if (out64) {
    if (a64) { if (t64) interp_state<double, double, double>(*d); else interp_state<double, double, float>(*d); }
    else     { if (t64) interp_state<double, float, double>(*d);  else interp_state<double, float, float>(*d); }
  } else {
    if (a64) { if (t64) interp_state<float, double, double>(*d); else interp_state<float, double, float>(*d); }
    else     { if (t64) interp_state<float, float, double>(*d);  else interp_state<float, float, float>(*d); }
  }
  return 0;
}
int neo_interp_state_f64(const NeInterpDesc* d) { return interp_dispatch(d, true); }
int neo_interp_state_f32(const NeInterpDesc* d) { return interp_dispatch(d, false); }

int neo_frac_indices_f64(const NeFracIndexDesc* d) {
  if (d->src_dtype == NE_F64) frac_indices<double, double>(*d); else frac_indices<double, float>(*d);
  return 0;
}
int neo_frac_indices_f32(const NeFracIndexDesc* d) {
  if (d->src_dtype == NE_F64) frac_indices<float, double>(*d); else frac_indices<float, float>(*d);
  return 0;
}

int neo_sea_ice_ocean_fluxes_f64(const NeSeaIceOceanDesc* d) { sea_ice_ocean<double>(*d); return 0; }
int neo_sea_ice_ocean_fluxes_f32(const NeSeaIceOceanDesc* d) { sea_ice_ocean<float>(*d); return 0; }
int neo_sea_ice_ocean_stress_f64(const NeSeaIceOceanStressDesc* d) { sea_ice_ocean_stress<double>(*d); return 0; }
int neo_sea_ice_ocean_stress_f32(const NeSeaIceOceanStressDesc* d) { sea_ice_ocean_stress<float>(*d); return 0; }
int neo_assemble_net_ocean_fluxes_f64(const NeAssembleOceanDesc* d) { assemble_ocean<double>(*d); return 0; }
int neo_assemble_net_ocean_fluxes_f32(const NeAssembleOceanDesc* d) { assemble_ocean<float>(*d); return 0; }
int neo_assemble_net_sea_ice_fluxes_f64(const NeAssembleSeaIceDesc* d) { assemble_sea_ice<double>(*d); return 0; }
int neo_assemble_net_sea_ice_fluxes_f32(const NeAssembleSeaIceDesc* d) { assemble_sea_ice<float>(*d); return 0; }
int neo_correct_atmosphere_elevation_f64(const NeElevationCorrectionDesc* d) { correct_atmosphere_elevation<double>(*d); return 0; }
int neo_correct_atmosphere_elevation_f32(const NeElevationCorrectionDesc* d) { correct_atmosphere_elevation<float>(*d); return 0; }
int neo_apply_radiative_fluxes_f64(const NeApplyRadiationDesc* d) { apply_radiation<double>(*d); return 0; }
int neo_apply_radiative_fluxes_f32(const NeApplyRadiationDesc* d) { apply_radiation<float>(*d); return 0; }

// ---- scalar probes used by the known-answer tests (tests/test_oracle_reference_kats.py) ---------
double neo_stability_f64(const NeStabilityProfile* s, double zeta) { return stability_profile<double, double>(*s, zeta); }
double neo_vsgs2_f64(const NeSubgridVelocity* s, double ustar, double bstar, double h_bl) { return vsgs2<double>(*s, ustar, bstar, h_bl); }
double neo_polynomial_drag_f64(const NePolynomialDrag* p, double U) { return polynomial_drag<double>(*p, U); }
double neo_momentum_roughness_f64(const NeRoughnessLength* r, double ustar, double U) {
  return momentum_roughness<double, double>(*r, r->nu, ustar, U);
}
double neo_scalar_roughness_f64(const NeRoughnessLength* r, double ell_u, double ustar) {
  return scalar_roughness<double, double>(*r, r->nu, ell_u, ustar);
}
double neo_saturation_vapor_pressure_f64(const NeThermoParams* t, double T, int phase) {
  return Thermo<double>(*t).saturation_vapor_pressure(T, phase);
}
double neo_surface_specific_humidity_f64(const NeInterfaceProperties* ip, const NeThermoParams* t, double p, double T, double S) {
  return surface_specific_humidity<double, double>(*ip, Thermo<double>(*t), p, T, S);
}
void neo_interpolator_f64(double f, int64_t* im, int64_t* ip, double* xi) {
  auto it = interpolator<double>(f); *im = it.im; *ip = it.ip; *xi = it.xi;
}
void neo_interpolator_f32(float f, int64_t* im, int64_t* ip, float* xi) {
  auto it = interpolator<float>(f); *im = it.im; *ip = it.ip; *xi = it.xi;
}

// op census control
void neo_count_ops(int enable) { g_count_ops = enable != 0; }
void neo_reset_op_counts(void) { g_ops = OpCounts{0, 0, 0, 0, 0, 0, 0, 0}; }
void neo_get_op_counts(uint64_t out[8]) {
  out[0] = g_ops.exp_; out[1] = g_ops.log_; out[2] = g_ops.pow_; out[3] = g_ops.atan_;
  out[4] = g_ops.cbrt_; out[5] = g_ops.sqrt_; out[6] = g_ops.iters; out[7] = g_ops.points;
}

}  // extern "C"
