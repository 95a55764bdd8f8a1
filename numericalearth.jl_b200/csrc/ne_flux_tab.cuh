// ne_flux_tab.cuh — table-driven iteration of the specialised a–o solve (default plugin tree).
//
// Same fixed point, same stopping rule and iteration count as ne_flux_fast.cuh (and therefore as
// compute_interface_state.jl:5-58 / similarity_theory_turbulent_fluxes.jl:315-385); what changes is
// HOW the elementary functions are evaluated (ne_fastmath.cuh): branch-free log/exp/cbrt/sqrt/rcp
// whose coefficients are constant-bank operands, and the unstable Edson ψ_m, ψ_s read from
// piecewise degree-13 polynomial tables staged in shared memory.  One iteration is ~190 FP64
// instructions with no data-dependent branch except (i) the stable side ζ ≥ 2^-6 (closed form with
// the custom exp) and (ii) |ζ| ≥ 2^7 on the unstable side (closed form through libdevice, rare).
#pragma once

#include "ne_fastmath.cuh"
#include "ne_flux_fast.cuh"

namespace ne {

struct TabParams {
  fm::MathConsts mc;
  double cbrt_floor;     // (gmin/β)³/8: below it β·cbrt(x) < gmin, so the clamp cannot change U_G
  double log_hd;         // log(surface_layer_height − d) when the height is a scalar (set per launch)
  int32_t same_exp;      // ψ_m and ψ_s stable branches share exp(−min(ζmax, A⁺ζ))
  int32_t general_psi;   // tables fitted to non-Edson stability functions: outside the table → generic closed forms
  int32_t f32_model;     // Float32 model: the generic closed forms take the Float32-rounded parameters
  int32_t far_fm;        // Edson far-unstable closed forms through the branch-free functions (set per launch)
  // Edson stable closed forms above ζ_sat = max(ζmax/A⁺)(1 + 1e-12): both exponentials equal exp(−ζmax) (ne_flux_tab2.cuh)
  double z_sat, em_sat, es_sat;
};

// ---- host: Chebyshev interpolation → monomials in w ∈ [−1, 1] (long double) -------------------------
template <class F>
inline void cheb_fit_monomial(F f, long double lo, long double hi, int deg, double* coef) {
  constexpr int MAXN = 24;
  const int N = deg + 1;
  const long double PI = 3.141592653589793238462643383279502884L;
  const long double mid = (lo + hi) / 2, half = (hi - lo) / 2;
  long double y[MAXN], c[MAXN];
  for (int k = 0; k < N; ++k) y[k] = f(mid + half * cosl(PI * (2 * k + 1) / (2 * N)));
  for (int j = 0; j < N; ++j) {
    long double a = 0;
    for (int k = 0; k < N; ++k) a += y[k] * cosl(PI * j * (2 * k + 1) / (2 * N));
    c[j] = a * 2 / N;
  }
  c[0] /= 2;
  long double T0[MAXN] = {0}, T1[MAXN] = {0}, T2[MAXN], px[MAXN] = {0};
  T0[0] = 1; T1[1] = 1;
  px[0] += c[0];
  if (N > 1) for (int m = 0; m < N; ++m) px[m] += c[1] * T1[m];
  for (int j = 2; j < N; ++j) {
    for (int m = 0; m < N; ++m) T2[m] = (m > 0 ? 2 * T1[m - 1] : 0) - T0[m];
    for (int m = 0; m < N; ++m) { px[m] += c[j] * T2[m]; T0[m] = T1[m]; T1[m] = T2[m]; }
  }
  for (int m = 0; m < N; ++m) coef[m] = (double)px[m];
}

// f32: the Float32 model evaluates Cp*Dp as a Float32 product before promotion (psi_edson_momentum<float, double>)
inline long double psi_m_stable_ld(const double* p, long double z, bool f32 = false) {
  const long double zmax = p[0], Ap = p[1], Bp = p[2], Cp = p[3], Dp = p[4];
  const long double dz = fminl(zmax, Ap * z);
  const long double CpDp = f32 ? (long double)((float)p[3] * (float)p[4]) : Cp * Dp;
  return -Bp * z - Cp * (z - Dp) * expl(-dz) - CpDp;
}
inline long double psi_s_stable_ld(const double* p, long double z) {
  const long double zmax = p[0], Ap = p[1], Bp = p[2], Cp = p[3], Dp = p[4], Ep = p[5];
  const long double dz = fminl(zmax, Ap * z);
  return -powl(1 + Bp * z, Cp) - Bp * (z - Dp) * expl(-dz) - Ep;
}

// Fills tab[TAB_SIZE] and T; returns the max error of the ψ polynomials (double evaluation vs the
// long-double closed forms, relative to max(1, |ψ|)) over all intervals.
// ---- every stability-function kind in long double (host; follows stability_fn<double, double>, ne_physics.cuh) ----
inline long double psi_fn_ld(const NeStabilityFn& f, long double z) {
  const double* p = f.p;
  switch (f.kind) {
    case NE_PSI_ZERO: return 0;
    case NE_PSI_EDSON_MOMENTUM: return z < 0 ? psi_m_unstable_ld(p, z) : psi_m_stable_ld(p, z);
    case NE_PSI_EDSON_SCALAR: return z < 0 ? psi_s_unstable_ld(p, z) : psi_s_stable_ld(p, z);
    case NE_PSI_SHEBA_MOMENTUM: {   // similarity_theory_turbulent_fluxes.jl:643-657
      const long double a = p[0], b = p[1], zp = fmaxl(0, z);
      const long double zz = cbrtl(1 + zp), B = cbrtl((1 - b) / b), rt3 = 1.7320508075688772;
      const long double P1 = -3 * a * (zz - 1) / b;
      const long double P2 = a * B / (2 * b) *
          (2 * logl((zz + B) / (1 + B)) - logl((zz * zz - B * zz + B * B) / (1 - B + B * B)) +
           2 * rt3 * (atanl((2 * zz - B) / (rt3 * B)) - atanl((2 - B) / (rt3 * B))));
      return P1 + P2;
    }
    case NE_PSI_SHEBA_SCALAR: {     // :665-677
      const long double a = p[0], b = p[1], c = p[2], B = sqrtl(c * c - 4), zp = fmaxl(0, z);
      const long double P1 = -b / 2 * logl(1 + c * zp + zp * zp);
      const long double P2 = (b * c / (2 * B) - a / B) * (logl((2 * zp + c - B) / (2 * zp + c + B)) - logl((c - B) / (c + B)));
      return P1 + P2;
    }
    case NE_PSI_PAULSON_MOMENTUM: { // :688-699
      const long double a = p[0], b = p[1], zm = fminl(0, z), zz = sqrtl(sqrtl(1 - a * zm));
      return 2 * logl((1 + zz) / 2) + logl((1 + zz * zz) / 2) - 2 * atanl(zz) + b;
    }
    case NE_PSI_PAULSON_SCALAR: {   // :705-710
      const long double a = p[0], zm = fminl(0, z), zz = sqrtl(sqrtl(1 - a * zm));
      return 2 * logl((1 + zz * zz) / 2);
    }
    case NE_PSI_LINEAR_STABLE: {    // :747-752
      const long double c = p[0], zmax = p[1], zp = fmaxl(0, z);
      return -c * fminl(zp, zmax);
    }
  }
  return NAN;
}
inline long double psi_profile_ld(const NeStabilityProfile& s, long double z) {   // SplitStabilityFunction :720-725
  if (!s.split) return psi_fn_ld(s.a, z);
  return z > 0 ? psi_fn_ld(s.a, z) : psi_fn_ld(s.b, z);
}
// one side of a profile as a function of |ζ|, continued to |ζ| = 0 from its own side (a SplitStabilityFunction
// evaluates ζ = 0 with its unstable member, :720-725: a one-point discontinuity of the size of the members' rounding
// at 0 — 1e-16 in Float64, 4e-8 with Float32-rounded parameters — that a polynomial should not chase)
inline long double psi_side_ld(const NeStabilityProfile& s, long double az, bool stable) {
  if (!s.split) return psi_fn_ld(s.a, stable ? az : -az);
  return stable ? psi_fn_ld(s.a, az) : psi_fn_ld(s.b, -az);
}
inline bool psi_is_plain_edson(const NeFluxFormulation& f) {
  return !f.psi_momentum.split && f.psi_momentum.a.kind == NE_PSI_EDSON_MOMENTUM &&
         !f.psi_temperature.split && f.psi_temperature.a.kind == NE_PSI_EDSON_SCALAR;
}

// f32: tables for the Float32 model — the ψ parameters are the Float32-rounded ones (and √E⁻, C⁺D⁺ are Float32
// operations), exactly what stability_fn<float, double> evaluates.
inline double build_solver_tables(const NeFluxFormulation& f, double* tab, TabParams& T, bool f32 = false) {
  using namespace fm;
  // log table
  for (int i = 0; i < LOG_N; ++i) {
    const double zlo = mk64(0x3fe6a09e + (i << 14), 0), zhi = mk64(0x3fe6a09e + ((i + 1) << 14), 0);
    const double c = 0.5 * (zlo + zhi);
    const double invc = 1.0 / c;
    tab[TAB_LOG + 2 * i] = invc;
    tab[TAB_LOG + 2 * i + 1] = (double)(-logl((long double)invc));
  }
  for (int j = 0; j < EXP2_N; ++j) tab[TAB_EXP2 + j] = (double)exp2l((long double)j / EXP2_N);
  fill_literals(T.mc);
  for (int k = 0; k < LOG_DEG; ++k) T.mc.logp[k] = ((k & 1) ? 1.0 : -1.0) / (double)(k + 2);   // −1/2, +1/3, −1/4 …
  {  // e^r on |r| ≤ 0.35, monomials in r
    double cw[EXP_DEG + 1];
    const long double h = 0.35L;
    cheb_fit_monomial([](long double x) { return expl(x); }, -h, h, EXP_DEG, cw);
    long double s = 1;
    for (int k = 0; k <= EXP_DEG; ++k) { T.mc.expp[k] = (double)((long double)cw[k] / s); s *= h; }
  }
  double pm[12], ps[12];
  for (int k = 0; k < 12; ++k) {
    pm[k] = f32 ? (double)(float)f.psi_momentum.a.p[k] : f.psi_momentum.a.p[k];
    ps[k] = f32 ? (double)(float)f.psi_temperature.a.p[k] : f.psi_temperature.a.p[k];
  }
  // any other combination of the shipped stability functions (SHEBA / Paulson / linear-stable / split / zero):
  // the same piecewise polynomials fitted to the general closed forms (Float64 models only)
  const bool general = !psi_is_plain_edson(f);
  // Float32 model with non-Edson functions: the closed forms at the Float32-rounded parameters (the generic kernel also
  // forms a few derived constants in Float32: a 1e-7 relative difference in ψ, far inside the Float32 bar of 1e-5)
  NeStabilityProfile gm = f.psi_momentum, gs = f.psi_temperature;
  if (general && f32)
    for (NeStabilityProfile* pr : {&gm, &gs})
      for (NeStabilityFn* fn : {&pr->a, &pr->b})
        for (int k = 0; k < 12; ++k) fn->p[k] = (double)(float)fn->p[k];
  double worst = 0;
  for (int iv = 0; iv < PSI_NI; ++iv) {
    long double lo, hi;
    const bool stable = iv >= PSI_NS;
    const int r = stable ? iv - PSI_NS : iv;
    if (r == 0) { lo = 0; hi = ldexpl(1, PSI_OCT_LO); }
    else {
      const int q = r - 1, e = PSI_OCT_LO + (q >> PSI_SUB_BITS), s = q & (PSI_SUB - 1);
      lo = ldexpl(1 + s / (long double)PSI_SUB, e); hi = ldexpl(1 + (s + 1) / (long double)PSI_SUB, e);
    }
    double* rec = tab + TAB_PSI + iv * PSI_REC;
    const long double half = (hi - lo) / 2, mid = (hi + lo) / 2;
    rec[0] = (double)(1 / half);
    rec[1] = (double)(-mid / half);
    double cm[PSI_DEG + 1], cs[PSI_DEG + 1];
    auto fmf = [&](long double az) {
      if (general) return psi_side_ld(gm, az, stable);
      return stable ? psi_m_stable_ld(pm, az, f32) : psi_m_unstable_ld(pm, -az, f32);
    };
    auto fsf = [&](long double az) {
      if (general) return psi_side_ld(gs, az, stable);
      return stable ? psi_s_stable_ld(ps, az) : psi_s_unstable_ld(ps, -az, f32);
    };
    cheb_fit_monomial(fmf, lo, hi, PSI_DEG, cm);
    cheb_fit_monomial(fsf, lo, hi, PSI_DEG, cs);
    for (int k = 0; k <= PSI_DEG; ++k) { rec[2 + 2 * k] = cm[k]; rec[2 + 2 * k + 1] = cs[k]; }
    for (int n = 0; n <= 48; ++n) {
      const double az = (double)(lo + (hi - lo) * n / 48.0L);
      double m, s;
      psi_pair(rec, az, m, s);
      const long double tm = fmf(az), ts = fsf(az);   // error relative to max(1, |ψ|): Π = log(h/ℓ) − ψ + ψ is O(1…|ψ|)
      const double em = (double)(fabsl((long double)m - tm) / fmaxl(1, fabsl(tm)));
      const double es = (double)(fabsl((long double)s - ts) / fmaxl(1, fabsl(ts)));
      if (!(em <= worst)) worst = em;   // also catches NaN
      if (!(es <= worst)) worst = es;
    }
  }
  for (int side = 0; side < 2; ++side) {   // |ζ| < 2^TINY_EXP records
    const bool stable = side == 1;
    const long double lo = 0, hi = ldexpl(1, TINY_EXP);
    double* rec = tab + TAB_TINY + side * TINY_REC;
    const long double half = (hi - lo) / 2, mid = (hi + lo) / 2;
    rec[0] = (double)(1 / half);
    rec[1] = (double)(-mid / half);
    double cm[TINY_DEG + 1], cs[TINY_DEG + 1];
    auto fmf = [&](long double az) {
      if (general) return psi_side_ld(gm, az, stable);
      return stable ? psi_m_stable_ld(pm, az, f32) : psi_m_unstable_ld(pm, -az, f32);
    };
    auto fsf = [&](long double az) {
      if (general) return psi_side_ld(gs, az, stable);
      return stable ? psi_s_stable_ld(ps, az) : psi_s_unstable_ld(ps, -az, f32);
    };
    cheb_fit_monomial(fmf, lo, hi, TINY_DEG, cm);
    cheb_fit_monomial(fsf, lo, hi, TINY_DEG, cs);
    for (int k = 0; k <= TINY_DEG; ++k) { rec[2 + 2 * k] = cm[k]; rec[2 + 2 * k + 1] = cs[k]; }
    for (int n = 0; n <= 48; ++n) {
      const double az = (double)(lo + (hi - lo) * n / 48.0L);
      double m, s;
      psi_tiny_pair(rec, az, az, m, s);
      const double em = (double)fabsl((long double)m - fmf(az)), es = (double)fabsl((long double)s - fsf(az));
      if (!(em <= worst)) worst = em;
      if (!(es <= worst)) worst = es;
    }
  }
  for (int side = 0; side < 2; ++side) {   // |ζ| < 2^MICRO_EXP records
    const bool stable = side == 1;
    const long double lo = 0, hi = ldexpl(1, MICRO_EXP);
    double* rec = tab + TAB_MICRO + side * MICRO_REC;
    const long double half = (hi - lo) / 2, mid = (hi + lo) / 2;
    rec[0] = (double)(1 / half);
    rec[1] = (double)(-mid / half);
    double cm[MICRO_DEG + 1], cs[MICRO_DEG + 1];
    auto fmf = [&](long double az) {
      if (general) return psi_side_ld(gm, az, stable);
      return stable ? psi_m_stable_ld(pm, az, f32) : psi_m_unstable_ld(pm, -az, f32);
    };
    auto fsf = [&](long double az) {
      if (general) return psi_side_ld(gs, az, stable);
      return stable ? psi_s_stable_ld(ps, az) : psi_s_unstable_ld(ps, -az, f32);
    };
    cheb_fit_monomial(fmf, lo, hi, MICRO_DEG, cm);
    cheb_fit_monomial(fsf, lo, hi, MICRO_DEG, cs);
    for (int k = 0; k <= MICRO_DEG; ++k) { rec[2 + 2 * k] = cm[k]; rec[2 + 2 * k + 1] = cs[k]; }
    for (int n = 0; n <= 48; ++n) {
      const double az = (double)(lo + (hi - lo) * n / 48.0L);
      double m, s2;
      psi_micro_pair(rec, az, az, m, s2);
      const double em = (double)fabsl((long double)m - fmf(az)), es = (double)fabsl((long double)s2 - fsf(az));
      if (!(em <= worst)) worst = em;
      if (!(es <= worst)) worst = es;
    }
  }
  const NeSubgridVelocity& g = f.subgrid_velocities;
  const double ratio = g.minimum_gustiness / g.gustiness_parameter;
  T.cbrt_floor = ratio * ratio * ratio / 8;
  T.same_exp = (pm[0] == ps[0] && pm[1] == ps[1]);
  T.general_psi = general ? 1 : 0;
  T.f32_model = f32 ? 1 : 0;
  T.far_fm = 0;
  T.log_hd = 0;
  T.z_sat = INFINITY; T.em_sat = T.es_sat = 0;
  if (!general && pm[1] > 0 && ps[1] > 0) {
    T.z_sat = std::fmax(pm[0] / pm[1], ps[0] / ps[1]) * (1 + 1e-12);
    T.em_sat = fm::exp_mid(T.mc, -pm[0]);   // the same operations as on the device: the same bits
    T.es_sat = fm::exp_mid(T.mc, -ps[0]);
  }
  return worst;
}

inline bool tab_path_eligible(const NeFluxFormulation& f) {
  const NeSubgridVelocity& g = f.subgrid_velocities;
  const double ratio = g.minimum_gustiness / g.gustiness_parameter;
  if (!(g.gustiness_parameter > 0) || !(ratio * ratio * ratio / 8 > 1e-290)) return false;
  return true;
}

// ---- Edson unstable closed forms for ζ ≤ −2^7 with the branch-free elementary functions -------------------------------
// The second trip of an unstable point starts from a tiny u★ and lands at ζ ~ −10^3…−10^5, outside the tables: more than
// half of the active warps pass here once (ncu source counters: 88 076 of 163 178 warp-tiles on C4), and the libdevice
// closed forms cost 805 SASS instructions per pass (5.9 % of the kernel).  Same formulas
// (similarity_theory_turbulent_fluxes.jl:501-532, 586-618) with log_pos / sqrt_pos / cbrt_pos and an arctangent that only
// has to handle large arguments: atan(x) = π/2 − atan(1/x), 1/x ≤ 0.16 here, odd Taylor series to x⁻¹⁹.
// Requires B⁻ = 2 for ψ_m (P.m_B2; the only value the reference ships) — the caller falls back to libdevice otherwise.
namespace fm {
NE_HD double atan_large(double x) {   // x ≥ 6: |error| ≲ 1 ulp of π/2
  const double y = rcp(x), y2 = y * y;
  double p = -1.0 / 19.0;
  p = fma_(p, y2, 1.0 / 17.0);
  p = fma_(p, y2, -1.0 / 15.0);
  p = fma_(p, y2, 1.0 / 13.0);
  p = fma_(p, y2, -1.0 / 11.0);
  p = fma_(p, y2, 1.0 / 9.0);
  p = fma_(p, y2, -1.0 / 7.0);
  p = fma_(p, y2, 1.0 / 5.0);
  p = fma_(p, y2, -1.0 / 3.0);
  p = fma_(p, y2, 1.0);
  return 1.5707963267948966 - y * p;
}
}  // namespace fm

NE_HD void psi_far_unstable_fm(const FastParams& P, const TabParams& T, const double* tab, double z, double& pm, double& ps) {
  const double z2 = z * z;
  const double fw = 1.0 - fm::rcp(1.0 + z2);                      // ζ²/(1 + ζ²)
  {  // momentum
    const double f1 = fm::sqrt_pos(fm::sqrt_pos(1.0 - P.m_Am * z));
    const double f1s = f1 * f1;
    const double psi1 = fm::log_pos(tab, T.mc, (1.0 + f1) * (1.0 + f1) * (1.0 + f1s) * 0.125) - 2.0 * fm::atan_large(f1) + P.m_Cm;
    const double f2 = fm::cbrt_pos(T.mc, 1.0 - P.m_Dm * z);
    const double psi2 = P.m_halfEm * fm::log_pos(tab, T.mc, (1.0 + f2 + f2 * f2) * P.m_iEm) -
                        P.m_rEm * fm::atan_large((1.0 + 2.0 * f2) * P.m_irEm) + P.m_Fm;
    pm = psi1 + fw * (psi2 - psi1);
  }
  {  // scalar
    const double f1 = fm::sqrt_pos(1.0 - P.s_Am * z);
    const double psi1 = P.s_Bm * fm::log_pos(tab, T.mc, (1.0 + f1) * P.s_iBm) + P.s_Cm;
    const double f2 = fm::cbrt_pos(T.mc, 1.0 - P.s_Dm * z);
    const double psi2 = P.s_halfEm * fm::log_pos(tab, T.mc, (1.0 + f2 + f2 * f2) * P.s_iEm) -
                        P.s_rEm * fm::atan_large((1.0 + 2.0 * f2) * P.s_irEm) + P.s_Fm;
    ps = psi1 + fw * (psi2 - psi1);
  }
}
// the arguments of atan_large above must stay ≥ 6 for ζ ≤ −2^7: checked on the host for the user's parameters
inline bool far_unstable_fm_ok(const FastParams& P) {
  if (!P.m_B2 || !(P.m_Am > 0) || !(P.m_Dm > 0) || !(P.s_Am > 0) || !(P.s_Dm > 0)) return false;
  const double z = -128.0;
  const double f1 = std::sqrt(std::sqrt(1.0 - P.m_Am * z));
  const double a_m = (1.0 + 2.0 * std::cbrt(1.0 - P.m_Dm * z)) * P.m_irEm, a_s = (1.0 + 2.0 * std::cbrt(1.0 - P.s_Dm * z)) * P.s_irEm;
  return f1 >= 6.0 && a_m >= 6.0 && a_s >= 6.0 && P.m_iEm > 0 && P.s_iEm > 0 && P.s_iBm > 0;
}

#if defined(__CUDACC__)
// unstable closed forms through libdevice for |ζ| ≥ 2^7 (parameter sets the branch-free version does not cover)
__device__ __forceinline__ void psi_far_unstable(const FastParams& P, double z, double& pm, double& ps) {
  pm = fast_psi_m(P, z);   // |z| ≥ 2^7 > P.zsmall: the closed-form branch
  ps = fast_psi_s(P, z);
}

static __device__ __noinline__ double pow_general(double x, double y) { return pow(x, y); }

// stable closed forms (ζ ≥ 2^-6) with the custom exp/sqrt
__device__ __forceinline__ void psi_stable_pair(const FastParams& P, const TabParams& T, double z, double& pm, double& ps) {
  const double em = fm::exp_mid(T.mc, -fmin(P.m_zmax, P.m_Ap * z));
  const double es = T.same_exp ? em : fm::exp_mid(T.mc, -fmin(P.s_zmax, P.s_Ap * z));
  pm = -P.m_Bp * z - P.m_Cp * (z - P.m_Dp) * em - P.m_CpDp;
  const double x = 1.0 + P.s_Bp * z;
  const double xp = P.s_C15 ? x * fm::sqrt_pos(x) : pow_general(x, P.s_Cp);
  ps = -xp - P.s_Bp * (z - P.s_Dp) * es - P.s_Ep;
}

// closed forms outside the table: stable side with the custom exp, far unstable side through libdevice.
// Out of line and returning by value, so the common path keeps ψ in registers.
// ff != nullptr: tables of a non-Edson pair — the generic closed forms of ne_physics.cuh.
static __device__ __noinline__ double2 psi_outside(const FastParams& P, const TabParams& T, const NeFluxFormulation* ff,
                                                   const double* tab, double z) {
  double pm, ps;
  if (ff && T.f32_model) {
    pm = stability_profile<float, double>(ff->psi_momentum, z);
    ps = stability_profile<float, double>(ff->psi_temperature, z);
  } else if (ff) {
    pm = stability_profile<double, double>(ff->psi_momentum, z);
    ps = stability_profile<double, double>(ff->psi_temperature, z);
  } else if (z > 0) psi_stable_pair(P, T, z, pm, ps);
  else if (T.far_fm) psi_far_unstable_fm(P, T, tab, z, pm, ps);
  else psi_far_unstable(P, z, pm, ps);
  return make_double2(pm, ps);
}

// ψ_m(ζ), ψ_s(ζ) at the same ζ
__device__ __forceinline__ void tab_psi_pair(const FastParams& P, const TabParams& T, const NeFluxFormulation* ff,
                                             const double* tab, double z, double& pm, double& ps) {
  bool outside;
  const int iv = fm::psi_interval(z, outside);
  if (!outside) {
    fm::psi_pair(tab + fm::TAB_PSI + iv * fm::PSI_REC, fabs(z), pm, ps);
  } else {
    const double2 r = psi_outside(P, T, ff, tab, z);
    pm = r.x; ps = r.y;
  }
}

// out-of-line general lookup for the case where ψ(ℓ/L★) leaves the tiny-|ζ| records (the first trips, while u★ is still
// tiny: 7 % of the warp-trips on C4): ψ_m at ζ_u and ψ_s at ζ_s, ONE call, each evaluating only the polynomial it needs
static __device__ __noinline__ double2 tab_psi_rare2(const FastParams& P, const TabParams& T, const NeFluxFormulation* ff,
                                              const double* tab, double zu, double zs) {
  double pm, ps;
  bool outside;
  int iv = fm::psi_interval(zu, outside);
  if (!outside) pm = fm::psi_single(tab + fm::TAB_PSI + iv * fm::PSI_REC, fabs(zu), 0);
  else pm = psi_outside(P, T, ff, tab, zu).x;
  iv = fm::psi_interval(zs, outside);
  if (!outside) ps = fm::psi_single(tab + fm::TAB_PSI + iv * fm::PSI_REC, fabs(zs), 1);
  else ps = psi_outside(P, T, ff, tab, zs).y;
  return make_double2(pm, ps);
}

// The Float64 core shared by the Float64 and the mixed-precision Float32 iterations: roughness lengths, the
// two logarithmic profiles with their ψ corrections and the transfer coefficients χ = ϰ/Π.
//   u★ (as a double), ru = 1/u★, ℓu (already clipped), 1/L★, Δh = z − d and log Δh  →  χ_u, χ_s
// EXT: the options beyond the strict default tree (wind-dependent waves, constant mesoscale term) and a domain guard —
// compiled out of the default-tree kernels, whose iterates stay positive.  (The COARE profile stays on the generic
// kernel: without the ψ(ℓ/L★) term Π changes sign on the way from the 1e-4 initial guess at ~1 % of the points, and
// what the reference then does with negative roughness lengths and NaNs is not worth imitating in a fast path.)
template <bool EXT = false>
__device__ __forceinline__ void tab_core(const FastParams& P, const TabParams& T, const double* tab, double ustar,
                                         double ru, double lu, double Linv, double hd, double log_hd,
                                         double& chi_u, double& chi_s, const NeFluxFormulation* ff = nullptr) {
  using fm::dmax;
  (void)ru;
  const double log_lu = fm::log_pos(tab, T.mc, lu);
  using fm::mul_;
  const double Rs = mul_(mul_(lu, ustar), P.nu_inv);
  const double log_Rs = fm::log_pos(tab, T.mc, Rs);
  const double log_ls_un = fm::fma_(-P.rb, log_Rs, P.log_rA);
  const bool clipped = log_ls_un > P.log_ls_max;
  const double log_ls = clipped ? P.log_ls_max : log_ls_un;
  const double ls_un = fm::exp_mid(T.mc, dmax(log_ls_un, -700.0));
  const double ls = clipped ? P.ls_max : ls_un;
  const bool lifted = 2.0 * lu > hd;
  const double dh = lifted ? 2.0 * lu : hd;
  const double log_dh = lifted ? T.mc.ln2 + log_lu : log_hd;
  double pm_h, ps_h, pm_l, ps_l;
  tab_psi_pair(P, T, ff, tab, mul_(dh, Linv), pm_h, ps_h);
  // ψ(ℓ/L★): |ℓ/L★| < 2^-12 except in the first trips from the 1e-4 initial guess
  const double zu = mul_(lu, Linv), zs = mul_(ls, Linv);
  if (fm::psi_is_micro(zu) && fm::psi_is_micro(zs)) {
    fm::psi_micro_pair(tab + fm::TAB_MICRO + (Linv < 0 ? 0 : fm::MICRO_REC), fabs(zu), fabs(zs), pm_l, ps_l);
  } else if (fm::psi_is_tiny(zu) && fm::psi_is_tiny(zs)) {
    fm::psi_tiny_pair(tab + fm::TAB_TINY + (Linv < 0 ? 0 : fm::TINY_REC), fabs(zu), fabs(zs), pm_l, ps_l);
  } else {
    const double2 r2 = tab_psi_rare2(P, T, ff, tab, zu, zs);
    pm_l = r2.x; ps_l = r2.y;
  }
  const double Pi_u = (log_dh - log_lu) - pm_h + pm_l;
  const double Pi_s = (log_dh - log_ls) - ps_h + ps_l;
  // χ = ϰ/Π for both profiles from ONE reciprocal: r = 1/(Π_u Π_s), ϰ/Π_u = ϰ Π_s r, ϰ/Π_s = ϰ Π_u r (each with a
  // residual correction, ≲ 1.5 ulp)
  const double r = fm::rcp(mul_(Pi_u, Pi_s));
  const double ru_ = mul_(Pi_s, r), rs_ = mul_(Pi_u, r);   // 1/Π_u, 1/Π_s
  chi_u = mul_(P.kappa, ru_); chi_s = mul_(P.kappa, rs_);
  chi_u = fm::fma_(fm::fma_(-Pi_u, chi_u, P.kappa), ru_, chi_u);
  chi_s = fm::fma_(fm::fma_(-Pi_s, chi_s, P.kappa), rs_, chi_s);
  if (EXT) {
    // Outside the domain of the branch-free log (a non-positive ℓu or R★: an iterate that went through Π ≤ 0, which
    // the non-default profiles allow) the reference's log / pow return NaN and the point stays NaN: do the same.
    if (!((lu > 0.0) & (Rs > 0.0))) { chi_u = __longlong_as_double(0x7ff8000000000000LL); chi_s = chi_u; }
  }
}

template <bool EXT = false>
__device__ __forceinline__ void tab_iteration(const FastParams& P, const TabParams& T, const double* tab, FastPoint& s,
                                              const NeFluxFormulation* ff = nullptr) {
  using fm::dmax;
  using fm::dmin;
  using fm::fma_;
  using fm::mul_;
  // every product is separately rounded or an explicit fma: the iterate does not depend on the kernel this is inlined in
  const double bstar = mul_(s.gTv, fma_(s.theta_star, s.c1, mul_(s.c2, s.q_star)));
  const double Jb = -mul_(s.ustar, bstar);
  const double UG = dmax(P.gmin, mul_(P.beta, fm::cbrt_pos(T.mc, dmax(mul_(dmax(0.0, Jb), s.h_bl), T.cbrt_floor))));
  const double U = fm::sqrt_pos(EXT ? fma_(UG, UG, s.dudv2) + P.sgs_const2 : fma_(UG, UG, s.dudv2));
  const double ru = fm::rcp(s.ustar);
  // 𝒞g/g: constant, or WindDependentWaveFormulation (roughness_lengths.jl:75) with the gusty wind speed U
  const double a1 = (EXT && P.wind_waves) ? mul_(dmax(0.0, fma_(P.wave_C1, dmin(U, P.wave_Umax), P.wave_C2)), P.inv_g_rough) : P.a1;
  const double lu = (EXT && P.const_rough) ? P.lu_c : dmin(fma_(mul_(a1, s.ustar), s.ustar, mul_(P.a2, ru)), P.lmax);
  const double Linv = mul_(mul_(mul_(P.kappa, bstar), ru), ru);   // 1/L★ (0 when b★ == 0, i.e. L★ = Inf)
  double chi_u, chi_s;
  tab_core<EXT>(P, T, tab, s.ustar, ru, lu, Linv, s.hd, s.log_hd, chi_u, chi_s, ff);
  s.ustar = mul_(chi_u, U);
  s.theta_star = mul_(chi_s, s.dtheta);
  s.q_star = mul_(chi_s, s.dq);
}

// ---- Float32 model, default tree --------------------------------------------------------------------
// The reference's promotion rules (the 1.5e-5 air viscosity is a Float64 literal,
// roughness_lengths.jl:94,126) make the Float32 iteration a MIXED-precision one: b★, the gustiness, U and L★ are
// Float32; the roughness lengths, log(Δh/ℓ), every ψ and χ = ϰ/Π are Float64 (with the Float32-rounded plugin
// parameters); u★, θ★, q★ are rounded to Float32 at the end of every trip.  The Float32 front below keeps each
// operation separately rounded (no FMA contraction), as the reference does.
struct FrontF32 {
  float gmin, beta, Cg, g_rough, kappa, tol;
  float g, d_zero;
};
struct FastPointF {
  float gTv, c1, c2, dudv2, h_bl, hd, dtheta, dq;
  double log_hd;
  float ustar, theta_star, q_star;
};

__device__ __forceinline__ void tab_iteration(const FastParams& P, const FrontF32& Q, const TabParams& T,
                                              const double* tab, FastPointF& s, const NeFluxFormulation* ff = nullptr) {
  const float bstar = __fmul_rn(s.gTv, __fadd_rn(__fmul_rn(s.theta_star, s.c1), __fmul_rn(s.c2, s.q_star)));
  const float Jb = -__fmul_rn(s.ustar, bstar);
  const float Jp = Jb > 0.0f ? Jb : 0.0f;
  const float Jh = __fmul_rn(Jp, s.h_bl);   // cbrt: correctly rounded Float32 (evaluated in Float64, rounded once; see ne_common.cuh)
  const float ug = __fmul_rn(Q.beta, Jh > 0.0f ? (float)fm::cbrt_pos(T.mc, (double)Jh) : 0.0f);
  const float UG = Q.gmin < ug ? ug : Q.gmin;
  const float U = sqrtf(__fadd_rn(s.dudv2, __fmul_rn(UG, UG)));
  const float u2 = __fmul_rn(s.ustar, s.ustar);
  const double ud = (double)s.ustar;
  const double ru = fm::rcp(ud);
  const float lW = __fdiv_rn(__fmul_rn(Q.Cg, u2), Q.g_rough);
  const double lu = fm::dmin(fm::fma_(P.a2, ru, (double)lW), P.lmax);
  // L★ = u★²/(ϰ b★) in Float32 (Inf when b★ == 0 or the quotient overflows: 1/L★ = 0)
  const float Lstar = __fdiv_rn(u2, __fmul_rn(Q.kappa, bstar));
  const float aL = fabsf(Lstar);
  const double Linv = (aL < 3.0e38f) ? ((aL > 1.0e-30f) ? fm::rcp((double)Lstar) : 1.0 / (double)Lstar) : 0.0;
  double chi_u, chi_s;
  tab_core<false>(P, T, tab, ud, ru, lu, Linv, (double)s.hd, s.log_hd, chi_u, chi_s, ff);
  s.ustar = (float)fm::mul_(chi_u, (double)U);
  s.theta_star = (float)fm::mul_(chi_s, (double)s.dtheta);
  s.q_star = (float)fm::mul_(chi_s, (double)s.dq);
}

// compute_interface_state.jl:10-18: the first trip always runs; then until drift < tol or it ≥ maxiter
// (FixedIterations: exactly maxiter trips)
template <bool EXT = false>
__device__ __forceinline__ int tab_solve(const FastParams& P, const TabParams& T, const double* tab, FastPoint& s,
                                         const NeFluxFormulation* ff = nullptr) {
  if (P.fixed && P.maxiter <= 0) return 0;
  const double tol = P.fixed ? -1.0 : P.tol;   // drift ≥ 0 > −1: never "converged" under FixedIterations
  const int maxiter = P.maxiter;
  int it = 0;
  double drift;
  do {
    const double pu = s.ustar, pt = s.theta_star, pq = s.q_star;
    tab_iteration<EXT>(P, T, tab, s, ff);
    drift = fabs(s.ustar - pu) + fabs(s.theta_star - pt) + fabs(s.q_star - pq);
    ++it;
  } while (!(drift < tol) && it < maxiter);
  return it;
}
#endif  // __CUDACC__

}  // namespace ne
