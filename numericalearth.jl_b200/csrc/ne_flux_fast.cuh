// ne_flux_fast.cuh — the specialised Float64 solve for the reference's DEFAULT plugin tree
// (ComponentInterfaces defaults, component_interfaces.jl:390-419):
//   SimilarityTheoryFluxes{Float64}: Edson momentum/scalar ψ (any parameter values), logarithmic
//   profile, MomentumRoughnessLength (constant wave parameter, constant viscosity),
//   ScalarRoughnessLength + ReynoldsScalingFunction shared by θ and q, ConvectiveGustiness,
//   ConvergenceStopCriteria or FixedIterations, BulkTemperature, any x_H2O, Float32 or Float64 thermodynamics.
// Any other tree runs through the generic InterfaceSolver of ne_flux_kernels.cu.
//
// What is different from the as-written iteration (results agree to ~1e-15 relative, far inside
// the 1e-10 bar; tests compare against the as-written oracle):
//   * all BulkTemperature invariants hoisted (qₛ, Δq, θₐ, Δθ, g/𝒯ₛ, 1+δqₛ, δ𝒯ₛ, Δu²+Δv², log Δh);
//   * ℓθ ≡ ℓq and ψθ ≡ ψq evaluated once;
//   * logs shared: log(Δh/ℓ) = log Δh − log ℓ with log ℓs = log A − b·log R★, and
//     A/R★^b = exp(log A − b·log R★): the pow and two logs of the as-written code disappear;
//   * 1/L★ formed once; only the ψ branch selected by sign(ζ) is evaluated;
//   * B⁻ = 2 (the only value the reference ships) merges two logs of the unstable momentum ψ.
#pragma once

#include <cstdlib>
#include <vector>

#include "ne_physics.cuh"

#define NE_FAST_PSI_DEG 12

namespace ne {

struct FastParams {
  // Edson momentum
  double m_zmax, m_Ap, m_Bp, m_Cp, m_Dp, m_CpDp, m_Am, m_Bm, m_Cm, m_Dm, m_Em, m_Fm, m_rEm, m_irEm, m_halfEm, m_iEm;
  // Edson scalar
  double s_zmax, s_Ap, s_Bp, s_Cp, s_Dp, s_Ep, s_Am, s_Bm, s_Cm, s_Dm, s_Em, s_Fm, s_rEm, s_irEm, s_halfEm, s_iEm, s_iBm;
  int32_t m_B2, s_C15;  // B⁻ == 2 ; C⁺ == 3/2
  // roughness
  double a1, a2, lmax, nu_inv, rA, log_rA, rb, ls_max, log_ls_max;
  // gustiness
  double beta, gmin;
  double kappa, d_zero, g;
  double tol;
  int32_t maxiter, fixed;
  // options of the table kernels beyond the strict default tree (all off for it)
  int32_t wind_waves;       // WindDependentWaveFormulation: 𝒞g = max(0, C1 min(U, Umax) + C2) (roughness_lengths.jl:75)
  double wave_C1, wave_C2, wave_Umax, inv_g_rough;
  double sgs_const2;        // SubgridVelocityCorrection: constant mesoscale term, already squared (:88-98)
  int32_t const_rough;      // constant roughness lengths (the land default, component_interfaces.jl:514-521): ℓu, ℓθ = ℓq fixed
  double lu_c, ls_c, log_lu_c, log_ls_c;
  // small-|ζ| unstable branch: ψ(ζ) ≈ Σ c_k (ζ/ζs)^k on [-ζs, 0], fitted on the host at Chebyshev nodes in
  // long double from the closed forms (max abs error ≲ 2e-16, checked at fit time; zsmall = 0 disables it)
  double zsmall, zsmall_inv;
  double pm[NE_FAST_PSI_DEG + 1], ps[NE_FAST_PSI_DEG + 1];
};

inline bool fast_path_eligible(const NeFluxFormulation& f, const NeInterfaceProperties& ip, const NeThermoParams& th) {
  if (f.kind != NE_FLUX_SIMILARITY_THEORY || f.similarity_form != NE_PROFILE_LOGARITHMIC) return false;
  (void)th;  // the thermodynamics element type only enters the hoisted prologue/epilogue (kernel template CT)
  if (ip.temperature_formulation != NE_TEMP_BULK) return false;
  if (f.psi_momentum.split || f.psi_momentum.a.kind != NE_PSI_EDSON_MOMENTUM) return false;
  if (f.psi_temperature.split || f.psi_temperature.a.kind != NE_PSI_EDSON_SCALAR) return false;
  if (std::memcmp(&f.psi_temperature, &f.psi_water_vapor, sizeof(NeStabilityProfile)) != 0) return false;
  if (std::memcmp(&f.ell_temperature, &f.ell_water_vapor, sizeof(NeRoughnessLength)) != 0) return false;
  const NeRoughnessLength& m = f.ell_momentum;
  const NeRoughnessLength& s = f.ell_temperature;
  if (m.kind != NE_ROUGH_MOMENTUM || m.wave_kind != NE_WAVE_CONSTANT || m.visc_kind != NE_VISC_CONSTANT) return false;
  if (s.kind != NE_ROUGH_SCALAR || s.visc_kind != NE_VISC_CONSTANT || s.nu != m.nu) return false;
  if (!(m.nu > 0) || !(s.reynolds_A > 0) || !(s.maximum_roughness_length > 0) || !(m.maximum_roughness_length > 0)) return false;
  if (!(m.smooth_wall_parameter > 0) && !(m.wave_constant > 0)) return false;
  const NeSubgridVelocity& g = f.subgrid_velocities;
  if (g.composite || g.convective_kind != NE_SGS_CONVECTIVE || !(g.minimum_gustiness > 0)) return false;
  const double* pm = f.psi_momentum.a.p;
  const double* ps = f.psi_temperature.a.p;
  if (!(pm[9] > 0) || !(ps[10] > 0) || !(ps[7] > 0) || !(pm[6] > 0)) return false;  // E⁻, B⁻ > 0
  return true;
}

// ---- host: closed forms in long double and the Chebyshev fit of the small-|ζ| unstable branch -------------
// f32: √E⁻ is a Float32 sqrt in the Float32 model (psi_edson_momentum<float, double>: FT rE = m_sqrt(Em))
inline long double psi_m_unstable_ld(const double* p, long double z, bool f32 = false) {
  const long double Am = p[5], Bm = p[6], Cm = p[7], Dm = p[8], Em = p[9], Fm = p[10];
  const long double rE = f32 ? (long double)std::sqrt((float)p[9]) : sqrtl(Em);
  const long double f1 = sqrtl(sqrtl(1 - Am * z));
  const long double psi1 = Bm * logl((1 + f1) / Bm) + logl((1 + f1 * f1) / Bm) - Bm * atanl(f1) + Cm;
  const long double f2 = cbrtl(1 - Dm * z);
  const long double psi2 = Em / 2 * logl((1 + f2 + f2 * f2) / Em) - rE * atanl((1 + 2 * f2) / rE) + Fm;
  const long double fw = z * z / (1 + z * z);
  return (1 - fw) * psi1 + fw * psi2;
}
inline long double psi_s_unstable_ld(const double* p, long double z, bool f32 = false) {
  const long double Am = p[6], Bm = p[7], Cm = p[8], Dm = p[9], Em = p[10], Fm = p[11];
  const long double rE = f32 ? (long double)std::sqrt((float)p[10]) : sqrtl(Em);
  const long double f1 = sqrtl(1 - Am * z);
  const long double psi1 = Bm * logl((1 + f1) / Bm) + Cm;
  const long double f2 = cbrtl(1 - Dm * z);
  const long double psi2 = Em / 2 * logl((1 + f2 + f2 * f2) / Em) - rE * atanl((1 + 2 * f2) / rE) + Fm;
  const long double fw = z * z / (1 + z * z);
  return (1 - fw) * psi1 + fw * psi2;
}

// Interpolate f(w·zs), w ∈ [-1, 0], at the N = DEG+1 Chebyshev nodes and convert to monomials in w (long double).
// Returns the max abs error of the double-precision Horner evaluation against the long-double closed form.
template <class F>
inline double fit_small_zeta(F f, double zs, double* coef) {
  constexpr int N = NE_FAST_PSI_DEG + 1;
  const long double PI = 3.141592653589793238462643383279502884L;
  long double y[N], c[N];
  for (int k = 0; k < N; ++k) {
    const long double x = cosl(PI * (2 * k + 1) / (2 * N));
    y[k] = f((x - 1) / 2 * (long double)zs);
  }
  for (int j = 0; j < N; ++j) {
    long double a = 0;
    for (int k = 0; k < N; ++k) a += y[k] * cosl(PI * j * (2 * k + 1) / (2 * N));
    c[j] = a * 2 / N;
  }
  c[0] /= 2;
  // Chebyshev -> monomial in x
  long double T0[N] = {0}, T1[N] = {0}, T2[N], px[N] = {0};
  T0[0] = 1; T1[1] = 1;
  px[0] += c[0];
  for (int m = 0; m < N; ++m) px[m] += c[1] * T1[m];
  for (int j = 2; j < N; ++j) {
    for (int m = 0; m < N; ++m) T2[m] = (m > 0 ? 2 * T1[m - 1] : 0) - T0[m];
    for (int m = 0; m < N; ++m) { px[m] += c[j] * T2[m]; T0[m] = T1[m]; T1[m] = T2[m]; }
  }
  // x = 2w + 1
  long double pw[N] = {0};
  for (int m = 0; m < N; ++m) {
    long double binom = 1;   // C(m, r)
    for (int r = 0; r <= m; ++r) {
      pw[r] += px[m] * binom * powl(2.0L, r);
      binom = binom * (m - r) / (r + 1);
    }
  }
  for (int m = 0; m < N; ++m) coef[m] = (double)pw[m];
  double err = 0;
  for (int i = 0; i <= 4000; ++i) {
    const double w = -((double)i / 4000.0) * ((double)i / 4000.0);   // denser near 0
    double r = coef[N - 1];
    for (int m = N - 2; m >= 0; --m) r = r * w + coef[m];
    const long double t = f((long double)w * (long double)zs);
    const double e = (double)fabsl((long double)r - t);
    if (e > err) err = e;
  }
  return err;
}

// f32: parameters of the Float32 model's mixed-precision iteration (ne_flux_tab.cuh): every plugin parameter is
// the Float32-rounded value promoted back to Float64, derived constants follow stability_fn<float, double>
// (√E⁻ and C⁺D⁺ are Float32 operations); the air viscosity stays the Float64 literal.
inline FastParams make_fast_params(const NeFluxFormulation& f, double g, bool f32 = false) {
  FastParams P;
  auto R = [f32](double x) { return f32 ? (double)(float)x : x; };
  const double* p = f.psi_momentum.a.p;
  P.m_zmax = R(p[0]); P.m_Ap = R(p[1]); P.m_Bp = R(p[2]); P.m_Cp = R(p[3]); P.m_Dp = R(p[4]);
  P.m_CpDp = f32 ? (double)((float)p[3] * (float)p[4]) : p[3] * p[4];
  P.m_Am = R(p[5]); P.m_Bm = R(p[6]); P.m_Cm = R(p[7]); P.m_Dm = R(p[8]); P.m_Em = R(p[9]); P.m_Fm = R(p[10]);
  P.m_rEm = f32 ? (double)std::sqrt((float)p[9]) : std::sqrt(p[9]);
  P.m_irEm = 1.0 / P.m_rEm; P.m_halfEm = P.m_Em / 2; P.m_iEm = 1.0 / P.m_Em;
  P.m_B2 = (P.m_Bm == 2.0);
  const double* q = f.psi_temperature.a.p;
  P.s_zmax = R(q[0]); P.s_Ap = R(q[1]); P.s_Bp = R(q[2]); P.s_Cp = R(q[3]); P.s_Dp = R(q[4]); P.s_Ep = R(q[5]);
  P.s_Am = R(q[6]); P.s_Bm = R(q[7]); P.s_Cm = R(q[8]); P.s_Dm = R(q[9]); P.s_Em = R(q[10]); P.s_Fm = R(q[11]);
  P.s_rEm = f32 ? (double)std::sqrt((float)q[10]) : std::sqrt(q[10]);
  P.s_irEm = 1.0 / P.s_rEm; P.s_halfEm = P.s_Em / 2; P.s_iEm = 1.0 / P.s_Em;
  P.s_iBm = 1.0 / P.s_Bm;
  P.s_C15 = (P.s_Cp == 1.5);
  const NeRoughnessLength& m = f.ell_momentum;
  const NeRoughnessLength& s = f.ell_temperature;
  P.a1 = m.wave_constant / m.gravitational_acceleration;   // Float64 model only (the Float32 front forms ℓ_W itself)
  P.a2 = R(m.smooth_wall_parameter) * m.nu;
  P.lmax = R(m.maximum_roughness_length);
  P.nu_inv = 1.0 / s.nu;
  P.rA = R(s.reynolds_A); P.log_rA = std::log(P.rA); P.rb = R(s.reynolds_b);
  P.ls_max = R(s.maximum_roughness_length); P.log_ls_max = std::log(P.ls_max);
  P.beta = R(f.subgrid_velocities.gustiness_parameter); P.gmin = R(f.subgrid_velocities.minimum_gustiness);
  P.kappa = R(f.von_karman_constant); P.d_zero = R(f.zero_plane_displacement); P.g = R(g);
  P.tol = R(f.stop.tolerance); P.maxiter = f.stop.maxiter; P.fixed = f.stop.kind == NE_STOP_FIXED_ITERATIONS;
  P.const_rough = m.kind == NE_ROUGH_CONSTANT && s.kind == NE_ROUGH_CONSTANT;
  P.lu_c = R(m.constant); P.ls_c = R(s.constant);
  P.log_lu_c = P.const_rough ? std::log(P.lu_c) : 0.0;
  P.log_ls_c = P.const_rough ? std::log(P.ls_c) : 0.0;
  if (P.const_rough) {   // the roughness-model constants are not used: keep them finite
    // … and such that tab_core's Reynolds-scaling branch always resolves to the clipped value ℓs = ls_c
    P.a1 = 0; P.a2 = 0; P.lmax = P.lu_c; P.nu_inv = 1; P.rA = 1; P.log_rA = 700.0; P.rb = 0; P.ls_max = P.ls_c; P.log_ls_max = P.log_ls_c;
  }
  P.wind_waves = m.wave_kind == NE_WAVE_WIND_DEPENDENT;
  P.wave_C1 = R(m.wave_C1); P.wave_C2 = R(m.wave_C2); P.wave_Umax = R(m.wave_Umax);
  P.inv_g_rough = 1.0 / R(m.gravitational_acceleration);
  {
    const NeSubgridVelocity& sg = f.subgrid_velocities;
    const double c = (sg.composite && sg.mesoscale_kind == NE_SGS_CONSTANT) ? R(sg.mesoscale_constant) : 0.0;
    P.sgs_const2 = c * c;
  }
  // the small-|ζ| polynomial belongs to the closed-form kernel only: see add_small_zeta_poly
  P.zsmall = 0; P.zsmall_inv = 0;
  for (int k = 0; k <= NE_FAST_PSI_DEG; ++k) { P.pm[k] = 0; P.ps[k] = 0; }
  return P;
}

// Small-|ζ| polynomials of the closed-form kernel (ao_flux_fast_kernel; ψ(ℓ/L★) always lands there: |ℓ/L★| ≲ 1e-3),
// fitted once per Edson parameter set and kept in a small per-thread cache (the table-driven kernels do not use them).
inline void add_small_zeta_poly(FastParams& P, const NeFluxFormulation& f) {
  struct Entry { double key[24]; double val[2 * (NE_FAST_PSI_DEG + 1) + 1]; };
  static thread_local std::vector<Entry> cache;
  const char* off = std::getenv("NE_B200_NO_SMALL_ZETA_POLY");
  if (off && off[0] == '1') return;
  const double* p = f.psi_momentum.a.p;
  const double* q = f.psi_temperature.a.p;
  double key[24];
  for (int k = 0; k < 12; ++k) { key[k] = p[k]; key[12 + k] = q[k]; }
  const Entry* hit = nullptr;
  for (const Entry& e : cache)
    if (std::memcmp(key, e.key, sizeof(key)) == 0) { hit = &e; break; }
  if (!hit) {
    Entry e;
    std::memcpy(e.key, key, sizeof(key));
    double zs = 1.0 / 64;
    while (zs > 1e-6) {
      const double em = fit_small_zeta([&](long double z) { return psi_m_unstable_ld(p, z); }, zs, e.val);
      const double es = fit_small_zeta([&](long double z) { return psi_s_unstable_ld(q, z); }, zs, e.val + NE_FAST_PSI_DEG + 1);
      if (em <= 4e-16 && es <= 4e-16) break;
      zs *= 0.5;
    }
    if (!(zs > 1e-6)) zs = 0;
    e.val[2 * (NE_FAST_PSI_DEG + 1)] = zs;
    if (cache.size() >= 16) cache.clear();
    cache.push_back(e);
    hit = &cache.back();
  }
  for (int k = 0; k <= NE_FAST_PSI_DEG; ++k) { P.pm[k] = hit->val[k]; P.ps[k] = hit->val[NE_FAST_PSI_DEG + 1 + k]; }
  P.zsmall = hit->val[2 * (NE_FAST_PSI_DEG + 1)];
  P.zsmall_inv = P.zsmall > 0 ? 1.0 / P.zsmall : 0.0;
}

// ψ_m(ζ), Edson et al. (2013) — similarity_theory_turbulent_fluxes.jl:501-532
__device__ __forceinline__ double fast_psi_poly(const double* c, double w) {
  double r = c[NE_FAST_PSI_DEG];
#pragma unroll
  for (int m = NE_FAST_PSI_DEG - 1; m >= 0; --m) r = fma(r, w, c[m]);
  return r;
}

__device__ __forceinline__ double fast_psi_m(const FastParams& P, double z) {
  if (z < 0) {
    if (z >= -P.zsmall) return fast_psi_poly(P.pm, z * P.zsmall_inv);
    const double f1 = sqrt(sqrt(1.0 - P.m_Am * z));
    const double f1s = f1 * f1;
    double psi1;
    if (P.m_B2) psi1 = log((1.0 + f1) * (1.0 + f1) * (1.0 + f1s) * 0.125) - 2.0 * atan(f1) + P.m_Cm;
    else psi1 = P.m_Bm * log((1.0 + f1) / P.m_Bm) + log((1.0 + f1s) / P.m_Bm) - P.m_Bm * atan(f1) + P.m_Cm;
    const double f2 = cbrt(1.0 - P.m_Dm * z);
    const double psi2 = P.m_halfEm * log((1.0 + f2 + f2 * f2) * P.m_iEm) - P.m_rEm * atan((1.0 + 2.0 * f2) * P.m_irEm) + P.m_Fm;
    const double z2 = z * z;
    const double fw = z2 / (1.0 + z2);
    return psi1 + fw * (psi2 - psi1);
  }
  const double dz = fmin(P.m_zmax, P.m_Ap * z);
  return -P.m_Bp * z - P.m_Cp * (z - P.m_Dp) * exp(-dz) - P.m_CpDp;
}

// ψ_s(ζ) — :586-618
__device__ __forceinline__ double fast_psi_s(const FastParams& P, double z) {
  if (z < 0) {
    if (z >= -P.zsmall) return fast_psi_poly(P.ps, z * P.zsmall_inv);
    const double f1 = sqrt(1.0 - P.s_Am * z);
    const double psi1 = P.s_Bm * log((1.0 + f1) * P.s_iBm) + P.s_Cm;
    const double f2 = cbrt(1.0 - P.s_Dm * z);
    const double psi2 = P.s_halfEm * log((1.0 + f2 + f2 * f2) * P.s_iEm) - P.s_rEm * atan((1.0 + 2.0 * f2) * P.s_irEm) + P.s_Fm;
    const double z2 = z * z;
    const double fw = z2 / (1.0 + z2);
    return psi1 + fw * (psi2 - psi1);
  }
  const double dz = fmin(P.s_zmax, P.s_Ap * z);
  const double x = 1.0 + P.s_Bp * z;
  const double xp = P.s_C15 ? x * sqrt(x) : pow(x, P.s_Cp);
  return -xp - P.s_Bp * (z - P.s_Dp) * exp(-dz) - P.s_Ep;
}

struct FastPoint {
  // inputs
  double gTv, c1, c2;        // g/𝒯ₛ, 1+δqₛ, δ𝒯ₛ
  double dudv2, h_bl, hd, log_hd;
  double dtheta, dq;
  // iterate
  double ustar, theta_star, q_star;
};

__device__ __forceinline__ void fast_iteration(const FastParams& P, FastPoint& s) {
  const double bstar = s.gTv * (s.theta_star * s.c1 + s.c2 * s.q_star);
  const double Jb = -s.ustar * bstar;
  const double UG = fmax(P.gmin, P.beta * cbrt(fmax(0.0, Jb) * s.h_bl));
  const double U = sqrt(s.dudv2 + UG * UG);
  const double ru = 1.0 / s.ustar;
  const double lu = fmin(P.a1 * s.ustar * s.ustar + P.a2 * ru, P.lmax);
  const double log_lu = log(lu);
  const double log_Rs = log(lu * s.ustar * P.nu_inv);
  const double log_ls_un = P.log_rA - P.rb * log_Rs;
  const bool clipped = log_ls_un > P.log_ls_max;
  const double log_ls = clipped ? P.log_ls_max : log_ls_un;
  const double ls = clipped ? P.ls_max : exp(log_ls_un);
  const bool lifted = 2.0 * lu > s.hd;
  const double dh = lifted ? 2.0 * lu : s.hd;
  const double log_dh = lifted ? 0.6931471805599453 + log_lu : s.log_hd;
  const double Linv = P.kappa * bstar * ru * ru;   // 1/L★ (0 when b★ == 0, i.e. L★ = Inf)
  const double zh = dh * Linv;
  const double Pi_u = (log_dh - log_lu) - fast_psi_m(P, zh) + fast_psi_m(P, lu * Linv);
  const double Pi_s = (log_dh - log_ls) - fast_psi_s(P, zh) + fast_psi_s(P, ls * Linv);
  const double chi_s = P.kappa / Pi_s;
  s.ustar = P.kappa / Pi_u * U;
  s.theta_star = chi_s * s.dtheta;
  s.q_star = chi_s * s.dq;
}

// compute_interface_state for the default tree; returns the iteration count
__device__ __forceinline__ int fast_solve(const FastParams& P, FastPoint& s) {
  int it = 0;
  double drift = 0;
  for (;;) {
    const bool go = P.fixed ? (it < P.maxiter) : (!((drift < P.tol) | (it >= P.maxiter)) | (it == 0));
    if (!go) break;
    const double pu = s.ustar, pt = s.theta_star, pq = s.q_star;
    fast_iteration(P, s);
    drift = fabs(s.ustar - pu) + fabs(s.theta_star - pt) + fabs(s.q_star - pq);
    ++it;
  }
  return it;
}

}  // namespace ne
