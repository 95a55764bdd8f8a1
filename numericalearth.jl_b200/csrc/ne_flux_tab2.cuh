// ne_flux_tab2.cuh — round-2 form of the table-driven a–o solve (default plugin tree, Float64 model).
//
// Same fixed point, stopping rule and trip counts as ne_flux_tab.cuh (compute_interface_state.jl:5-58,
// similarity_theory_turbulent_fluxes.jl:315-385); what changes is the SHAPE of the code the hardware sees
// (profiles/r01_ncu_full_v10_summary.txt: FP64 pipe 55 %, issue 62 %, shared-memory wavefronts 61 %, lane
// efficiency 25.6/32, `wait` the largest stall):
//
//   * one straight-line basic block per trip: the ψ(Δh/L★) pair and the |ζ| < 2^-12 ψ(ℓ/L★) pair are ALWAYS
//     evaluated from (clamped) table records and the rare cases (|ζ| ≥ 2^7, ℓ/L★ outside the micro records)
//     overwrite the result afterwards — the scheduler can interleave the cbrt → sqrt chain of the gustiness,
//     the two logs, the exp and the four Horner chains instead of running them block after block;
//   * every FP64 operation goes through an ops policy (ne_fastmath.cuh): the OpsCount instantiation of the same
//     source counts the FP64 instructions a launch executes (bench.py's roofline numerator, no profiler);
//   * prologue and epilogue (once per point) use the branch-free reciprocal / log / exp instead of IEEE division
//     and libdevice powf / expf (~650 → ~350 instructions per point); a Float32 q_sat (Float32 thermodynamics:
//     interface_states.jl:56-59) is evaluated as Float32(pow_Float64) · Float32(exp_Float64), i.e. correctly
//     rounded Float32 functions, which is what Julia's Float32 `^` is (Base.Math.pow_body widens to Float64)
//     and what every other kernel and the oracle do (ne_common.cuh m_pow / m_exp);
//   * points are dealt to warps in the order of the PREVIOUS step's trip counts (ne_flux_tab2.cu:
//     trip_order_kernel), so the 32 lanes of a warp leave the loop together.
#pragma once

#include "ne_flux_tab.cuh"

namespace ne {

#if defined(__CUDACC__)

namespace fm {
template <class O> __device__ __forceinline__ double atan_large(O& o, double x) {   // x ≥ 6: |error| ≲ 1 ulp of π/2
  const double y = rcp(o, x), y2 = o.mul(y, y);
  double p = -1.0 / 19.0;
  p = o.fma(p, y2, 1.0 / 17.0);
  p = o.fma(p, y2, -1.0 / 15.0);
  p = o.fma(p, y2, 1.0 / 13.0);
  p = o.fma(p, y2, -1.0 / 11.0);
  p = o.fma(p, y2, 1.0 / 9.0);
  p = o.fma(p, y2, -1.0 / 7.0);
  p = o.fma(p, y2, 1.0 / 5.0);
  p = o.fma(p, y2, -1.0 / 3.0);
  p = o.fma(p, y2, 1.0);
  return o.fma(-y, p, 1.5707963267948966);
}
}  // namespace fm

// ---- rare ψ paths -----------------------------------------------------------------------------------------
// `which`: 0 = ψ_m only, 1 = ψ_s only, 2 = both (the half that is not asked for is not evaluated)
// Edson unstable closed forms for ζ ≤ −2^7 (similarity_theory_turbulent_fluxes.jl:501-532, 586-618), B⁻ = 2
template <class O>
__device__ __forceinline__ void tab2_psi_far_unstable(O& o, const FastParams& P, const TabParams& T, const double* tab,
                                                      double z, int which, double& pm, double& ps) {
  const double z2 = o.mul(z, z);
  const double fw = o.sub(1.0, fm::rcp(o, o.add(1.0, z2)));                      // ζ²/(1 + ζ²)
  if (which != 1) {  // momentum
    const double f1 = fm::sqrt_pos(o, fm::sqrt_pos(o, o.fma(-P.m_Am, z, 1.0)));
    const double f1s = o.mul(f1, f1), f1p = o.add(1.0, f1);
    const double arg = o.mul(o.mul(o.mul(f1p, f1p), o.add(1.0, f1s)), 0.125);
    const double psi1 = o.add(o.fma(-2.0, fm::atan_large(o, f1), fm::log_pos(o, tab, T.mc, arg)), P.m_Cm);
    const double f2 = fm::cbrt_pos(o, T.mc, o.fma(-P.m_Dm, z, 1.0));
    const double l2 = fm::log_pos(o, tab, T.mc, o.mul(o.fma(f2, f2, o.add(1.0, f2)), P.m_iEm));
    const double a2 = fm::atan_large(o, o.mul(o.fma(2.0, f2, 1.0), P.m_irEm));
    const double psi2 = o.add(o.fma(-P.m_rEm, a2, o.mul(P.m_halfEm, l2)), P.m_Fm);
    pm = o.fma(fw, o.sub(psi2, psi1), psi1);
  }
  if (which != 0) {  // scalar
    const double f1 = fm::sqrt_pos(o, o.fma(-P.s_Am, z, 1.0));
    const double psi1 = o.fma(P.s_Bm, fm::log_pos(o, tab, T.mc, o.mul(o.add(1.0, f1), P.s_iBm)), P.s_Cm);
    const double f2 = fm::cbrt_pos(o, T.mc, o.fma(-P.s_Dm, z, 1.0));
    const double l2 = fm::log_pos(o, tab, T.mc, o.mul(o.fma(f2, f2, o.add(1.0, f2)), P.s_iEm));
    const double a2 = fm::atan_large(o, o.mul(o.fma(2.0, f2, 1.0), P.s_irEm));
    const double psi2 = o.add(o.fma(-P.s_rEm, a2, o.mul(P.s_halfEm, l2)), P.s_Fm);
    ps = o.fma(fw, o.sub(psi2, psi1), psi1);
  }
}

// stable closed forms (ζ ≥ 2^7: the tables cover the rest) with the branch-free exp / sqrt (:519-531, 605-617).  Every
// solve starts from u★ = θ★ = q★ = 1e-4 (atmosphere_ocean_fluxes.jl:131-137), i.e. from b★ > 0 and ζ ~ 1e5: its first
// trips evaluate ψ(Δh/L★), ψ(ℓu/L★) and ψ(ℓs/L★) HERE, where both exponentials have long saturated at
// exp(−ζmax) — above ζ_sat = max(ζmax/A⁺) they are the two constants T.em_sat, T.es_sat (the same bits exp_mid returns)
template <class O>
__device__ __forceinline__ void tab2_psi_stable(O& o, const FastParams& P, const TabParams& T, double z, int which, double& pm, double& ps) {
  double em, es;
  if (z >= T.z_sat) { em = T.em_sat; es = T.es_sat; }
  else {
    em = fm::exp_mid(o, T.mc, -fm::dmin(P.m_zmax, o.mul(P.m_Ap, z)));
    es = T.same_exp ? em : fm::exp_mid(o, T.mc, -fm::dmin(P.s_zmax, o.mul(P.s_Ap, z)));
  }
  if (which != 1) pm = o.sub(o.fma(-o.mul(P.m_Cp, o.sub(z, P.m_Dp)), em, -o.mul(P.m_Bp, z)), P.m_CpDp);
  if (which != 0) {
    const double x = o.fma(P.s_Bp, z, 1.0);
    double xp;
    if (P.s_C15) xp = o.mul(x, fm::sqrt_pos(o, x));
    else { xp = pow_general(x, P.s_Cp); o.other(60); }
    ps = o.sub(o.fma(-o.mul(P.s_Bp, o.sub(z, P.s_Dp)), es, -xp), P.s_Ep);
  }
}

// Out of line, own policy object (an ops reference would force the caller's counters into local memory); the
// counting instantiation adds its operations to the launch's global counters itself.
template <class O>
static __device__ __noinline__ double2 tab2_psi_outside(const FastParams& P, const TabParams& T, const double* tab, double z,
                                                        int which, unsigned long long* counts) {
  O o;
  double pm = 0, ps = 0;
  if (z > 0) tab2_psi_stable(o, P, T, z, which, pm, ps);
  else if (T.far_fm) tab2_psi_far_unstable(o, P, T, tab, z, which, pm, ps);
  else { psi_far_unstable(P, z, pm, ps); o.other(700); }
  o.flush(counts);
  return make_double2(pm, ps);
}

// ψ_m(ζ_u), ψ_s(ζ_s) when ℓ/L★ leaves the micro records (the first trips, while u★ is still near its 1e-4 guess)
template <class O>
static __device__ __noinline__ double2 tab2_psi_small(const FastParams& P, const TabParams& T, const double* tab, double zu,
                                                      double zs, double Linv, unsigned long long* counts) {
  O o;
  double pm, ps;
  if (fm::psi_is_tiny(zu) && fm::psi_is_tiny(zs)) {
    fm::psi_tiny_pair(o, tab + fm::TAB_TINY + (Linv < 0 ? 0 : fm::TINY_REC), fabs(zu), fabs(zs), pm, ps);
  } else {
    bool outside;
    int iv = fm::psi_interval(zu, outside);
    if (!outside) pm = fm::psi_single(o, tab + fm::TAB_PSI + iv * fm::PSI_REC, fabs(zu), 0);
    else pm = tab2_psi_outside<O>(P, T, tab, zu, 0, counts).x;
    iv = fm::psi_interval(zs, outside);
    if (!outside) ps = fm::psi_single(o, tab + fm::TAB_PSI + iv * fm::PSI_REC, fabs(zs), 1);
    else ps = tab2_psi_outside<O>(P, T, tab, zs, 1, counts).y;
  }
  o.flush(counts);
  return make_double2(pm, ps);
}

// ---- one trip of iterate_interface_state (compute_interface_state.jl:69-122 + similarity_theory…:315-385) ---
// the two |ζ| < 2^-12 records (unstable, stable) as a kernel parameter: constant-bank operands
struct Micro { double rec[2][fm::MICRO_REC]; };

// What a thread keeps in registers across the loop: u★, χ_s = ϰ/Π_s and b★.  With BulkTemperature θ★ = χ_s Δθ and
// q★ = χ_s Δq are one number times two per-point invariants, so
//     b★ = g/𝒯ₛ (θ★ (1 + δqₛ) + δ𝒯ₛ q★) = χ_s C,          C = A Δθ + B Δq,  A = g/𝒯ₛ (1 + δqₛ), B = g/𝒯ₛ δ𝒯ₛ
//     |θ★ − θ★'| + |q★ − q★'| = |χ_s − χ_s'| D,            D = |Δθ| + |Δq|
// (identities in real arithmetic; in Float64 they move b★ and the drift by a few ulp — the drift is compared with 1e-8,
// the iterate itself is unchanged: θ★ and q★ are formed from the final χ_s with the reference's own multiplication).
// C, D and Δu² + Δv² wait in the thread's own column of shared memory (conflict-free 64-bit accesses) and are read where
// a trip needs them: 6 registers of loop-carried state instead of 24 (the 80-register build of round 1 reloaded three
// spilled doubles at the head of every trip), 3 shared-memory loads per trip instead of 5.  The first trip starts from
// θ★ = q★ = 1e-4 (atmosphere_ocean_fluxes.jl:131-137), which is not of that form: its b★ and its drift are formed as written.
// Slots of the column (stride NT doubles):
enum { SL_DU, SL_DV, SL_TA, SL_PA, SL_QA, SL_TS, SL_DTH, SL_DQ, SL_C, SL_D, SL_DUDV2, SL_COUNT };
struct Tab2Point { double ustar, chi_s, bstar; };
struct Tab2Heights { double h_bl, hd, log_hd; };

// a shared-memory read the compiler may not hoist out of the loop (hoisting is what spills)
__device__ __forceinline__ double slot_ld(const double* p) { return *(const volatile double*)p; }

// `lrep`: the log table replicated per lane class (fm::log_pos_rep), `lc` = lane mod 8
// `rare`: copies of P and T in shared memory for the out-of-line closed forms (a reference to a kernel parameter handed to a
// non-inlined function turns every use into a generic load from the parameter window: long-scoreboard stalls, ncu)
struct Tab2Rare { FastParams P; TabParams T; };

template <int NT, class O>
__device__ __forceinline__ void tab2_iteration(O& o, const FastParams& P, const TabParams& T, const Tab2Rare& rare, const Micro& Mi, const double* tab,
                                               const double* lrep, int lc, const Tab2Heights& H, Tab2Point& s, const double* col,
                                               unsigned long long* counts, int& record) {
  using fm::dmax;
  using fm::dmin;
  // b★, gustiness, U (similarity_theory…:354-358, 417-425)
  const double bstar = s.bstar;
  const double Jb = -o.mul(s.ustar, bstar);
  // U_G = max(floor, β ∛(max(0, J_b) h_bl)) is its floor wherever the buoyancy flux is not destabilising: a warp whose lanes
  // are all stable (trip-ordered lanes share the stability regime) skips the cube root — same bits, 25 instructions fewer
  double UG = P.gmin;
  if (__any_sync(__activemask(), Jb > 0.0))
    UG = dmax(P.gmin, o.mul(P.beta, fm::cbrt3(o, T.mc, dmax(o.mul(dmax(0.0, Jb), H.h_bl), T.cbrt_floor))));
  const double U = fm::sqrt3(o, o.fma(UG, UG, slot_ld(col + SL_DUDV2 * NT)));
  // roughness lengths (roughness_lengths.jl:197-246) and 1/L★
  const double ru = fm::rcp3(o, s.ustar);
  const double lu = dmin(o.fma(o.mul(P.a1, s.ustar), s.ustar, o.mul(P.a2, ru)), P.lmax);
  const double Linv = o.mul(o.mul(o.mul(P.kappa, bstar), ru), ru);   // 0 when b★ == 0, i.e. L★ = Inf (:372)
  const double log_lu = fm::log_pos_rep(o, lrep, lc, T.mc, lu);
  const double Rs = o.mul(o.mul(lu, s.ustar), P.nu_inv);
  const double log_Rs = fm::log_pos_rep(o, lrep, lc, T.mc, Rs);
  const double log_ls_un = o.fma(-P.rb, log_Rs, P.log_rA);
  const bool clipped = log_ls_un > P.log_ls_max;
  const double log_ls = clipped ? P.log_ls_max : log_ls_un;
  // ℓs itself only scales the argument of ψ_s(ℓs/L★), |ψ| < 1e-3 against Π_s ~ 10: the short exponential (2e-12) is ample
  const double ls_un = fm::exp_lo(o, tab, T.mc, dmax(log_ls_un, -700.0));
  const double ls = clipped ? P.ls_max : ls_un;
  const double lu2 = o.add(lu, lu);
  const bool lifted = lu2 > H.hd;                                    // Δh = max(Δh − d, 2ℓu) (:313)
  const double dh = lifted ? lu2 : H.hd;
  const double log_dh = lifted ? o.add(T.mc.ln2, log_lu) : H.log_hd;
  // ψ(Δh/L★): always from the (clamped) record; |ζ| ≥ 2^7 overwrites below
  const double zh = o.mul(dh, Linv);
  bool outside;
  const int iv = fm::psi_interval(zh, outside);
  record = iv;
  double pm_h, ps_h;
  fm::psi_pair_bits(o, tab + fm::TAB_PSI + iv * fm::PSI_REC, fabs(zh), pm_h, ps_h);
  // ψ(ℓ/L★): always from the |ζ| < 2^-12 record of the side of L★.  The two records also sit in the kernel-parameter
  // constant bank (Micro): a warp whose lanes are all on one side — nearly every warp, L★ varies smoothly in space —
  // takes its coefficients from there and issues no shared-memory load for this lookup
  const double zu = o.mul(lu, Linv), zs = o.mul(ls, Linv);
  double pm_l, ps_l;
  {
    const bool unstable = Linv < 0;
    const unsigned act = __activemask();
    const unsigned neg = __ballot_sync(act, unstable);
    if (neg == act) fm::psi_micro_pair(o, Mi.rec[0], fabs(zu), fabs(zs), pm_l, ps_l);
    else if (neg == 0u) fm::psi_micro_pair(o, Mi.rec[1], fabs(zu), fabs(zs), pm_l, ps_l);
    else fm::psi_micro_pair(o, tab + fm::TAB_MICRO + (unstable ? 0 : fm::MICRO_REC), fabs(zu), fabs(zs), pm_l, ps_l);
  }
  if (outside) {
    const double2 r = tab2_psi_outside<O>(rare.P, rare.T, tab, zh, 2, counts);
    pm_h = r.x; ps_h = r.y;
  }
  if (!(fm::psi_is_micro(zu) && fm::psi_is_micro(zs))) {
    const double2 r = tab2_psi_small<O>(rare.P, rare.T, tab, zu, zs, Linv, counts);
    pm_l = r.x; ps_l = r.y;
  }
  // Π = log(Δh/ℓ) − ψ(Δh/L★) + ψ(ℓ/L★), χ = ϰ/Π (:242-247, 375-377)
  const double Pi_u = o.add(o.sub(o.sub(log_dh, log_lu), pm_h), pm_l);
  const double Pi_s = o.add(o.sub(o.sub(log_dh, log_ls), ps_h), ps_l);
  const double r = fm::rcp3(o, o.mul(Pi_u, Pi_s));
  const double ru_ = o.mul(Pi_s, r), rs_ = o.mul(Pi_u, r);           // 1/Π_u, 1/Π_s
  double chi_u = o.mul(P.kappa, ru_), chi_s = o.mul(P.kappa, rs_);
  chi_u = o.fma(o.fma(-Pi_u, chi_u, P.kappa), ru_, chi_u);
  chi_s = o.fma(o.fma(-Pi_s, chi_s, P.kappa), rs_, chi_s);
  s.ustar = o.mul(chi_u, U);
  s.chi_s = chi_s;
  s.bstar = o.mul(chi_s, slot_ld(col + SL_C * NT));
}

// ---- the first trip ---------------------------------------------------------------------------------------
// Every solve starts from u★ = θ★ = q★ = 1e-4 (atmosphere_ocean_fluxes.jl:131-137).  With that iterate the roughness
// lengths, their logarithms and Δh are the same numbers at every point (formed once per launch on the host with the
// functions of ne_fastmath.cuh), b★ > 0 (stable: U_G is its floor), and L★ is so small (ζ ~ 1e5) that ψ(Δh/L★) and
// ψ(ℓu/L★) sit on the saturated stable branch, where the Edson functions are a line and a 3/2 power.  The generic trip
// spends ~800 instructions there (three out-of-line closed-form calls); this one does the same arithmetic in ~150.
// Taken only when every lane of the warp qualifies; otherwise the generic trip runs (same result either way to the
// last ulp of the host-formed constants).
struct Tab2First { double ru0, lu0, log_lu0, ls0, log_ls0, dh0, log_dh0; int32_t ok; };

inline Tab2First make_tab2_first(const FastParams& P, const TabParams& T, const double* host_tab, double hd, double log_hd) {
  using fm::dmax; using fm::dmin;
  fm::OpsPlain o;
  Tab2First F;
  const double u0 = 1e-4;
  F.ru0 = fm::rcp(o, u0);
  F.lu0 = dmin(o.fma(o.mul(P.a1, u0), u0, o.mul(P.a2, F.ru0)), P.lmax);
  F.log_lu0 = fm::log_pos(o, host_tab, T.mc, F.lu0);
  const double Rs = o.mul(o.mul(F.lu0, u0), P.nu_inv);
  const double log_ls_un = o.fma(-P.rb, fm::log_pos(o, host_tab, T.mc, Rs), P.log_rA);
  const bool clipped = log_ls_un > P.log_ls_max;
  F.log_ls0 = clipped ? P.log_ls_max : log_ls_un;
  F.ls0 = clipped ? P.ls_max : fm::exp_lo(o, host_tab, T.mc, dmax(log_ls_un, -700.0));
  const double lu2 = o.add(F.lu0, F.lu0);
  const bool lifted = lu2 > hd;
  F.dh0 = lifted ? lu2 : hd;
  F.log_dh0 = lifted ? o.add(T.mc.ln2, F.log_lu0) : log_hd;
  F.ok = (T.z_sat < 1e300 && P.s_C15 && F.lu0 > 0 && F.ls0 > 0 && Rs > 0 && !(P.fixed && P.maxiter <= 0)) ? 1 : 0;
  return F;
}

template <int NT, class O>
__device__ __forceinline__ bool tab2_first_trip(O& o, const FastParams& P, const TabParams& T, const Tab2First& F, const double* tab,
                                                Tab2Point& s, const double* col, double& drift, int& record) {
  const double Linv = o.mul(o.mul(o.mul(P.kappa, s.bstar), F.ru0), F.ru0);
  const double zh = o.mul(F.dh0, Linv), zu = o.mul(F.lu0, Linv), zs = o.mul(F.ls0, Linv);
  bool out_s, out_h;
  const int ivs = fm::psi_interval(zs, out_s);
  const bool mine = F.ok && s.bstar > 0.0 && zh >= T.z_sat && zu >= T.z_sat && !out_s && !fm::psi_is_tiny(zs);
  if (!__all_sync(__activemask(), mine)) return false;
  record = fm::psi_interval(zh, out_h);
  const double U = fm::sqrt3(o, o.fma(P.gmin, P.gmin, slot_ld(col + SL_DUDV2 * NT)));
  // ψ_m, ψ_s on the saturated stable branch (tab2_psi_stable with ζ ≥ ζ_sat: the same operations)
  const double pm_h = o.sub(o.fma(-o.mul(P.m_Cp, o.sub(zh, P.m_Dp)), T.em_sat, -o.mul(P.m_Bp, zh)), P.m_CpDp);
  const double pm_l = o.sub(o.fma(-o.mul(P.m_Cp, o.sub(zu, P.m_Dp)), T.em_sat, -o.mul(P.m_Bp, zu)), P.m_CpDp);
  const double x = o.fma(P.s_Bp, zh, 1.0);
  const double ps_h = o.sub(o.fma(-o.mul(P.s_Bp, o.sub(zh, P.s_Dp)), T.es_sat, -o.mul(x, fm::sqrt3(o, x))), P.s_Ep);
  const double ps_l = fm::psi_single(o, tab + fm::TAB_PSI + ivs * fm::PSI_REC, zs, 1);
  const double Pi_u = o.add(o.sub(o.sub(F.log_dh0, F.log_lu0), pm_h), pm_l);
  const double Pi_s = o.add(o.sub(o.sub(F.log_dh0, F.log_ls0), ps_h), ps_l);
  const double r = fm::rcp3(o, o.mul(Pi_u, Pi_s));
  const double ru_ = o.mul(Pi_s, r), rs_ = o.mul(Pi_u, r);
  double chi_u = o.mul(P.kappa, ru_), chi_s = o.mul(P.kappa, rs_);
  chi_u = o.fma(o.fma(-Pi_u, chi_u, P.kappa), ru_, chi_u);
  chi_s = o.fma(o.fma(-Pi_s, chi_s, P.kappa), rs_, chi_s);
  const double pu = s.ustar;
  s.ustar = o.mul(chi_u, U);
  s.chi_s = chi_s;
  s.bstar = o.mul(chi_s, slot_ld(col + SL_C * NT));
  const double th = o.mul(chi_s, slot_ld(col + SL_DTH * NT)), q = o.mul(chi_s, slot_ld(col + SL_DQ * NT));
  drift = o.add(o.add(fabs(o.sub(s.ustar, pu)), fabs(o.sub(th, 1e-4))), fabs(o.sub(q, 1e-4)));
  return true;
}

// compute_interface_state.jl:10-18: the first trip always runs; then until drift < tol or it ≥ maxiter
// `record`: the ψ(Δh/L★) table record of the last trip (the next step's ordering hint)
template <int NT, class O>
__device__ __forceinline__ int tab2_solve(O& o, const FastParams& P, const TabParams& T, const Tab2Rare& rare, const Micro& Mi, const double* tab,
                                          const double* lrep, int lc, const Tab2Heights& H, const Tab2First& F, Tab2Point& s,
                                          const double* col, unsigned long long* counts, int& record) {
  record = 0;
  if (P.fixed && P.maxiter <= 0) return 0;
  const double tol = P.fixed ? -1.0 : P.tol;
  const int maxiter = P.maxiter;
  int it = 0;
  double drift;
  if (tab2_first_trip<NT>(o, P, T, F, tab, s, col, drift, record)) {
    it = 1;
    o.trip();
    if (!(!(drift < tol) && it < maxiter)) return it;
  }
  do {
    const double pu = s.ustar, pc = s.chi_s;
    tab2_iteration<NT>(o, P, T, rare, Mi, tab, lrep, lc, H, s, col, counts, record);
    if (it == 0) {   // against the initial guess u★ = θ★ = q★ = 1e-4, as written
      const double th = o.mul(s.chi_s, slot_ld(col + SL_DTH * NT)), q = o.mul(s.chi_s, slot_ld(col + SL_DQ * NT));
      drift = o.add(o.add(fabs(o.sub(s.ustar, pu)), fabs(o.sub(th, 1e-4))), fabs(o.sub(q, 1e-4)));
    } else {
      drift = o.fma(fabs(o.sub(s.chi_s, pc)), slot_ld(col + SL_D * NT), fabs(o.sub(s.ustar, pu)));
    }
    ++it;
    o.trip();
  } while (!(drift < tol) && it < maxiter);
  return it;
}

// ---- once per point: q_sat and the iteration invariants ---------------------------------------------------
// Float32 functions of a Float32 thermodynamics, correctly rounded through Float64 (see the header comment)
template <class O>
__device__ __forceinline__ float tab2_powf(O& o, const TabParams& T, const double* tab, float x, float y) {
  return (float)fm::exp_mid(o, T.mc, o.mul((double)y, fm::log_pos(o, tab, T.mc, (double)x)));
}

// p_sat = p_tr (T/T_tr)^(Δcp/R_v) exp[(ℒ₀ − Δcp T₀)/R_v (1/T_tr − 1/T)].  The ARGUMENTS of the power and of the exponential are
// formed with the reference's own operations in its order (IEEE divisions; the constant factors on the host, Thermo::make):
// a one-ulp difference there is multiplied by |argument| ≈ 20 in the result, and q_sat feeds Δq = qₐ − qₛ, whose
// cancellation the pointwise parity criterion sees.  Only the two transcendental evaluations differ from the oracle's
// library calls (≲ 3 ulp).
template <class O>
__device__ __forceinline__ double tab2_psat(O& o, const Thermo<float>& th, const TabParams& T, const double* tab, int phase, double Ts) {
  const float Tf = (float)Ts;
  const int ph = phase == NE_PHASE_LIQUID ? 0 : 1;
  const float x = __fdiv_rn(Tf, th.T_triple);
  const float a = __fmul_rn(th.psat_exp[ph], __fsub_rn(th.inv_T_triple, __fdiv_rn(1.0f, Tf)));
  const float pw = tab2_powf(o, T, tab, x, th.psat_pow[ph]);
  const float ex = (float)fm::exp_mid(o, T.mc, (double)a);
  return (double)__fmul_rn(__fmul_rn(th.press_triple, pw), ex);
}
template <class O>
__device__ __forceinline__ double tab2_psat(O& o, const Thermo<double>& th, const TabParams& T, const double* tab, int phase, double Ts) {
  const int ph = phase == NE_PHASE_LIQUID ? 0 : 1;
  const double x = __ddiv_rn(Ts, th.T_triple);
  const double a = o.mul(th.psat_exp[ph], o.sub(th.inv_T_triple, __ddiv_rn(1.0, Ts)));
  o.other(2);
  const double pw = fm::exp_mid(o, T.mc, o.mul(th.psat_pow[ph], fm::log_pos(o, tab, T.mc, x)));
  return o.mul(o.mul(th.press_triple, pw), fm::exp_mid(o, T.mc, a));
}

// surface_specific_humidity (interface_states.jl:55-74) for Float64 exchange fields
template <class O, class CT>
__device__ __forceinline__ double tab2_surface_humidity(O& o, const NeInterfaceProperties& ip, const Thermo<CT>& th,
                                                        const TabParams& T, const double* tab, double p_at, double Ts, double Ss) {
  const CT p = (CT)p_at;
  // the branch-free log/exp want a positive, finite argument; anything else (it would be a NaN in the reference
  // as well) goes through the library functions
  if (!(Ts > 150.0 && Ts < 400.0 && p > (CT)0)) { o.other(400); return surface_specific_humidity<double, CT>(ip, th, p_at, Ts, Ss); }
  const double psat = tab2_psat(o, th, T, tab, ip.phase, Ts);
  double pv;
  if (ip.x_h2o_kind == NE_XH2O_ONE) pv = psat;
  else if (ip.x_h2o_kind == NE_XH2O_CONSTANT) pv = o.mul(ip.x_h2o, psat);
  else { pv = water_mole_fraction<double>(ip, Ss) * psat; o.other(40); }
  const double lim = (double)((CT)0.999 * p);
  pv = lim < pv ? lim : pv;
  const double num = o.mul((double)th.eps_inv, pv);
  const double den = o.sub((double)p, o.mul((double)((CT)1 - th.eps_inv), pv));
  return fm::div(o, num, den);
}

// what the flux epilogue needs of a point, parked in shared memory across the solve (one slot per thread: conflict-free
// 64-bit accesses, 2 wavefronts each) instead of being re-read from global memory (trip-ordered lanes touch up to 32
// cache lines per load) or kept in registers (the loop already spills at 80)
struct Parked { double du, dv, Ta, pa, qa, Ts; };

// loads of one point: atmosphere state, ocean velocity at the cell centre, surface temperature in Kelvin.  Nothing here
// depends on the land mask, so the kernel requests the mask and the fields together (one memory latency per group)
template <bool HS>
__device__ __forceinline__ void tab2_load(const NeAtmosOceanDesc& d, const Layout& L, int64_t idx, bool celsius, bool relative,
                                          Parked& k, double& uo, double& vo, double& So) {
  k.du = __ldg((const double*)d.ua + idx); k.dv = __ldg((const double*)d.va + idx);
  k.Ta = __ldg((const double*)d.Ta + idx); k.pa = __ldg((const double*)d.pa + idx); k.qa = __ldg((const double*)d.qa + idx);
  uo = vo = 0;
  if (relative) {
    uo = d.uo.ptr ? (slot_at<double>(d.uo, idx) + slot_at<double>(d.uo, idx + 1)) / 2 : d.uo.value;
    vo = d.vo.ptr ? (slot_at<double>(d.vo, idx) + slot_at<double>(d.vo, idx + L.sx)) / 2 : d.vo.value;
  }
  double To = slot_at<double>(d.To, idx);
  if (celsius) To = To + 273.15;
  k.Ts = To;
  So = slot_at<double>(d.So, idx);
}

// iteration invariants of a solved point (BulkTemperature: everything but the iterate is fixed) into the thread's column
template <int NT, class O, class CT>
__device__ __forceinline__ void tab2_invariants(O& o, const NeAtmosOceanDesc& d, const Thermo<CT>& th, const FastParams& P,
                                                const TabParams& T, const double* tab, const Parked& k, double So,
                                                Tab2Point& s, double* col) {
  const double az = d.surface_layer_height.value;
  const double To = k.Ts;
  const double qs = tab2_surface_humidity(o, d.properties, th, T, tab, k.pa, To, So);
  const double Rm = o.add(o.mul((double)th.R_d, o.sub(1.0, qs)), o.mul((double)th.R_v, qs));   // R_d(1 − q) + R_v q
  const double Tv = fm::div(o, o.mul(To, Rm), (double)th.R_d);                                  // virtual_temperature
  const double gTv = fm::div(o, P.g, Tv);
  const double A = o.mul(gTv, o.fma((double)th.delta, qs, 1.0));                                // g/𝒯ₛ (1 + δ qₛ)
  const double B = o.mul(gTv, o.mul((double)th.delta, Tv));                                     // g/𝒯ₛ δ𝒯ₛ
  col[SL_DUDV2 * NT] = o.fma(k.du, k.du, o.mul(k.dv, k.dv));
  const double cpm = o.add(o.mul((double)th.cp_d, o.sub(1.0, k.qa)), o.mul((double)th.cp_v, k.qa));
  const double dth = o.sub(o.add(k.Ta, fm::div(o, o.mul(P.g, az), cpm)), To);                   // θₐ − Tₛ (interface_states.jl:308-317)
  const double dq = o.sub(k.qa, qs);
  col[SL_DTH * NT] = dth;
  col[SL_DQ * NT] = dq;
  col[SL_C * NT] = o.fma(A, dth, o.mul(B, dq));
  col[SL_D * NT] = o.add(fabs(dth), fabs(dq));
  s.ustar = 1e-4;                             // u★ = θ★ = q★ = 1e-4 (atmosphere_ocean_fluxes.jl:131-137)
  s.chi_s = 0;                                // not used by the first trip
  s.bstar = o.fma(A, 1e-4, o.mul(B, 1e-4));
}

// flux epilogue + stores (atmosphere_ocean_fluxes.jl:160-196)
template <class O, class CT>
__device__ __forceinline__ void tab2_epilogue(O& o, const NeAtmosOceanDesc& d, const Thermo<CT>& th, int64_t idx, bool celsius,
                                              bool not_water, const Parked& k, double ustar, double theta_star, double q_star, int iters) {
  const double du = k.du, dv = k.dv;
  double Ts = k.Ts;
  if (not_water) {  // zero_interface_state (interface_states.jl:800-803): Δu = uₐ − 0 (parked as such)
    ustar = 0; theta_star = 0; q_star = 0; Ts = 273.15;
  }
  double Qv, Qc, Jv, tx, ty;
  if (k.Ta > 150.0 && k.pa > 0.0) {
    const double dU2 = o.fma(du, du, o.mul(dv, dv));
    double taux = 0, tauy = 0;
    if (dU2 > 1e-290) {                                               // τ = −u★² Δu/ΔU with the RESOLVED ΔU (:166-170)
      const double m = -o.mul(o.mul(ustar, ustar), fm::rcp(o, fm::sqrt_pos(o, dU2)));
      taux = o.mul(m, du); tauy = o.mul(m, dv);
    } else if (dU2 != 0.0) {
      const double dU = sqrt(dU2);
      taux = -(ustar * ustar) * du / dU; tauy = -(ustar * ustar) * dv / dU;
    }
    const double Rm = o.add(o.mul((double)th.R_d, o.sub(1.0, k.qa)), o.mul((double)th.R_v, k.qa));
    const double rho = fm::div(o, k.pa, o.mul(Rm, k.Ta));              // air_density
    const double cpm = o.add(o.mul((double)th.cp_d, o.sub(1.0, k.qa)), o.mul((double)th.cp_v, k.qa));
    const double Lv = o.fma((double)(th.cp_v - th.cp_l), o.sub(k.Ta, (double)th.T_0), (double)th.LH_v0);
    const double ru = o.mul(-rho, ustar);
    Qv = o.mul(o.mul(o.mul(-rho, Lv), ustar), q_star);
    Qc = o.mul(o.mul(o.mul(-rho, cpm), ustar), theta_star);
    Jv = o.mul(ru, q_star);
    tx = o.mul(rho, taux);
    ty = o.mul(rho, tauy);
  } else {
    AtmosState<double> a;
    a.u = 0; a.v = 0; a.z = 0; a.h_bl = 0; a.T = k.Ta; a.p = k.pa; a.q = k.qa;
    FluxEpilogue<double, CT> e(th, a, ustar, theta_star, q_star, du, dv, false);
    Qv = e.Qv; Qc = e.Qc; Jv = e.Jv; tx = e.tx; ty = e.ty;
    o.other(150);
  }
  ((double*)d.latent_heat)[idx] = Qv;
  ((double*)d.sensible_heat)[idx] = Qc;
  ((double*)d.water_vapor)[idx] = Jv;
  ((double*)d.x_momentum)[idx] = tx;
  ((double*)d.y_momentum)[idx] = ty;
  ((double*)d.interface_temperature)[idx] = celsius ? Ts - 273.15 : Ts;
  ((double*)d.friction_velocity)[idx] = ustar;
  ((double*)d.temperature_scale)[idx] = theta_star;
  ((double*)d.water_vapor_scale)[idx] = q_star;
  if (d.iterations) d.iterations[idx] = iters;
}

#endif  // __CUDACC__

}  // namespace ne
