// fastmath_check.cu — host-side accuracy check of ne_fastmath.cuh / ne_flux_tab.cuh (no GPU needed).
//   nvcc -O2 -std=c++17 -o /tmp/fastmath_check tools/fastmath_check.cu && /tmp/fastmath_check
// Prints one JSON object: max relative error (in units of 2^-53) of each elementary function against
// long double, and the max abs error of the ψ tables for the default Edson parameters.
#include <cstdio>
#include <random>

#include "../numericalearth.jl_b200/csrc/ne_flux_tab.cuh"

using namespace ne;

static void default_formulation(NeFluxFormulation& f) {
  std::memset(&f, 0, sizeof(f));
  double m[11] = {50, 0.35, 0.7, 0.75, 5 / 0.35, 15, 2, 3.141592653589793 / 2, 10.15, 3, 3.141592653589793 / std::sqrt(3.0)};
  double s[12] = {50, 0.35, 2.0 / 3, 1.5, 14.28, 8.525, 15, 2, 0, 34.15, 3, 3.141592653589793 / std::sqrt(3.0)};
  f.psi_momentum.a.kind = NE_PSI_EDSON_MOMENTUM;
  f.psi_temperature.a.kind = f.psi_water_vapor.a.kind = NE_PSI_EDSON_SCALAR;
  for (int k = 0; k < 11; ++k) f.psi_momentum.a.p[k] = m[k];
  for (int k = 0; k < 12; ++k) { f.psi_temperature.a.p[k] = s[k]; f.psi_water_vapor.a.p[k] = s[k]; }
  f.subgrid_velocities.gustiness_parameter = 1.2;
  f.subgrid_velocities.minimum_gustiness = 0.01;
}

// atmosphere_sea_ice_stability_functions (similarity_theory_turbulent_fluxes.jl:779-789): Split(SHEBA, Paulson)
static void sea_ice_formulation(NeFluxFormulation& f) {
  default_formulation(f);
  auto split = [](NeStabilityProfile& s, int stable_kind, const double* ps, int np, int unstable_kind, const double* pu, int nu) {
    std::memset(&s, 0, sizeof(s));
    s.split = 1;
    s.a.kind = stable_kind;
    for (int k = 0; k < np; ++k) s.a.p[k] = ps[k];
    s.b.kind = unstable_kind;
    for (int k = 0; k < nu; ++k) s.b.p[k] = pu[k];
  };
  const double sm[2] = {6.5, 1.3}, ss[3] = {5.0, 5.0, 3.0}, pm[2] = {16.0, 3.141592653589793 / 2}, pp[1] = {16.0};
  split(f.psi_momentum, NE_PSI_SHEBA_MOMENTUM, sm, 2, NE_PSI_PAULSON_MOMENTUM, pm, 2);
  split(f.psi_temperature, NE_PSI_SHEBA_SCALAR, ss, 3, NE_PSI_PAULSON_SCALAR, pp, 1);
  f.psi_water_vapor = f.psi_temperature;
}
// large_yeager_stability_functions (:766-771): Split(LinearStable, Paulson)
static void large_yeager_formulation(NeFluxFormulation& f) {
  sea_ice_formulation(f);
  const double ls[2] = {5.0, 10.0};
  for (NeStabilityProfile* s : {&f.psi_momentum, &f.psi_temperature, &f.psi_water_vapor}) {
    std::memset(&s->a, 0, sizeof(s->a));
    s->a.kind = NE_PSI_LINEAR_STABLE;
    s->a.p[0] = ls[0]; s->a.p[1] = ls[1];
  }
}

// dense check of a general (non-Edson) table against psi_profile_ld, both sides, relative to max(1, |ψ|)
static double dense_general(const NeFluxFormulation& f, const double* tab, std::mt19937_64& rng) {
  std::uniform_real_distribution<double> u(0, 1);
  double worst = 0;
  for (int i = 0; i < 200000; ++i) {
    const double az = std::exp(std::log(1e-9) + u(rng) * (std::log(127.99) - std::log(1e-9)));
    for (int side = 0; side < 2; ++side) {
      const double z = side ? az : -az;
      bool outside;
      const int iv = fm::psi_interval(z, outside);
      if (outside) return 1e300;
      double m, s;
      fm::psi_pair(tab + fm::TAB_PSI + iv * fm::PSI_REC, az, m, s);
      const long double tm = psi_profile_ld(f.psi_momentum, z), ts = psi_profile_ld(f.psi_temperature, z);
      const double em = (double)(fabsl((long double)m - tm) / fmaxl(1, fabsl(tm)));
      const double es = (double)(fabsl((long double)s - ts) / fmaxl(1, fabsl(ts)));
      if (!(em <= worst)) worst = em;
      if (!(es <= worst)) worst = es;
      if (az < std::ldexp(1.0, fm::TINY_EXP)) {
        fm::psi_tiny_pair(tab + fm::TAB_TINY + side * fm::TINY_REC, az, az, m, s);
        const double e2 = (double)fmaxl(fabsl((long double)m - tm), fabsl((long double)s - ts));
        if (!(e2 <= worst)) worst = e2;
      }
    }
  }
  return worst;
}

template <class F, class G>
static double max_ulp(F f, G truth, double lo, double hi, bool logspace, int n, std::mt19937_64& rng) {
  std::uniform_real_distribution<double> u(0, 1);
  double worst = 0;
  for (int i = 0; i < n; ++i) {
    double t = u(rng);
    double x = logspace ? std::exp(std::log(lo) + t * (std::log(hi) - std::log(lo))) : lo + t * (hi - lo);
    long double ref = truth((long double)x);
    double got = f(x);
    double e = (double)(fabsl((long double)got - ref) / fabsl(ref)) * 9007199254740992.0;
    if (!(e <= worst)) worst = e;
  }
  return worst;
}

int main() {
  NeFluxFormulation f;
  default_formulation(f);
  static double tab[fm::TAB_SIZE];
  TabParams T;
  double fit = build_solver_tables(f, tab, T);
  std::mt19937_64 rng(20261017);
  double e_rcp = max_ulp([](double x) { return fm::rcp(x); }, [](long double x) { return 1 / x; }, 1e-6, 1e6, true, 400000, rng);
  double e_div = max_ulp([](double x) { return fm::div(0.4, x); }, [](long double x) { return 0.4L * 1 / x * 1.0L == 0 ? 0 : (long double)0.4 / x; }, 1e-3, 1e3, true, 400000, rng);
  double e_sqrt = max_ulp([](double x) { return fm::sqrt_pos(x); }, [](long double x) { return sqrtl(x); }, 1e-8, 1e8, true, 400000, rng);
  double e_cbrt = max_ulp([&](double x) { return fm::cbrt_pos(T.mc, x); }, [](long double x) { return cbrtl(x); }, 1e-12, 1e12, true, 400000, rng);
  double e_cbrt2 = max_ulp([&](double x) { return fm::cbrt_pos(T.mc, x); }, [](long double x) { return cbrtl(x); }, 1e-300, 1e300, true, 400000, rng);
  fm::OpsPlain ops;
  double e_rcp3 = max_ulp([&](double x) { return fm::rcp3(ops, x); }, [](long double x) { return 1 / x; }, 1e-6, 1e6, true, 400000, rng);
  double e_sqrt3 = max_ulp([&](double x) { return fm::sqrt3(ops, x); }, [](long double x) { return sqrtl(x); }, 1e-8, 1e8, true, 400000, rng);
  double e_cbrt3 = max_ulp([&](double x) { return fm::cbrt3(ops, T.mc, x); }, [](long double x) { return cbrtl(x); }, 1e-300, 1e300, true, 400000, rng);
  // log: relative error away from 1, absolute error (in units of 2^-53) near 1
  double e_log = max_ulp([&](double x) { return fm::log_pos(tab, T.mc, x); }, [](long double x) { return logl(x); }, 1e-12, 0.5, true, 400000, rng);
  double e_log_hi = max_ulp([&](double x) { return fm::log_pos(tab, T.mc, x); }, [](long double x) { return logl(x); }, 2.0, 1e12, true, 400000, rng);
  double e_log_abs = 0;
  {
    std::uniform_real_distribution<double> u(0.5, 2.0);
    for (int i = 0; i < 400000; ++i) {
      double x = u(rng);
      double e = (double)fabsl((long double)fm::log_pos(tab, T.mc, x) - logl((long double)x)) * 9007199254740992.0;
      if (!(e <= e_log_abs)) e_log_abs = e;
    }
  }
  double e_exp = max_ulp([&](double x) { return fm::exp_mid(T.mc, x); }, [](long double x) { return expl(x); }, -700, 700, false, 400000, rng);
  double e_exp2 = max_ulp([&](double x) { return fm::exp_mid(T.mc, x); }, [](long double x) { return expl(x); }, -30, 1, false, 400000, rng);
  // the short exponential (relative error, not ulp) and the saturated-exponential constants of the stable closed forms
  double e_exp_lo = 0;
  {
    std::uniform_real_distribution<double> u(-700, 40);
    fm::OpsPlain o;
    for (int i = 0; i < 400000; ++i) {
      const double x = u(rng);
      const long double t = expl((long double)x);
      const double e = (double)(fabsl((long double)fm::exp_lo(o, tab, T.mc, x) - t) / t);
      if (!(e <= e_exp_lo)) e_exp_lo = e;
    }
  }
  const double e_sat = std::fmax(std::fabs(T.em_sat - std::exp(-f.psi_momentum.a.p[0])) / std::exp(-f.psi_momentum.a.p[0]),
                                 std::fabs(T.es_sat - std::exp(-f.psi_temperature.a.p[0])) / std::exp(-f.psi_temperature.a.p[0]));
  const int sat_ok = T.z_sat * f.psi_momentum.a.p[1] >= f.psi_momentum.a.p[0] && T.z_sat * f.psi_temperature.a.p[1] >= f.psi_temperature.a.p[0] &&
                     T.z_sat < 1e3;
  // bit-derived w and the replicated log table: the same bits as the stored (a, b) / the plain table
  int bits_ok = 1;
  {
    static double lrep[2 * fm::LOG_N * fm::LOG_REP];
    for (int k = 0; k < fm::LOG_N * fm::LOG_REP; ++k) { lrep[2 * k] = tab[fm::TAB_LOG + 2 * (k / fm::LOG_REP)]; lrep[2 * k + 1] = tab[fm::TAB_LOG + 2 * (k / fm::LOG_REP) + 1]; }
    std::uniform_real_distribution<double> u(0, 1);
    fm::OpsPlain o;
    for (int i = 0; i < 400000; ++i) {
      const double az = std::exp(std::log(1e-12) + u(rng) * (std::log(127.99) - std::log(1e-12)));
      for (int side = 0; side < 2; ++side) {
        bool outside;
        const int iv = fm::psi_interval(side ? az : -az, outside);
        double m1, s1, m2, s2;
        fm::psi_pair(o, tab + fm::TAB_PSI + iv * fm::PSI_REC, az, m1, s1);
        fm::psi_pair_bits(o, tab + fm::TAB_PSI + iv * fm::PSI_REC, az, m2, s2);
        if (outside || std::memcmp(&m1, &m2, 8) || std::memcmp(&s1, &s2, 8)) bits_ok = 0;
      }
      const double x = std::exp(-30 + 60 * u(rng));
      const double l1 = fm::log_pos(o, tab, T.mc, x), l2 = fm::log_pos_rep(o, lrep, i & (fm::LOG_REP - 1), T.mc, x);
      if (std::memcmp(&l1, &l2, 8)) bits_ok = 0;
    }
    // interval edges: exact powers of two and sub-interval boundaries
    for (int e = fm::PSI_OCT_LO - 2; e < fm::PSI_OCT_HI; ++e)
      for (int k = 0; k < fm::PSI_SUB; ++k)
        for (int d = -1; d <= 1; ++d) {
          double az = std::ldexp(1.0 + (double)k / fm::PSI_SUB, e);
          if (d) az = std::nextafter(az, d > 0 ? 1e300 : 0.0);
          bool outside;
          const int iv = fm::psi_interval(az, outside);
          double m1, s1, m2, s2;
          fm::psi_pair(o, tab + fm::TAB_PSI + iv * fm::PSI_REC, az, m1, s1);
          fm::psi_pair_bits(o, tab + fm::TAB_PSI + iv * fm::PSI_REC, az, m2, s2);
          if (std::memcmp(&m1, &m2, 8) || std::memcmp(&s1, &s2, 8)) bits_ok = 0;
        }
  }
  // ψ tables: dense independent check
  const double* pm = f.psi_momentum.a.p;
  const double* ps = f.psi_temperature.a.p;
  double e_psi = 0;
  {
    std::uniform_real_distribution<double> u(0, 1);
    for (int i = 0; i < 400000; ++i) {
      double az = std::exp(std::log(1e-9) + u(rng) * (std::log(127.99) - std::log(1e-9)));
      bool outside;
      int iv = fm::psi_interval(-az, outside);
      if (outside) { e_psi = 1e300; break; }
      double m, s;
      fm::psi_pair(tab + fm::TAB_PSI + iv * fm::PSI_REC, az, m, s);
      auto rel = [](double got, long double t) { return (double)(fabsl((long double)got - t) / fmaxl(1, fabsl(t))); };
      double em = rel(m, psi_m_unstable_ld(pm, -(long double)az));
      double es = rel(s, psi_s_unstable_ld(ps, -(long double)az));
      if (!(em <= e_psi)) e_psi = em;
      if (!(es <= e_psi)) e_psi = es;
      {
        iv = fm::psi_interval(az, outside);
        if (outside) { e_psi = 1e300; break; }
        fm::psi_pair(tab + fm::TAB_PSI + iv * fm::PSI_REC, az, m, s);
        em = rel(m, psi_m_stable_ld(pm, (long double)az));
        es = rel(s, psi_s_stable_ld(ps, (long double)az));
        if (!(em <= e_psi)) e_psi = em;
        if (!(es <= e_psi)) e_psi = es;
      }
    }
  }
  auto iv = [](double z) { bool o; int i = fm::psi_interval(z, o); return o ? -1 : i; };
  // geometry-independent statements of the interval logic: 2^PSI_OCT_LO opens record 1, 0.5 = 2^-1 opens the first record
  // of its octave, everything below 2^PSI_OCT_LO shares record 0, |ζ| ≥ 2^7 and NaN are outside
  const double lo = std::ldexp(1.0, fm::PSI_OCT_LO);
  int iv_ok = iv(-128.0) == -1 && iv(128.0) == -1 && iv(0.5) == fm::PSI_NS + 1 + fm::PSI_SUB * (-1 - fm::PSI_OCT_LO) &&
              iv(0.0) == fm::PSI_NS && iv(-1e-300) == 0 &&
              iv(-lo) == 1 && iv(-127.9) == fm::PSI_NQ && iv(NAN) == -1 && iv(-1e300) == -1 && iv(0.96 * lo) == fm::PSI_NS &&
              iv(lo * (1.0 + 1.0 / fm::PSI_SUB)) == fm::PSI_NS + 2 &&
              iv(127.9) == fm::PSI_NS + fm::PSI_NQ &&
              fm::psi_is_tiny(std::ldexp(0.99, fm::TINY_EXP)) && !fm::psi_is_tiny(std::ldexp(1.0, fm::TINY_EXP)) && fm::psi_is_tiny(-1e-9) && !fm::psi_is_tiny(NAN);
  double e_tiny = 0;
  {
    std::uniform_real_distribution<double> u(0, 1);
    for (int i = 0; i < 200000; ++i) {
      double az = std::exp(std::log(1e-12) + u(rng) * (std::log(std::ldexp(0.9999, fm::TINY_EXP)) - std::log(1e-12)));
      double m, s;
      fm::psi_tiny_pair(tab + fm::TAB_TINY, az, az * 0.7, m, s);
      double em = (double)fabsl((long double)m - psi_m_unstable_ld(pm, -(long double)az));
      double es = (double)fabsl((long double)s - psi_s_unstable_ld(ps, -(long double)(az * 0.7)));
      if (!(em <= e_tiny)) e_tiny = em;
      if (!(es <= e_tiny)) e_tiny = es;
      fm::psi_tiny_pair(tab + fm::TAB_TINY + fm::TINY_REC, az, az * 0.7, m, s);
      em = (double)fabsl((long double)m - psi_m_stable_ld(pm, (long double)az));
      es = (double)fabsl((long double)s - psi_s_stable_ld(ps, (long double)(az * 0.7)));
      if (!(em <= e_tiny)) e_tiny = em;
      if (!(es <= e_tiny)) e_tiny = es;
    }
  }
  double e_micro = 0;
  {
    std::uniform_real_distribution<double> u(0, 1);
    for (int i = 0; i < 200000; ++i) {
      double az = std::exp(std::log(1e-14) + u(rng) * (std::log(std::ldexp(0.9999, fm::MICRO_EXP)) - std::log(1e-14)));
      double m, s2;
      fm::psi_micro_pair(tab + fm::TAB_MICRO, az, az * 0.7, m, s2);
      double em = (double)fabsl((long double)m - psi_m_unstable_ld(pm, -(long double)az));
      double es = (double)fabsl((long double)s2 - psi_s_unstable_ld(ps, -(long double)(az * 0.7)));
      if (!(em <= e_micro)) e_micro = em;
      if (!(es <= e_micro)) e_micro = es;
      fm::psi_micro_pair(tab + fm::TAB_MICRO + fm::MICRO_REC, az, az * 0.7, m, s2);
      em = (double)fabsl((long double)m - psi_m_stable_ld(pm, (long double)az));
      es = (double)fabsl((long double)s2 - psi_s_stable_ld(ps, (long double)(az * 0.7)));
      if (!(em <= e_micro)) e_micro = em;
      if (!(es <= e_micro)) e_micro = es;
    }
    if (!(fm::psi_is_micro(std::ldexp(0.99, fm::MICRO_EXP)) && !fm::psi_is_micro(std::ldexp(1.0, fm::MICRO_EXP)) && fm::psi_is_micro(-1e-9))) e_micro = 1e300;
  }
  // far-unstable closed forms with the branch-free functions (ζ ≤ −2^7) against long double
  double e_far = 0, e_atan = 0;
  {
    const FastParams P = make_fast_params(f, 9.80665);
    std::uniform_real_distribution<double> u(0, 1);
    for (int i = 0; i < 200000; ++i) {
      const double az = std::exp(std::log(128.0) + u(rng) * (std::log(1e9) - std::log(128.0)));
      double m, s2;
      psi_far_unstable_fm(P, T, tab, -az, m, s2);
      const long double tm = psi_m_unstable_ld(pm, -(long double)az), ts = psi_s_unstable_ld(ps, -(long double)az);
      const double em = (double)(fabsl((long double)m - tm) / fmaxl(1, fabsl(tm))), es = (double)(fabsl((long double)s2 - ts) / fmaxl(1, fabsl(ts)));
      if (!(em <= e_far)) e_far = em;
      if (!(es <= e_far)) e_far = es;
      const double x = 6.0 + az * 1e-3;
      const double ea = (double)fabsl((long double)fm::atan_large(x) - atanl((long double)x));
      if (!(ea <= e_atan)) e_atan = ea;
    }
    if (!far_unstable_fm_ok(P)) e_far = 1e300;
  }
  NeFluxFormulation fi, fl;
  sea_ice_formulation(fi);
  large_yeager_formulation(fl);
  static double tab_i[fm::TAB_SIZE], tab_l[fm::TAB_SIZE];
  TabParams Ti, Tl;
  const double fit_i = build_solver_tables(fi, tab_i, Ti), fit_l = build_solver_tables(fl, tab_l, Tl);
  const double dense_i = dense_general(fi, tab_i, rng), dense_l = dense_general(fl, tab_l, rng);
  printf("{\"psi_far_err\": %.3e, \"atan_large_abs\": %.3e, \"psi_micro_abs\": %.3e, ", e_far, e_atan, e_micro);
  printf("\"bit_w_and_replicated_log_same_bits\": %d, ", bits_ok);
  printf("\"rcp3_ulp\": %.3f, \"sqrt3_ulp\": %.3f, \"cbrt3_ulp\": %.3f, ", e_rcp3, e_sqrt3, e_cbrt3);
  printf("\"exp_lo_rel\": %.3e, \"exp_sat_rel\": %.3e, \"z_sat_ok\": %d, ", e_exp_lo, e_sat, sat_ok);
  printf("\"psi_seaice_fit_err\": %.3e, \"psi_seaice_dense_err\": %.3e, \"psi_seaice_general\": %d, "
         "\"psi_ly_fit_err\": %.3e, \"psi_ly_dense_err\": %.3e, ", fit_i, dense_i, Ti.general_psi, fit_l, dense_l);
  printf("\"rcp_ulp\": %.3f, \"div_ulp\": %.3f, \"sqrt_ulp\": %.3f, \"cbrt_ulp\": %.3f, \"cbrt_wide_ulp\": %.3f, "
         "\"log_ulp_small\": %.3f, \"log_ulp_large\": %.3f, \"log_abs_near1_ulp1\": %.3f, \"exp_ulp\": %.3f, \"exp_ulp_mid\": %.3f, "
         "\"psi_fit_err\": %.3e, \"psi_dense_err\": %.3e, \"psi_tiny_abs\": %.3e, \"interval_logic_ok\": %d, \"cbrt_floor\": %.6e, \"same_exp\": %d}\n",
         e_rcp, e_div, e_sqrt, e_cbrt, e_cbrt2, e_log, e_log_hi, e_log_abs, e_exp, e_exp2, fit, e_psi, e_tiny, iv_ok, T.cbrt_floor, T.same_exp);
  return 0;
}
