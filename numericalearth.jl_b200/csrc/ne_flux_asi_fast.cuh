// ne_flux_asi_fast.cuh — atmosphere–sea-ice turbulent fluxes, default plugin tree, on the work-queue kernel.
//
// Replaces _compute_atmosphere_sea_ice_interface_state! (atmosphere_sea_ice_fluxes.jl:65-185) for
//   SimilarityTheoryFluxes (logarithmic profile; ANY pair of the shipped stability functions whose piecewise
//   polynomial tables verify — the default is Split(SHEBA stable, Paulson unstable),
//   similarity_theory_turbulent_fluxes.jl:779-794 —, momentum + Reynolds-scaled scalar roughness lengths with
//   constant viscosity, convective gustiness), SkinTemperature(ConductiveFlux | IceSnowConductiveFlux) or
//   BulkTemperature, any saturation-humidity variant, Float64 exchange grid, Float32 or Float64 thermodynamics.
// Everything else keeps the generic kernel (ne_flux_generic.cuh).
//
// Why a kernel of its own: sea ice covers a few percent of the exchange grid in contiguous polar caps, its skin
// temperature makes the fixed point slower (mean 34 trips, 19 % of the synthetic points sit on a limit cycle and
// run all 100), so one thread per point leaves most lanes idle.  The work-queue kernel compacts the ice points
// into dense rounds and defers the long runners (ne_flux_queue.cuh).  One trip = skin temperature from the
// conductive flux balance (interface_states.jl:468-508), q_sat over ice at the new Tₛ, then the same table-driven
// similarity step as the a–o kernel (ne_flux_tab.cuh).
#pragma once

#include "ne_flux_queue.cuh"
#include "ne_flux_tab2.cuh"   // tab2_surface_humidity: q_sat from the branch-free log / exp (same rounding rules as the a-o prologue)

namespace ne {

template <class FT>
struct AsiPoint {
  // invariants
  FT theta_a, dudv2, h_bl, hd;
  double log_hd;
  FT ap, aq;
  FT rho_c, rho_L;            // ρₐ cₚₘ, ρₐ ℒˢ
  FT Qd, R, Tb, sig_eps;
  // iterate
  FT ustar, theta_star, q_star, Ts;
};

template <class FT, class CT>
__device__ __noinline__ void asi_write_outputs(const NeAtmosSeaIceDesc& d, const Thermo<CT>& th, int64_t idx,
                                               FT ustar, FT theta_star, FT q_star, FT Ts, int iters) {
  AtmosState<FT> a;
  a.u = __ldg((const FT*)d.ua + idx);
  a.v = __ldg((const FT*)d.va + idx);
  a.T = __ldg((const FT*)d.Ta + idx);
  a.p = __ldg((const FT*)d.pa + idx);
  a.q = __ldg((const FT*)d.qa + idx);
  a.z = 0; a.h_bl = 0;
  FluxEpilogue<FT, CT> e(th, a, ustar, theta_star, q_star, a.u, a.v, true);   // Δu = uₐ − 0 (:97-98)
  ((FT*)d.latent_heat)[idx] = e.Qv;
  ((FT*)d.sensible_heat)[idx] = e.Qc;
  ((FT*)d.water_vapor)[idx] = e.Jv;
  ((FT*)d.x_momentum)[idx] = e.tx;
  ((FT*)d.y_momentum)[idx] = e.ty;
  ((FT*)d.interface_temperature)[idx] = d.sea_ice.temperature_units == NE_DEGREES_CELSIUS ? Ts - (FT)273.15 : Ts;
  if (d.iterations) d.iterations[idx] = iters;
}

// FT = double: Float64 model.  FT = float: Float32 model with Float32 thermodynamics and the Float64-literal viscosity —
// skin temperature and q_sat in Float32, the similarity step mixed precision exactly as in the a–o Float32 kernel
// (tab_iteration(…, FastPointF&), ne_flux_tab.cuh).
template <class FT_, class CT, bool HS>
struct AsiProblem {
  using FT = FT_;
  using Point = AsiPoint<FT>;
  static constexpr int NSTATE = 4;
  static constexpr bool F32 = std::is_same<FT, float>::value;
  struct Params {
    NeAtmosSeaIceDesc d;
    Layout L;
    Thermo<CT> th;
    FastParams P;
    TabParams T;
    FrontF32 Q;
  };
  __device__ static __forceinline__ const Layout& layout(const Params& p) { return p.L; }
  __device__ static __forceinline__ const FastParams& fast(const Params& p) { return p.P; }
  __device__ static __forceinline__ FT tolerance(const Params& p) { return F32 ? (FT)p.Q.tol : (FT)p.P.tol; }
  __device__ static __forceinline__ FT gravity(const Params& p) { return F32 ? (FT)p.Q.g : (FT)p.P.g; }
  __device__ static __forceinline__ FT zero_plane(const Params& p) { return F32 ? (FT)p.Q.d_zero : (FT)p.P.d_zero; }

  __device__ static __forceinline__ FT initial_Ts(const Params& p, int32_t idx) {
    FT Ts0 = ((const FT*)p.d.interface_temperature)[idx];
    if (p.d.sea_ice.temperature_units == NE_DEGREES_CELSIUS) Ts0 = Ts0 + (FT)273.15;
    return Ts0;
  }

  __device__ static __forceinline__ bool admit(const Params& p, int32_t idx) {
    const NeAtmosSeaIceDesc& d = p.d;
    const bool not_water = d.inactive ? (d.inactive[idx] != 0) : false;
    const bool ice_free = slot_at<FT>(d.concentration, idx) == 0;
    if ((!p.P.fixed && not_water) || ice_free) {   // :141-142: zero scales, Tₛ = ocean surface temperature
      FT To = slot_at<FT>(d.To, idx);
      if (d.ocean.temperature_units == NE_DEGREES_CELSIUS) To = To + (FT)273.15;
      asi_write_outputs<FT, CT>(d, p.th, idx, (FT)0, (FT)0, (FT)0, To, 0);
      return false;
    }
    if (p.P.fixed && p.P.maxiter <= 0) {           // no trip: the initial state (:127-131)
      const FT x0 = (FT)1e-4f;
      asi_write_outputs<FT, CT>(d, p.th, idx, x0, x0, x0, initial_Ts(p, idx), 0);
      return false;
    }
    return true;
  }

  __device__ static __forceinline__ void prologue(const Params& p, int32_t idx, Point& s, bool fresh) {
    const NeAtmosSeaIceDesc& d = p.d;
    const Thermo<CT>& th = p.th;
    const FT au = __ldg((const FT*)d.ua + idx), av = __ldg((const FT*)d.va + idx);
    const FT aT = __ldg((const FT*)d.Ta + idx), ap = __ldg((const FT*)d.pa + idx), aq = __ldg((const FT*)d.qa + idx);
    const FT az = HS ? (FT)d.surface_layer_height.value : slot_at<FT>(d.surface_layer_height, idx);
    s.theta_a = aT + gravity(p) * az / th.cp_m(aq);   // surface_atmosphere_temperature interface_states.jl:308-317
    s.dudv2 = au * au + av * av;                      // ice velocity forced to 0 (:97-98)
    s.h_bl = HS ? (FT)d.boundary_layer_height.value : slot_at<FT>(d.boundary_layer_height, idx);
    s.hd = az - zero_plane(p);
    s.log_hd = HS ? p.T.log_hd : log((double)s.hd);
    s.ap = ap; s.aq = aq;
    const auto rho_a = th.air_density(aT, ap, aq);
    s.rho_c = rho_a * th.cp_m(aq);
    s.rho_L = rho_a * th.latent_heat_sublim(aT);      // sublimation enthalpy for every surface (:542-544)
    const int32_t j = (int32_t)(idx / p.L.sx) - (p.L.hy - 1);
    const RadState<FT> rad = radiation_state<FT>(d.radiation, p.L, idx, j);
    s.sig_eps = rad.sigma * rad.eps;
    s.Qd = -(1 - rad.alpha) * rad.sw - rad.eps * rad.lw;
    const NeInterfaceProperties& ip = d.properties;
    const FT hi = slot_at<FT>(d.hi, idx), hc = slot_at<FT>(d.hc, idx);
    if (ip.temperature_formulation == NE_TEMP_SKIN_CONDUCTIVE) s.R = hi / (FT)ip.ice_conductivity;
    else s.R = slot_at<FT>(d.hs, idx) / (FT)ip.snow_conductivity + hi / (FT)ip.ice_conductivity;
    FT Tb = (FT)d.sea_ice.liquidus_freshwater_melting_temperature - (FT)d.sea_ice.liquidus_slope * slot_at<FT>(d.So, idx);
    if (d.sea_ice.temperature_units == NE_DEGREES_CELSIUS) Tb = Tb + (FT)273.15;
    s.Tb = Tb;
    if (!(hi >= hc)) s.R = (FT)-1;                    // thin ice: Tₛ = T_b whatever the balance says (:505-507)
    if (fresh) {
      s.ustar = s.theta_star = s.q_star = (FT)1e-4f;  // convert(FT, 1f-4) :127
      s.Ts = initial_Ts(p, idx);
    }
  }
  __device__ static __forceinline__ void get_state(const Point& s, FT* v) { v[0] = s.ustar; v[1] = s.theta_star; v[2] = s.q_star; v[3] = s.Ts; }
  __device__ static __forceinline__ void set_state(Point& s, const FT* v) { s.ustar = v[0]; s.theta_star = v[1]; s.q_star = v[2]; s.Ts = v[3]; }

  // conductive_flux_balance_temperature (interface_states.jl:468-508)
  __device__ static __forceinline__ FT skin_temperature(const Params& p, const Point& s) {
    const NeInterfaceProperties& ip = p.d.properties;
    const FT Tsm = s.Ts;
    FT Tm = (FT)p.d.sea_ice.liquidus_freshwater_melting_temperature;
    if (p.d.sea_ice.temperature_units == NE_DEGREES_CELSIUS) Tm = Tm + (FT)273.15;
    if (s.R < 0) return s.Tb;
    const FT lw_up = s.sig_eps * pow4(Tsm);
    const FT QT = -s.rho_c * s.ustar * s.theta_star;
    const FT Qv = -s.rho_L * s.ustar * s.q_star;
    const FT dT = s.theta_a - Tsm;
    const FT Qa = Qv + lw_up + s.Qd;
    // Float64 model: the three quotients of the balance through the reciprocal-based division (≲ 1 ulp; a sea-ice point's
    // solve is a chain of up to 100 dependent trips and the kernel's time is the latency of its slowest points, so every
    // IEEE division — ~35 dependent instructions — is on the critical path); arguments that would not be normal numbers
    // (ΔT = 0, D = 0) are replaced by the selects below exactly as in the reference
    auto quot = [](FT a, FT b) -> FT {
      if constexpr (F32) return a / b;
      else return (b == 0) ? a / b : (FT)fm::div((double)a, (double)b);
    };
    const FT Oc = (dT == 0) ? (FT)0 : quot(QT, dT);
    const FT beta = quot(4 * lw_up, Tsm);
    const FT R = s.R;
    const FT D = 1 + beta * R - Oc * R;
    FT Tstar = quot(s.Tb + beta * R * Tsm - Oc * R * s.theta_a - Qa * R, D);
    Tstar = (D == 0) ? Tsm : Tstar;
    Tstar = (Tstar != Tstar) ? Tsm : Tstar;
    const FT maxdT = (FT)ip.max_dT;
    const FT Tsp = Tsm + clampv<FT>(Tstar - Tsm, -maxdT, maxdT);
    return mn(Tsp, Tm);
  }

  __device__ static __forceinline__ FT trip(const Params& p, const double* tab, Point& s, int) {
    const NeInterfaceProperties& ip = p.d.properties;
    if (ip.temperature_formulation != NE_TEMP_BULK) s.Ts = skin_temperature(p, s);
    FT qs, Tv;
    typename QPointOf<FT>::type f;
    if constexpr (F32) {
      qs = surface_specific_humidity<FT, CT>(ip, p.th, s.ap, s.Ts, (FT)0);   // humidity scalar 0 over ice (:737)
      Tv = p.th.virtual_temperature(s.Ts, qs);
      f.gTv = gravity(p) / Tv;
    } else {   // q_sat(Tₛ) is re-evaluated every trip: the power and the exponential from the branch-free log / exp
      fm::OpsPlain o;
      qs = tab2_surface_humidity(o, ip, p.th, p.T, tab, (double)s.ap, (double)s.Ts, 0.0);
      Tv = fm::div(s.Ts * p.th.gas_constant_air(qs), (double)p.th.R_d);          // virtual_temperature
      f.gTv = fm::div((double)gravity(p), Tv);
    }
    f.c1 = 1 + p.th.delta * qs;
    f.c2 = p.th.delta * Tv;
    f.dudv2 = s.dudv2; f.h_bl = s.h_bl; f.hd = s.hd; f.log_hd = s.log_hd;
    f.dtheta = s.theta_a - s.Ts;
    f.dq = s.aq - qs;
    f.ustar = s.ustar; f.theta_star = s.theta_star; f.q_star = s.q_star;
    const NeFluxFormulation* ff = p.T.general_psi ? &p.d.flux : nullptr;
    if constexpr (F32) tab_iteration(p.P, p.Q, p.T, tab, f, ff);
    else tab_iteration<true>(p.P, p.T, tab, f, ff);
    const FT drift = m_abs(f.ustar - s.ustar) + m_abs(f.theta_star - s.theta_star) + m_abs(f.q_star - s.q_star);
    s.ustar = f.ustar; s.theta_star = f.theta_star; s.q_star = f.q_star;
    return drift;
  }

  __device__ static __forceinline__ void finish(const Params& p, int32_t idx, const Point& s, int it) {
    asi_write_outputs<FT, CT>(p.d, p.th, idx, s.ustar, s.theta_star, s.q_star, s.Ts, it);
  }
};

// the roughness / gustiness / profile part of the default tree (shared with the a–o eligibility test)
inline bool default_roughness_gustiness(const NeFluxFormulation& f) {
  if (f.kind != NE_FLUX_SIMILARITY_THEORY) return false;
  if (f.similarity_form != NE_PROFILE_LOGARITHMIC) return false;
  if (std::memcmp(&f.psi_temperature, &f.psi_water_vapor, sizeof(NeStabilityProfile)) != 0) return false;
  if (std::memcmp(&f.ell_temperature, &f.ell_water_vapor, sizeof(NeRoughnessLength)) != 0) return false;
  const NeRoughnessLength& m = f.ell_momentum;
  const NeRoughnessLength& s = f.ell_temperature;
  if (m.kind != NE_ROUGH_MOMENTUM || m.visc_kind != NE_VISC_CONSTANT) return false;
  if (m.wave_kind != NE_WAVE_CONSTANT && m.wave_kind != NE_WAVE_WIND_DEPENDENT) return false;
  if (s.kind != NE_ROUGH_SCALAR || s.visc_kind != NE_VISC_CONSTANT || s.nu != m.nu) return false;
  if (!(m.nu > 0) || !(s.reynolds_A > 0) || !(s.maximum_roughness_length > 0) || !(m.maximum_roughness_length > 0)) return false;
  if (!(m.gravitational_acceleration > 0)) return false;
  // ℓu must stay positive: a smooth-wall term, or a wave term that cannot vanish
  if (!(m.smooth_wall_parameter > 0) && !(m.wave_kind == NE_WAVE_CONSTANT && m.wave_constant > 0)) return false;
  const NeSubgridVelocity& g = f.subgrid_velocities;
  if (g.convective_kind != NE_SGS_CONVECTIVE || !(g.minimum_gustiness > 0)) return false;
  if (g.composite && g.mesoscale_kind != NE_SGS_NONE && g.mesoscale_kind != NE_SGS_CONSTANT) return false;
  return true;
}

// strict default tree: logarithmic profile, constant wave parameter, no mesoscale term (the EXT = false kernels)
inline bool strict_default_options(const NeFluxFormulation& f) {
  const NeSubgridVelocity& g = f.subgrid_velocities;
  return f.similarity_form == NE_PROFILE_LOGARITHMIC && f.ell_momentum.wave_kind == NE_WAVE_CONSTANT &&
         !(g.composite && g.mesoscale_kind == NE_SGS_CONSTANT && g.mesoscale_constant != 0.0);
}

inline bool asi_fast_path_eligible(const NeFluxFormulation& f, const NeInterfaceProperties& ip) {
  if (!default_roughness_gustiness(f) || !tab_path_eligible(f)) return false;
  const int tf = ip.temperature_formulation;
  return tf == NE_TEMP_BULK || tf == NE_TEMP_SKIN_CONDUCTIVE || tf == NE_TEMP_SKIN_ICE_SNOW;
}

}  // namespace ne
