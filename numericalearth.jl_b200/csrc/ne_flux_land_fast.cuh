// ne_flux_land_fast.cuh — atmosphere–land turbulent fluxes, default plugin tree, on the work-queue kernel.
//
// Replaces _compute_atmosphere_land_interface_state! (atmosphere_land_fluxes.jl:147-251) for the reference's land defaults
// (component_interfaces.jl:501-521): SimilarityTheoryFluxes with constant roughness lengths (0.1 / 0.01 / 0.01 m), any
// tabulatable stability-function pair (default: the Large–Yeager set), convective gustiness, BulkTemperature, and the
// humidity closures BulkHumidity / FractionalHumidity (invariants of the iteration) or SkinHumidity (re-solved every trip,
// interface_states.jl:625-651), Float64 exchange grid.  Everything else keeps the generic kernel.  Every `:xy` cell is
// solved (the reference applies no mask here); the work-queue kernel evens out the trip counts.
#pragma once

#include "ne_flux_queue.cuh"

namespace ne {

struct LandPoint {
  FastPoint f;                 // invariants of the similarity step + the iterate (u★, θ★, q★)
  double Ts, ap, aq;           // interface (= bulk land) temperature, atmosphere pressure and humidity
  // SkinHumidity / DryLayerHumidity (vapor-flux balances re-solved every trip): air density, conductance G (κ/d, or
  // ρₐ Dᵛ / max(δᵛ, δᵛmin)), source humidity (reservoir / evaporation front), saturated-skin value and logistic weight
  // (SkinHumidity: q⁺ unused, σ = 1)
  double rho_a, G, q_src, q_sat_skin, sigma;
  double qs;                   // surface specific humidity (carried from trip to trip by the balances)
};

template <class CT>
__device__ __noinline__ void al_write_outputs(const NeAtmosLandDesc& d, const Thermo<CT>& th, int64_t idx,
                                              double ustar, double theta_star, double q_star, double Ts, int iters) {
  using FT = double;
  AtmosState<FT> a;
  a.u = __ldg((const FT*)d.ua + idx);
  a.v = __ldg((const FT*)d.va + idx);
  a.T = __ldg((const FT*)d.Ta + idx);
  a.p = __ldg((const FT*)d.pa + idx);
  a.q = __ldg((const FT*)d.qa + idx);
  a.z = 0; a.h_bl = 0;
  FluxEpilogue<FT, CT> e(th, a, ustar, theta_star, q_star, a.u, a.v, false);   // surface at rest: Δu = uₐ for both velocity formulations
  ((FT*)d.latent_heat)[idx] = e.Qv;
  ((FT*)d.sensible_heat)[idx] = e.Qc;
  ((FT*)d.water_vapor)[idx] = e.Jv;
  ((FT*)d.x_momentum)[idx] = e.tx;
  ((FT*)d.y_momentum)[idx] = e.ty;
  ((FT*)d.interface_temperature)[idx] = Ts;
  ((FT*)d.friction_velocity)[idx] = ustar;
  ((FT*)d.temperature_scale)[idx] = theta_star;
  ((FT*)d.water_vapor_scale)[idx] = q_star;
  if (d.iterations) d.iterations[idx] = iters;
}

template <class CT, bool HS>
struct LandProblem {
  using FT = double;
  using Point = LandPoint;
  static constexpr int NSTATE = 4;
  struct Params {
    NeAtmosLandDesc d;
    Layout L;
    Thermo<CT> th;
    FastParams P;
    TabParams T;
  };
  __device__ static __forceinline__ const Layout& layout(const Params& p) { return p.L; }
  __device__ static __forceinline__ const FastParams& fast(const Params& p) { return p.P; }
  __device__ static __forceinline__ FT tolerance(const Params& p) { return p.P.tol; }

  // saturation_specific_humidity interface_states.jl:79-90
  __device__ static __forceinline__ CT qsat(const Params& p, FT T, FT pa) {
    CT Tc = (CT)T, pc = (CT)pa;
    CT pv = p.th.saturation_vapor_pressure(Tc, p.d.humidity.phase);
    pv = mn(pv, (CT)0.999 * pc);
    return p.th.eps_inv * pv / (pc - (1 - p.th.eps_inv) * pv);
  }
  __device__ static __forceinline__ void set_humidity(const Params& p, Point& s, FT qs) {
    const FT Tv = p.th.virtual_temperature(s.Ts, qs);
    s.qs = qs;
    s.f.gTv = p.P.g / Tv;
    s.f.c1 = 1 + p.th.delta * qs;
    s.f.c2 = p.th.delta * Tv;
    s.f.dq = s.aq - qs;
  }

  __device__ static __forceinline__ bool admit(const Params& p, int32_t idx) {
    if (p.P.fixed && p.P.maxiter <= 0) {   // no trip: the initial state (:204-210)
      al_write_outputs<CT>(p.d, p.th, idx, 1e-4, 1e-4, 1e-4, slot_at<FT>(p.d.land_temperature, idx), 0);
      return false;
    }
    return true;
  }

  __device__ static __forceinline__ void prologue(const Params& p, int32_t idx, Point& s, bool fresh) {
    const NeAtmosLandDesc& d = p.d;
    const FT au = __ldg((const FT*)d.ua + idx), av = __ldg((const FT*)d.va + idx);
    const FT aT = __ldg((const FT*)d.Ta + idx), ap = __ldg((const FT*)d.pa + idx), aq = __ldg((const FT*)d.qa + idx);
    const FT az = HS ? (FT)d.surface_layer_height.value : slot_at<FT>(d.surface_layer_height, idx);
    s.Ts = slot_at<FT>(d.land_temperature, idx);
    s.ap = ap; s.aq = aq;
    s.f.dudv2 = au * au + av * av;
    s.f.h_bl = HS ? (FT)d.boundary_layer_height.value : slot_at<FT>(d.boundary_layer_height, idx);
    s.f.hd = az - p.P.d_zero;
    s.f.log_hd = HS ? p.T.log_hd : log(s.f.hd);
    s.f.dtheta = (aT + p.P.g * az / p.th.cp_m(aq)) - s.Ts;
    const NeLandHumidity& h = d.humidity;
    if (h.kind == NE_LANDQ_SKIN || h.kind == NE_LANDQ_DRY_LAYER) {
      s.rho_a = p.th.air_density(aT, ap, aq);
      if (h.kind == NE_LANDQ_SKIN) {
        s.G = h.vapor_diffusivity / h.surface_thickness;
        s.q_src = qsat(p, s.Ts, ap);                    // the reservoir sits at the bulk land temperature (= T_s under BulkTemperature)
        s.q_sat_skin = 0; s.sigma = 1;
      } else {   // dry_layer_humidity.jl: under BulkTemperature T_in = T_la, so everything but the balance itself is invariant
        const FT S = slot_at<FT>(d.saturation, idx), Tin = s.Ts, Tla = s.Ts;
        const FT sc = mn(S / (FT)h.dry_layer_onset_saturation, (FT)1);
        const FT dv = (FT)h.maximum_dry_layer_depth * pow(mx((FT)1 - sc, (FT)0), (FT)h.dry_layer_exponent);
        const FT dvmin = (FT)h.minimum_dry_layer_depth;
        const FT chi = clampv<FT>(dv / (FT)h.thermal_exchange_depth, (FT)0, (FT)1);
        const FT Te = Tin + chi * (Tla - Tin);
        s.q_src = qsat(p, Te, ap);
        const FT theta_l = S * (FT)h.porosity;
        FT Dv = (FT)h.molecular_diffusivity;
        if (h.tortuosity != NE_TORTUOSITY_CONSTANT) {
          const FT nu = (FT)h.porosity, tg = mx(nu - theta_l, (FT)0);
          Dv = (FT)h.molecular_diffusivity * pow(tg, (FT)10 / (FT)3) / (nu * nu);
        }
        s.G = s.rho_a * Dv / mx(dv, dvmin);
        s.q_sat_skin = qsat(p, Tin, ap);
        const FT dvw = (FT)h.wet_transition_width;
        const FT z = 10 * (dv - dvmin - dvw / 2) / mx(dvw, (FT)2.220446049250313e-16);
        s.sigma = 1 / (1 + exp(-z));
      }
      if (fresh) set_humidity(p, s, (FT)qsat(p, s.Ts, ap));   // initial qₛ (:205); replaced before the first similarity step
    } else {
      const FT S = slot_at<FT>(d.saturation, idx);
      const CT qv = qsat(p, s.Ts, ap);
      FT qs;
      if (h.kind == NE_LANDQ_BULK) qs = (FT)((S > 0) ? qv : (CT)0);
      else if (h.kind == NE_LANDQ_FRACTIONAL_CRITICAL) qs = (FT)(mn(S / (FT)h.critical_saturation, (FT)1) * qv);
      else qs = (FT)(h.efficiency * qv);
      set_humidity(p, s, qs);
      s.rho_a = 0; s.G = 0; s.q_src = 0; s.q_sat_skin = 0; s.sigma = 0;
    }
    if (fresh) s.f.ustar = s.f.theta_star = s.f.q_star = 1e-4;   // convert(FT, 1e-4) :204
  }
  __device__ static __forceinline__ void get_state(const Point& s, FT* v) { v[0] = s.f.ustar; v[1] = s.f.theta_star; v[2] = s.f.q_star; v[3] = s.qs; }
  __device__ static __forceinline__ void set_state(Point& s, const FT* v) { s.f.ustar = v[0]; s.f.theta_star = v[1]; s.f.q_star = v[2]; s.qs = v[3]; }

  __device__ static __forceinline__ FT trip(const Params& p, const double* tab, Point& s, int) {
    const NeLandHumidity& h = p.d.humidity;
    if (h.kind == NE_LANDQ_SKIN || h.kind == NE_LANDQ_DRY_LAYER) {
      // compute_interface_humidity(::SkinHumidity) :625-651 / (::DryLayerHumidity) with the previous iterate
      const FT Ja = -s.rho_a * s.f.ustar * s.f.q_star;
      const FT dq = s.qs - s.aq;
      const FT D = s.G * dq + Ja;
      const FT qbal = (D == 0) ? s.qs : (s.G * s.q_src * dq + Ja * s.aq) / D;
      set_humidity(p, s, h.kind == NE_LANDQ_SKIN ? qbal : s.q_sat_skin + s.sigma * (qbal - s.q_sat_skin));
    }
    const FT pu = s.f.ustar, pt = s.f.theta_star, pq = s.f.q_star;
    tab_iteration<true>(p.P, p.T, tab, s.f, p.T.general_psi ? &p.d.flux : nullptr);
    return fabs(s.f.ustar - pu) + fabs(s.f.theta_star - pt) + fabs(s.f.q_star - pq);
  }

  __device__ static __forceinline__ void finish(const Params& p, int32_t idx, const Point& s, int it) {
    al_write_outputs<CT>(p.d, p.th, idx, s.f.ustar, s.f.theta_star, s.f.q_star, s.Ts, it);
  }
};

inline bool land_fast_path_eligible(const NeFluxFormulation& f, const NeInterfaceProperties& ip) {
  if (f.kind != NE_FLUX_SIMILARITY_THEORY || f.similarity_form != NE_PROFILE_LOGARITHMIC) return false;
  if (ip.temperature_formulation != NE_TEMP_BULK) return false;
  if (std::memcmp(&f.psi_temperature, &f.psi_water_vapor, sizeof(NeStabilityProfile)) != 0) return false;
  if (f.ell_momentum.kind != NE_ROUGH_CONSTANT || f.ell_temperature.kind != NE_ROUGH_CONSTANT ||
      f.ell_water_vapor.kind != NE_ROUGH_CONSTANT) return false;
  if (f.zero_plane_displacement_kind != NE_DISPLACEMENT_CONSTANT) return false;   // per-cell displacement: generic kernel
  if (f.ell_temperature.constant != f.ell_water_vapor.constant) return false;
  if (!(f.ell_momentum.constant > 0) || !(f.ell_temperature.constant > 0)) return false;
  const NeSubgridVelocity& g = f.subgrid_velocities;
  if (g.convective_kind != NE_SGS_CONVECTIVE || !(g.minimum_gustiness > 0)) return false;
  if (g.composite && g.mesoscale_kind != NE_SGS_NONE && g.mesoscale_kind != NE_SGS_CONSTANT) return false;
  return tab_path_eligible(f);
}

}  // namespace ne
