// ne_flux_generic.cuh — the generic (every plugin variant, every element-type combination) turbulent
// flux solver and its two kernels.  Instantiated per element-type combination in
// ne_flux_generic_{ao,asi}_{f64,f32}.cu so the variants build as separate translation units.
//
// Replaces (citations relative to /root/reference/src/EarthSystemModels/InterfaceComputations/):
//   _compute_atmosphere_ocean_interface_state!    atmosphere_ocean_fluxes.jl:80-197
//   _compute_atmosphere_sea_ice_interface_state!  atmosphere_sea_ice_fluxes.jl:65-185
//   compute_interface_state / iterating           compute_interface_state.jl:5-58
//   iterate_interface_state                       compute_interface_state.jl:69-122
//   iterate_interface_fluxes (similarity theory)  similarity_theory_turbulent_fluxes.jl:315-385
//   iterate_interface_fluxes (coefficient based)  coefficient_based_turbulent_fluxes.jl:346-371
#pragma once

#include "ne_physics.cuh"

namespace ne {

struct SolveFlags {
  int32_t scalar_shared;  // ℓθ ≡ ℓq and ψθ ≡ ψq: evaluate the scalar profile once
  int32_t pad_;
};

// One point's fixed-point solve.  ICE selects the AirIceInterfaceState conventions
// (humidity scalar 0, interface_states.jl:737) vs AirSeaInterfaceState (salinity, :717).
template <class FT, class CT, class VT, bool ICE>
struct InterfaceSolver {
  const NeFluxFormulation& ff;
  const NeInterfaceProperties& ip;
  const NeMediumProperties& medium;
  const Thermo<CT>& th;
  FT g;
  SolveFlags flags;

  // iterate state
  FT ustar, theta_star, q_star, Ts, qs;
  // constants of the point
  AtmosState<FT> a;
  Interior<FT> in;
  RadState<FT> rad;
  FT surf_u, surf_v, surf_S;

  // land interface (AirLandInterfaceState, interface_states.jl:740-777): humidity closure + the land state it reads.
  // Trailing members: aggregate initialisation of the ocean / sea-ice kernels leaves them null / zero.
  const NeLandHumidity* landq;
  FT land_saturation, land_T;
  // NE_ROUGH_LAND / NE_DISPLACEMENT_LAND resolved for this cell (local_roughness_lengths, local_zero_plane_displacement,
  // similarity_theory_turbulent_fluxes.jl:265-303): momentum, temperature, water vapor; displacement.  `land_cell` says the
  // kernel filled them (ocean / sea-ice kernels never carry the land markers: the entry points resolve them on the host).
  FT ell_land[3], d_land;
  int land_cell;

  using WT = decltype(FT() + CT());

  // saturation_specific_humidity interface_states.jl:79-90 (pressure-based, thermodynamics' element type)
  template <class T, class P>
  __device__ __forceinline__ CT saturation_specific_humidity(T Tk, P p_at, int phase) const {
    CT Tc = (CT)Tk, p = (CT)p_at;
    CT pv = th.saturation_vapor_pressure(Tc, phase);
    pv = mn(pv, (CT)0.999 * p);
    return th.eps_inv * pv / (p - (1 - th.eps_inv) * pv);
  }

  // compute_interface_humidity for the land closures: BulkHumidity :120-126, FractionalHumidity :592-598 (+ :151-157),
  // SkinHumidity :625-651 (previous iterate's u★, q★, qˢ)
  __device__ __forceinline__ FT land_humidity() const {
    const NeLandHumidity& h = *landq;
    if (h.kind == NE_LANDQ_BULK) {
      CT qv = saturation_specific_humidity(Ts, a.p, h.phase);
      return (FT)((land_saturation > 0) ? qv : (CT)0);
    }
    if (h.kind == NE_LANDQ_FRACTIONAL_CRITICAL) {
      FT beta = mn(land_saturation / (FT)h.critical_saturation, (FT)1);
      CT qv = saturation_specific_humidity(Ts, a.p, h.phase);
      return (FT)(beta * qv);
    }
    if (h.kind == NE_LANDQ_FRACTIONAL_CONSTANT) {
      CT qv = saturation_specific_humidity(Ts, a.p, h.phase);
      return (FT)(h.efficiency * qv);   // β::Number keeps its own (Float64) type
    }
    if (h.kind == NE_LANDQ_DRY_LAYER) {   // compute_interface_humidity(::DryLayerHumidity) dry_layer_humidity.jl
      auto rho_d = th.air_density(a.T, a.p, a.q);
      const FT S = land_saturation, Tin = Ts, Tla = land_T;
      const FT sc = mn(S / (FT)h.dry_layer_onset_saturation, (FT)1);
      const FT dv = (FT)h.maximum_dry_layer_depth * m_pow(mx((FT)1 - sc, (FT)0), (FT)h.dry_layer_exponent);
      const FT dvmin = (FT)h.minimum_dry_layer_depth, lT = (FT)h.thermal_exchange_depth;
      const FT chi = clampv<FT>(dv / lT, (FT)0, (FT)1);
      const FT Te = Tin + chi * (Tla - Tin);
      CT qe = saturation_specific_humidity(Te, a.p, h.phase);
      const FT theta_l = S * (FT)h.porosity;
      FT Dv;
      if (h.tortuosity == NE_TORTUOSITY_CONSTANT) Dv = (FT)h.molecular_diffusivity;
      else {
        const FT nu = (FT)h.porosity, tg = mx(nu - theta_l, (FT)0);
        Dv = (FT)h.molecular_diffusivity * m_pow(tg, (FT)10 / (FT)3) / (nu * nu);
      }
      auto Ge = rho_d * Dv / mx(dv, dvmin);
      auto Jd = -rho_d * ustar * q_star;
      FT dqd = qs - a.q;
      auto Dd = Ge * dqd + Jd;
      auto qbal = (Ge * qe * dqd + Jd * a.q) / Dd;
      qbal = (Dd == 0) ? (decltype(qbal))qs : qbal;
      CT qinp = saturation_specific_humidity(Tin, a.p, h.phase);
      const FT dvw = (FT)h.wet_transition_width;
      const FT eps_ft = std::is_same<FT, double>::value ? (FT)2.220446049250313e-16 : (FT)1.1920929e-07f;
      const FT z = 10 * (dv - dvmin - dvw / 2) / mx(dvw, eps_ft);
      const FT sigma = 1 / (1 + m_exp(-z));
      return (FT)(qinp + sigma * (qbal - qinp));
    }
    auto rho_a = th.air_density(a.T, a.p, a.q);
    CT qv = saturation_specific_humidity(land_T, a.p, h.phase);
    double gs = h.vapor_diffusivity / h.surface_thickness;
    auto Ja = -rho_a * ustar * q_star;
    FT dq = qs - a.q;
    auto D = gs * dq + Ja;
    auto q = (gs * qv * dq + Ja * a.q) / D;
    return (FT)((D == 0) ? (decltype(q))qs : q);
  }

  // compute_interface_temperature(::SkinTemperature) interface_states.jl:526-577
  __device__ __forceinline__ FT skin_temperature(WT theta_a) const {
    auto rho_a = th.air_density(a.T, a.p, a.q);
    auto c_a = th.cp_m(a.q);
    auto Li = th.latent_heat_sublim(a.T);  // sublimation enthalpy for every surface (:542-544)
    FT Tsm = Ts;
    FT lw_up = rad.sigma * rad.eps * pow4(Tsm);
    FT Qd = -(1 - rad.alpha) * rad.sw - rad.eps * rad.lw;
    auto QT = -rho_a * c_a * ustar * theta_star;
    auto Qv = -rho_a * Li * ustar * q_star;
    const int tf = ip.temperature_formulation;
    if (tf == NE_TEMP_SKIN_DIFFUSIVE || tf == NE_TEMP_SKIN_DIFFUSIVE_INTERIOR) {  // :434-457
      FT kappa = tf == NE_TEMP_SKIN_DIFFUSIVE ? (FT)ip.kappa : mx(in.kappa, (FT)ip.kappa);
      FT delta = (FT)ip.delta;
      FT lambda = 1 / ((FT)medium.reference_density * (FT)medium.heat_capacity);
      auto Qa = Qv + lw_up + Qd;
      auto JT = Qa * lambda;
      auto dT = theta_a - Tsm;
      auto Om = QT * lambda;
      auto D = kappa * dT - Om * delta;
      auto Tstar = (in.T * kappa * dT - (JT * dT + Om * theta_a) * delta) / D;
      using W = decltype(Tstar);
      Tstar = (D == 0) ? (W)Tsm : Tstar;
      W maxdT = (W)(FT)ip.max_dT;
      return (FT)(in.T + clampv<W>(Tstar - in.T, -maxdT, maxdT));
    }
    // conductive_flux_balance_temperature :468-508
    FT R;
    if (tf == NE_TEMP_SKIN_CONDUCTIVE) R = in.hi / (FT)ip.ice_conductivity;
    else R = in.hs / (FT)ip.snow_conductivity + in.hi / (FT)ip.ice_conductivity;
    FT Tb = (FT)medium.liquidus_freshwater_melting_temperature - (FT)medium.liquidus_slope * in.S;
    FT Tm = (FT)medium.liquidus_freshwater_melting_temperature;
    if (medium.temperature_units == NE_DEGREES_CELSIUS) { Tb = Tb + (FT)273.15; Tm = Tm + (FT)273.15; }
    auto dT = theta_a - Tsm;
    auto Qa = Qv + lw_up + Qd;
    using WO = decltype(QT / dT);
    WO Oc = (dT == 0) ? (WO)0 : QT / dT;
    FT beta = 4 * lw_up / Tsm;
    auto D = 1 + beta * R - Oc * R;
    auto Tstar = (Tb + beta * R * Tsm - Oc * R * theta_a - Qa * R) / D;
    using W = decltype(Tstar);
    Tstar = (D == 0) ? (W)Tsm : Tstar;
    Tstar = (Tstar != Tstar) ? (W)Tsm : Tstar;
    W maxdT = (W)ip.max_dT;
    auto Tsp = Tsm + clampv<W>(Tstar - Tsm, -maxdT, maxdT);
    Tsp = mn(Tsp, Tm);
    Tsp = (in.hi >= in.hc) ? Tsp : (W)Tb;
    return (FT)Tsp;
  }

  __device__ __forceinline__ void similarity_step(WT dtheta, FT dq, FT du, FT dv) {
    // buoyancy_scale similarity_theory_turbulent_fluxes.jl:417-425
    auto Tv = th.virtual_temperature(Ts, qs);
    auto bstar = g / Tv * (theta_star * (1 + th.delta * qs) + th.delta * Tv * q_star);
    auto Usg2 = vsgs2<FT>(ff.subgrid_velocities, ustar, bstar, a.h_bl);
    auto U = m_sqrt(sq(du) + sq(dv) + Usg2);

    using LW = decltype(FT() * VT() * U);
    LW lu, lq, lt;
    if (ff.ell_momentum.kind == NE_ROUGH_CONSTANT) lu = (FT)ff.ell_momentum.constant;
    else if (ff.ell_momentum.kind == NE_ROUGH_LAND) lu = ell_land[0];
    else lu = momentum_roughness<FT, VT>(ff.ell_momentum, air_viscosity<FT, VT>(ff.ell_momentum, Ts), ustar, U);
    if (ff.ell_water_vapor.kind == NE_ROUGH_CONSTANT) lq = (FT)ff.ell_water_vapor.constant;
    else if (ff.ell_water_vapor.kind == NE_ROUGH_LAND) lq = ell_land[2];
    else lq = scalar_roughness<FT, VT>(ff.ell_water_vapor, air_viscosity<FT, VT>(ff.ell_water_vapor, Ts), lu, ustar);
    if (flags.scalar_shared) lt = lq;
    else if (ff.ell_temperature.kind == NE_ROUGH_CONSTANT) lt = (FT)ff.ell_temperature.constant;
    else if (ff.ell_temperature.kind == NE_ROUGH_LAND) lt = ell_land[1];
    else lt = scalar_roughness<FT, VT>(ff.ell_temperature, air_viscosity<FT, VT>(ff.ell_temperature, Ts), lu, ustar);

    const FT d_zero = (land_cell && ff.zero_plane_displacement_kind == NE_DISPLACEMENT_LAND) ? d_land : (FT)ff.zero_plane_displacement;
    auto dh = mx(a.z - d_zero, 2 * lu);  // displaced_profile_height :313
    FT kappa = (FT)ff.von_karman_constant;
    using BW = decltype(sq(ustar) / (kappa * bstar));
    BW Lstar = (bstar == 0) ? Inf<BW>::v() : sq(ustar) / (kappa * bstar);

    auto chi_u = kappa / similarity_profile<FT>(ff.similarity_form, ff.psi_momentum, dh, lu, Lstar);
    auto chi_t = kappa / similarity_profile<FT>(ff.similarity_form, ff.psi_temperature, dh, lt, Lstar);
    auto chi_q = flags.scalar_shared ? chi_t
                                     : kappa / similarity_profile<FT>(ff.similarity_form, ff.psi_water_vapor, dh, lq, Lstar);
    ustar = (FT)(chi_u * U);
    theta_star = (FT)(chi_t * dtheta);
    q_star = (FT)(chi_q * dq);
  }

  __device__ __forceinline__ void coefficient_step(WT dtheta, FT dq, FT du, FT dv) {
    using W = WT;
    W Cd, Ch, Cq, dU;
    if (ff.kind == NE_FLUX_LARGE_YEAGER) {  // evaluate_coefficients(::LargeYeagerTransferCoefficients) :288-340
      const NeLargeYeager& ly = ff.large_yeager;
      FT Umin = (FT)ly.neutral_drag.minimum_wind_speed;
      dU = mx(m_sqrt(sq(du) + sq(dv)), Umin);
      FT kap = (FT)ly.von_karman_constant, h0 = (FT)ly.reference_height;
      FT dh = a.z;
      auto Tv = th.virtual_temperature(Ts, qs);
      auto bstar = g / Tv * (theta_star * (1 + th.delta * qs) + th.delta * Tv * q_star);
      using BW = decltype(sq(ustar) / (kap * bstar));
      BW Lstar = (bstar == 0) ? (BW)Inf<FT>::v() : sq(ustar) / (kap * bstar);
      auto zeta = dh / Lstar;
      auto psi_m = stability_profile<FT>(ly.psi_momentum, zeta);
      auto psi_h = stability_profile<FT>(ly.psi_temperature, zeta);
      W Cdp = (ustar == 0) ? (W)polynomial_drag<FT>(ly.neutral_drag, dU) : (W)(sq(ustar) / sq(dU));
      W lg = m_log(dh / h0);
      W UN10 = dU / (1 + m_sqrt(Cdp) / kap * (lg - psi_m));
      UN10 = mx(UN10, Umin);
      W CdN = polynomial_drag<FT>(ly.neutral_drag, UN10);
      W rCdN = m_sqrt(CdN);
      W ChN = rCdN / 1000 * ((zeta > 0) ? (FT)ly.stable_heat : (FT)ly.unstable_heat);
      W CqN = rCdN / 1000 * (FT)ly.moisture;
      W xi_m = rCdN / kap * (lg - psi_m);
      Cd = CdN / sq(1 + xi_m);
      W xi_h = rCdN / kap * (lg - psi_h);
      W ratio = m_sqrt(Cd) / rCdN;
      Ch = ChN * ratio / (1 + ChN * xi_h);
      Cq = CqN * ratio / (1 + CqN * xi_h);
    } else {
      dU = mx(m_sqrt(sq(du) + sq(dv)), (FT)0);
      W c[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const NeTransferCoefficient& tc = ff.coefficients[k];
        c[k] = tc.kind == NE_COEFF_CONSTANT ? (W)(FT)tc.constant : (W)polynomial_drag<FT>(tc.polynomial, dU);
      }
      Cd = c[0]; Ch = c[1]; Cq = c[2];
    }
    W rCd = m_sqrt(Cd);
    ustar = (FT)(rCd * dU);
    theta_star = (Cd == 0) ? (FT)0 : (FT)(Ch / rCd * dtheta);
    q_star = (Cd == 0) ? (FT)0 : (FT)(Cq / rCd * dq);
  }

  // compute_interface_state (compute_interface_state.jl:31-58); returns the iteration count
  __device__ __forceinline__ int solve() {
    const bool bulk = ip.temperature_formulation == NE_TEMP_BULK;
    const bool fixed = ff.stop.kind == NE_STOP_FIXED_ITERATIONS;
    const int maxiter = ff.stop.maxiter;
    const FT tol = (FT)ff.stop.tolerance;
    // iteration invariants
    const WT theta_a = a.T + g * a.z / th.cp_m(a.q);  // surface_atmosphere_temperature interface_states.jl:308-317
    FT du, dv;
    if (ip.velocity_formulation == NE_VEL_RELATIVE) { du = a.u - surf_u; dv = a.v - surf_v; } else { du = a.u; dv = a.v; }
    WT dtheta = 0;
    FT dq = 0;
    int it = 0;
    FT drift = 0;
    for (;;) {
      bool go = fixed ? (it < maxiter) : (!((drift < tol) | (it >= maxiter)) | (it == 0));
      if (!go) break;
      if (!bulk) Ts = skin_temperature(theta_a);
      if (!bulk || it == 0 || (landq && (landq->kind == NE_LANDQ_SKIN || landq->kind == NE_LANDQ_DRY_LAYER))) {
        qs = landq ? land_humidity() : surface_specific_humidity<FT, CT>(ip, th, a.p, Ts, ICE ? (FT)0 : surf_S);
        dq = a.q - qs;
        dtheta = theta_a - Ts;
      }
      FT pu = ustar, pt = theta_star, pq = q_star;
      if (ff.kind == NE_FLUX_SIMILARITY_THEORY) similarity_step(dtheta, dq, du, dv);
      else coefficient_step(dtheta, dq, du, dv);
      drift = m_abs(ustar - pu) + m_abs(theta_star - pt) + m_abs(q_star - pq);
      ++it;
    }
    return it;
  }
};

// ---- atmosphere–ocean kernel -----------------------------------------------------------------------
template <class FT, class CT, class VT>
__global__ void __launch_bounds__(128)
ao_flux_kernel(const __grid_constant__ NeAtmosOceanDesc d, const __grid_constant__ Layout L,
               const __grid_constant__ Thermo<CT> th, const __grid_constant__ SolveFlags flags) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)L.ni * L.nj) return;
  const int32_t jj = (int32_t)(t / L.ni);
  const int32_t i = L.i_lo + (int32_t)(t - (int64_t)jj * L.ni);
  const int32_t j = L.j_lo + jj;
  const int64_t idx = L.at(i, j);

  InterfaceSolver<FT, CT, VT, false> s{d.flux, d.properties, d.ocean, th, (FT)d.gravitational_acceleration, flags};
  s.a.u = __ldg((const FT*)d.ua + idx);
  s.a.v = __ldg((const FT*)d.va + idx);
  s.a.T = __ldg((const FT*)d.Ta + idx);
  s.a.p = __ldg((const FT*)d.pa + idx);
  s.a.q = __ldg((const FT*)d.qa + idx);
  s.a.z = slot_at<FT>(d.surface_layer_height, idx);
  s.a.h_bl = slot_at<FT>(d.boundary_layer_height, idx);
  // ℑxᶜᵃᵃ u, ℑyᵃᶜᵃ v (atmosphere_ocean_fluxes.jl:64-65)
  s.in.u = d.uo.ptr ? (slot_at<FT>(d.uo, idx) + slot_at<FT>(d.uo, idx + 1)) / 2 : (FT)d.uo.value;
  s.in.v = d.vo.ptr ? (slot_at<FT>(d.vo, idx) + slot_at<FT>(d.vo, idx + L.sx)) / 2 : (FT)d.vo.value;
  const bool celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  FT To = slot_at<FT>(d.To, idx);
  if (celsius) To = To + (FT)273.15;
  s.in.T = To;
  s.in.S = slot_at<FT>(d.So, idx);
  s.in.kappa = 0; s.in.hi = 0; s.in.hs = 0; s.in.hc = 0;
  if (d.properties.temperature_formulation == NE_TEMP_SKIN_DIFFUSIVE_INTERIOR) s.in.kappa = __ldg((const FT*)d.kappa + idx);
  s.rad = radiation_state<FT>(d.radiation, L, idx, j);
  s.surf_u = s.in.u; s.surf_v = s.in.v; s.surf_S = s.in.S;

  const bool not_water = d.inactive ? (d.inactive[idx] != 0) : false;
  const bool needs_to_converge = d.flux.stop.kind == NE_STOP_CONVERGENCE;

  // initial interface state (:131-137)
  s.ustar = s.theta_star = s.q_star = (FT)1e-4;
  s.Ts = To;
  s.qs = 0;
  int iters = 0;
  if (!(needs_to_converge && not_water)) iters = s.solve();
  FT ustar = s.ustar, theta_star = s.theta_star, q_star = s.q_star, Ts = s.Ts;
  FT su = s.surf_u, sv = s.surf_v;
  if (not_water) {  // zero_interface_state interface_states.jl:800-803, applied at :145,:158
    ustar = 0; theta_star = 0; q_star = 0; Ts = (FT)273.15; su = 0; sv = 0;
  }
  FT du, dv;
  if (d.properties.velocity_formulation == NE_VEL_RELATIVE) { du = s.a.u - su; dv = s.a.v - sv; } else { du = s.a.u; dv = s.a.v; }
  FluxEpilogue<FT, CT> e(th, s.a, ustar, theta_star, q_star, du, dv, false);
  ((FT*)d.latent_heat)[idx] = e.Qv;
  ((FT*)d.sensible_heat)[idx] = e.Qc;
  ((FT*)d.water_vapor)[idx] = e.Jv;
  ((FT*)d.x_momentum)[idx] = e.tx;
  ((FT*)d.y_momentum)[idx] = e.ty;
  ((FT*)d.interface_temperature)[idx] = celsius ? Ts - (FT)273.15 : Ts;
  ((FT*)d.friction_velocity)[idx] = ustar;
  ((FT*)d.temperature_scale)[idx] = theta_star;
  ((FT*)d.water_vapor_scale)[idx] = q_star;
  if (d.iterations) d.iterations[idx] = iters;
}

// ---- atmosphere–sea-ice kernel ----------------------------------------------------------------------
template <class FT, class CT, class VT>
__global__ void __launch_bounds__(128)
asi_flux_kernel(const __grid_constant__ NeAtmosSeaIceDesc d, const __grid_constant__ Layout L,
                const __grid_constant__ Thermo<CT> th, const __grid_constant__ SolveFlags flags) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)L.ni * L.nj) return;
  const int32_t jj = (int32_t)(t / L.ni);
  const int32_t i = L.i_lo + (int32_t)(t - (int64_t)jj * L.ni);
  const int32_t j = L.j_lo + jj;
  const int64_t idx = L.at(i, j);

  InterfaceSolver<FT, CT, VT, true> s{d.flux, d.properties, d.sea_ice, th, (FT)d.gravitational_acceleration, flags};
  s.a.u = __ldg((const FT*)d.ua + idx);
  s.a.v = __ldg((const FT*)d.va + idx);
  s.a.T = __ldg((const FT*)d.Ta + idx);
  s.a.p = __ldg((const FT*)d.pa + idx);
  s.a.q = __ldg((const FT*)d.qa + idx);
  s.a.z = slot_at<FT>(d.surface_layer_height, idx);
  s.a.h_bl = slot_at<FT>(d.boundary_layer_height, idx);
  const bool ocean_celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  const bool ice_celsius = d.sea_ice.temperature_units == NE_DEGREES_CELSIUS;
  FT To = slot_at<FT>(d.To, idx);
  if (ocean_celsius) To = To + (FT)273.15;
  s.in.u = 0; s.in.v = 0; s.in.T = To;   // ice velocity forced to 0 (:97-98)
  s.in.S = slot_at<FT>(d.So, idx);
  s.in.kappa = 0;
  s.in.hi = slot_at<FT>(d.hi, idx);
  s.in.hs = slot_at<FT>(d.hs, idx);
  s.in.hc = slot_at<FT>(d.hc, idx);
  const FT conc = slot_at<FT>(d.concentration, idx);
  FT* Tsurf = (FT*)d.interface_temperature;
  FT Ts0 = Tsurf[idx];
  if (ice_celsius) Ts0 = Ts0 + (FT)273.15;
  s.rad = radiation_state<FT>(d.radiation, L, idx, j);
  s.surf_u = 0; s.surf_v = 0; s.surf_S = 0;

  const bool not_water = d.inactive ? (d.inactive[idx] != 0) : false;
  const bool needs_to_converge = d.flux.stop.kind == NE_STOP_CONVERGENCE;
  const bool ice_free = conc == 0;

  s.ustar = s.theta_star = s.q_star = (FT)1e-4f;   // convert(FT, 1f-4) :127
  s.Ts = Ts0;
  s.qs = 0;
  int iters = 0;
  FT ustar, theta_star, q_star, Ts;
  if ((needs_to_converge && not_water) || ice_free) {   // :141-142
    ustar = 0; theta_star = 0; q_star = 0; Ts = To;
  } else {
    iters = s.solve();
    ustar = s.ustar; theta_star = s.theta_star; q_star = s.q_star; Ts = s.Ts;
  }
  FluxEpilogue<FT, CT> e(th, s.a, ustar, theta_star, q_star, s.a.u, s.a.v, true);  // Δu = uₐ - 0 for both formulations
  ((FT*)d.latent_heat)[idx] = e.Qv;
  ((FT*)d.sensible_heat)[idx] = e.Qc;
  ((FT*)d.water_vapor)[idx] = e.Jv;
  ((FT*)d.x_momentum)[idx] = e.tx;
  ((FT*)d.y_momentum)[idx] = e.ty;
  Tsurf[idx] = ice_celsius ? Ts - (FT)273.15 : Ts;
  if (d.iterations) d.iterations[idx] = iters;
}

// ---- atmosphere–land kernel (_compute_atmosphere_land_interface_state!, atmosphere_land_fluxes.jl:147-251) -----------
template <class FT, class CT, class VT>
__global__ void __launch_bounds__(128)
al_flux_kernel(const __grid_constant__ NeAtmosLandDesc d, const __grid_constant__ Layout L,
               const __grid_constant__ Thermo<CT> th, const __grid_constant__ SolveFlags flags,
               const __grid_constant__ NeMediumProperties medium) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)L.ni * L.nj) return;
  const int32_t jj = (int32_t)(t / L.ni);
  const int64_t idx = L.at(L.i_lo + (int32_t)(t - (int64_t)jj * L.ni), L.j_lo + jj);

  InterfaceSolver<FT, CT, VT, false> s{d.flux, d.properties, medium, th, (FT)d.gravitational_acceleration, flags};
  s.a.u = __ldg((const FT*)d.ua + idx);
  s.a.v = __ldg((const FT*)d.va + idx);
  s.a.T = __ldg((const FT*)d.Ta + idx);
  s.a.p = __ldg((const FT*)d.pa + idx);
  s.a.q = __ldg((const FT*)d.qa + idx);
  s.a.z = slot_at<FT>(d.surface_layer_height, idx);
  s.a.h_bl = slot_at<FT>(d.boundary_layer_height, idx);
  const FT Tland = slot_at<FT>(d.land_temperature, idx);
  s.in.u = 0; s.in.v = 0; s.in.T = Tland; s.in.S = 0;      // surface velocities are zero for land (:189-191)
  s.in.kappa = 0; s.in.hi = 0; s.in.hs = 0; s.in.hc = 0;
  s.rad = RadState<FT>{0, 0, 0, 0, 0};
  s.surf_u = 0; s.surf_v = 0; s.surf_S = 0;
  s.landq = &d.humidity;
  s.land_saturation = slot_at<FT>(d.saturation, idx);
  s.land_T = Tland;
  {   // local_atmosphere_land_surface_properties: what the land model provides for this cell
    const NeRoughnessLength* r[3] = {&d.flux.ell_momentum, &d.flux.ell_temperature, &d.flux.ell_water_vapor};
    const FT* field[3] = {(const FT*)d.momentum_roughness_length, (const FT*)d.scalar_roughness_length, (const FT*)d.scalar_roughness_length};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const FT lmin = (FT)r[k]->land_minimum_roughness_length;
      const FT candidate = field[k] ? mx(__ldg(field[k] + idx), lmin) : lmin;
      s.ell_land[k] = mx((FT)r[k]->land_multiplier * candidate, lmin);
    }
    s.d_land = d.zero_plane_displacement ? __ldg((const FT*)d.zero_plane_displacement + idx) : (FT)0;
    s.land_cell = 1;
  }
  s.ustar = s.theta_star = s.q_star = (FT)1e-4;              // convert(FT, 1e-4) :204
  s.Ts = Tland;
  s.qs = (FT)s.saturation_specific_humidity(Tland, s.a.p, d.humidity.phase);   // :205
  const int iters = s.solve();
  FT du, dv;
  if (d.properties.velocity_formulation == NE_VEL_RELATIVE) { du = s.a.u - s.surf_u; dv = s.a.v - s.surf_v; } else { du = s.a.u; dv = s.a.v; }
  FluxEpilogue<FT, CT> e(th, s.a, s.ustar, s.theta_star, s.q_star, du, dv, false);
  ((FT*)d.latent_heat)[idx] = e.Qv;
  ((FT*)d.sensible_heat)[idx] = e.Qc;
  ((FT*)d.water_vapor)[idx] = e.Jv;
  ((FT*)d.x_momentum)[idx] = e.tx;
  ((FT*)d.y_momentum)[idx] = e.ty;
  ((FT*)d.interface_temperature)[idx] = s.Ts;
  ((FT*)d.friction_velocity)[idx] = s.ustar;
  ((FT*)d.temperature_scale)[idx] = s.theta_star;
  ((FT*)d.water_vapor_scale)[idx] = s.q_star;
  if (d.iterations) d.iterations[idx] = iters;
}

inline SolveFlags make_flags(const NeFluxFormulation& f) {
  SolveFlags fl = {0, 0};
  if (f.kind == NE_FLUX_SIMILARITY_THEORY &&
      std::memcmp(&f.ell_temperature, &f.ell_water_vapor, sizeof(NeRoughnessLength)) == 0 &&
      std::memcmp(&f.psi_temperature, &f.psi_water_vapor, sizeof(NeStabilityProfile)) == 0)
    fl.scalar_shared = 1;
  return fl;
}


template <class FT, class CT, class VT>
int launch_ao(const NeAtmosOceanDesc& d, cudaStream_t stream) {
  Layout L = make_layout(d.grid);
  Thermo<CT> th = Thermo<CT>::make(d.thermo);
  SolveFlags fl = make_flags(d.flux);
  const int64_t n = (int64_t)L.ni * L.nj;
  const int threads = 128;
  const int64_t blocks = (n + threads - 1) / threads;
  ao_flux_kernel<FT, CT, VT><<<(unsigned)blocks, threads, 0, stream>>>(d, L, th, fl);
  NE_CUDA_CHECK_LAUNCH("ne_atmosphere_ocean_fluxes");
  return NE_OK;
}

template <class FT, class CT, class VT>
int launch_asi(const NeAtmosSeaIceDesc& d, cudaStream_t stream) {
  Layout L = make_layout(d.grid);
  Thermo<CT> th = Thermo<CT>::make(d.thermo);
  SolveFlags fl = make_flags(d.flux);
  const int64_t n = (int64_t)L.ni * L.nj;
  const int threads = 128;
  const int64_t blocks = (n + threads - 1) / threads;
  asi_flux_kernel<FT, CT, VT><<<(unsigned)blocks, threads, 0, stream>>>(d, L, th, fl);
  NE_CUDA_CHECK_LAUNCH("ne_atmosphere_sea_ice_fluxes");
  return NE_OK;
}

template <class FT, class CT, class VT>
int launch_al(const NeAtmosLandDesc& d, cudaStream_t stream) {
  Layout L = make_layout(d.grid);
  Thermo<CT> th = Thermo<CT>::make(d.thermo);
  SolveFlags fl = make_flags(d.flux);
  NeMediumProperties medium;
  std::memset(&medium, 0, sizeof(medium));   // only read by skin temperatures, which the land interface does not use
  const int64_t n = (int64_t)L.ni * L.nj;
  const int threads = 128;
  const int64_t blocks = (n + threads - 1) / threads;
  al_flux_kernel<FT, CT, VT><<<(unsigned)blocks, threads, 0, stream>>>(d, L, th, fl, medium);
  NE_CUDA_CHECK_LAUNCH("ne_atmosphere_land_fluxes");
  return NE_OK;
}

}  // namespace ne
