// ne_flux_kernels.cu — atmosphere–ocean and atmosphere–sea-ice turbulent flux kernels (sm_100a).
//
// Replaces (citations relative to /root/reference/src/EarthSystemModels/InterfaceComputations/):
//   _compute_atmosphere_ocean_interface_state!    atmosphere_ocean_fluxes.jl:80-197
//   _compute_atmosphere_sea_ice_interface_state!  atmosphere_sea_ice_fluxes.jl:65-185
//   compute_interface_state / iterating           compute_interface_state.jl:5-58
//   iterate_interface_state                       compute_interface_state.jl:69-122
//   iterate_interface_fluxes (similarity theory)  similarity_theory_turbulent_fluxes.jl:315-385
//   iterate_interface_fluxes (coefficient based)  coefficient_based_turbulent_fluxes.jl:346-371
//
// Design: one thread per exchange point, consecutive threads along x (coalesced SoA loads of the
// surface state through the read-only path); the fixed-point iterate (u★, θ★, q★, Tₛ, qₛ) lives in
// registers; each lane leaves the loop at exactly the iteration the reference's `iterating`
// predicate stops at, the warp reconverges after its slowest lane.  Iteration-invariant terms of
// the BulkTemperature variant (qₛ, Δq, θₐ, Δθ, 𝒯ₛ, g/𝒯ₛ, Δu, Δv) are hoisted.  No tensor cores:
// nothing here is a contraction.  FP64 transcendentals are libdevice (no fast-math).
#include <algorithm>
#include <cstdlib>

#include <mutex>
#include <vector>

#include "ne_flux_fast.cuh"
#include "ne_flux_tab.cuh"
#include "ne_physics.cuh"

namespace ne {

// generic kernels: defined in ne_flux_generic.cuh, instantiated in ne_flux_generic_*.cu
template <class FT, class CT, class VT> int launch_ao(const NeAtmosOceanDesc& d, cudaStream_t stream);
template <class FT, class CT, class VT> int launch_asi(const NeAtmosSeaIceDesc& d, cudaStream_t stream);

// ---- atmosphere–ocean kernel, default plugin tree (Float64), see ne_flux_fast.cuh --------------------
template <class CT, int MINB>
__global__ void __launch_bounds__(128, MINB)
ao_flux_fast_kernel(const __grid_constant__ NeAtmosOceanDesc d, const __grid_constant__ Layout L,
                    const __grid_constant__ Thermo<CT> th, const __grid_constant__ FastParams P) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)L.ni * L.nj) return;
  const int32_t jj = (int32_t)(t / L.ni);
  const int32_t i = L.i_lo + (int32_t)(t - (int64_t)jj * L.ni);
  const int32_t j = L.j_lo + jj;
  const int64_t idx = L.at(i, j);
  using FT = double;

  AtmosState<FT> a;
  a.u = __ldg((const FT*)d.ua + idx);
  a.v = __ldg((const FT*)d.va + idx);
  a.T = __ldg((const FT*)d.Ta + idx);
  a.p = __ldg((const FT*)d.pa + idx);
  a.q = __ldg((const FT*)d.qa + idx);
  a.z = slot_at<FT>(d.surface_layer_height, idx);
  a.h_bl = slot_at<FT>(d.boundary_layer_height, idx);
  FT uo = d.uo.ptr ? (slot_at<FT>(d.uo, idx) + slot_at<FT>(d.uo, idx + 1)) / 2 : (FT)d.uo.value;
  FT vo = d.vo.ptr ? (slot_at<FT>(d.vo, idx) + slot_at<FT>(d.vo, idx + L.sx)) / 2 : (FT)d.vo.value;
  const bool celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  FT To = slot_at<FT>(d.To, idx);
  if (celsius) To = To + 273.15;
  const FT So = slot_at<FT>(d.So, idx);
  const bool not_water = d.inactive ? (d.inactive[idx] != 0) : false;
  const bool skip = not_water && !P.fixed;   // needs_to_converge && not_water (:144)

  FT ustar = 0, theta_star = 0, q_star = 0, Ts = To;
  int iters = 0;
  FT du, dv;
  if (d.properties.velocity_formulation == NE_VEL_RELATIVE) { du = a.u - uo; dv = a.v - vo; } else { du = a.u; dv = a.v; }
  if (!skip) {
    FastPoint s;
    const FT qs = surface_specific_humidity<FT, CT>(d.properties, th, a.p, To, So);
    const FT Tv = th.virtual_temperature(To, qs);
    s.gTv = P.g / Tv;
    s.c1 = 1 + th.delta * qs;
    s.c2 = th.delta * Tv;
    s.dudv2 = du * du + dv * dv;
    s.h_bl = a.h_bl;
    s.hd = a.z - P.d_zero;
    s.log_hd = log(s.hd);
    s.dtheta = (a.T + P.g * a.z / th.cp_m(a.q)) - To;
    s.dq = a.q - qs;
    s.ustar = s.theta_star = s.q_star = 1e-4;
    iters = fast_solve(P, s);
    ustar = s.ustar; theta_star = s.theta_star; q_star = s.q_star;
  }
  if (not_water) {  // zero_interface_state (interface_states.jl:800-803)
    ustar = 0; theta_star = 0; q_star = 0; Ts = 273.15;
    if (d.properties.velocity_formulation == NE_VEL_RELATIVE) { du = a.u; dv = a.v; }
  }
  FluxEpilogue<FT, CT> e(th, a, ustar, theta_star, q_star, du, dv, false);
  ((FT*)d.latent_heat)[idx] = e.Qv;
  ((FT*)d.sensible_heat)[idx] = e.Qc;
  ((FT*)d.water_vapor)[idx] = e.Jv;
  ((FT*)d.x_momentum)[idx] = e.tx;
  ((FT*)d.y_momentum)[idx] = e.ty;
  ((FT*)d.interface_temperature)[idx] = celsius ? Ts - 273.15 : Ts;
  ((FT*)d.friction_velocity)[idx] = ustar;
  ((FT*)d.temperature_scale)[idx] = theta_star;
  ((FT*)d.water_vapor_scale)[idx] = q_star;
  if (d.iterations) d.iterations[idx] = iters;
}

// ---- atmosphere–ocean kernel, default plugin tree, table-driven iteration (ne_flux_tab.cuh) -----------
// 256-thread CTAs; each CTA stages the 27 KB solver table (log table + ψ polynomials) in shared memory once
// and then walks 256-point tiles with a grid stride, so the staging cost is amortised over several tiles.
// Register budget: only the 9 per-point invariants and the 3-component iterate live across the loop; the
// atmosphere state needed by the flux epilogue is re-read from global memory (L2 hits) after the solve.
// HS: surface-layer and boundary-layer heights are scalars (PrescribedAtmosphere default) → uniform.
template <class CT, int MINB, bool HS>
__global__ void __launch_bounds__(256, MINB)
ao_flux_tab_kernel(const __grid_constant__ NeAtmosOceanDesc d, const __grid_constant__ Layout L,
                   const __grid_constant__ Thermo<CT> th, const __grid_constant__ FastParams P,
                   const __grid_constant__ TabParams T, const double* __restrict__ gtab) {
  __shared__ __align__(16) double tab[fm::TAB_SIZE];
  for (int k = threadIdx.x; k < fm::TAB_SIZE / 2; k += 256)
    reinterpret_cast<double2*>(tab)[k] = __ldg(reinterpret_cast<const double2*>(gtab) + k);
  __syncthreads();
  using FT = double;
  const int64_t n = (int64_t)L.ni * L.nj;
  const bool celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  const bool relative = d.properties.velocity_formulation == NE_VEL_RELATIVE;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < n; t += (int64_t)gridDim.x * 256) {
    const int32_t jj = (int32_t)(t / L.ni);
    const int64_t idx = L.at(L.i_lo + (int32_t)(t - (int64_t)jj * L.ni), L.j_lo + jj);
    const bool not_water = d.inactive ? (d.inactive[idx] != 0) : false;
    const bool skip = not_water && !P.fixed;   // needs_to_converge && not_water (:144)

    FT ustar = 0, theta_star = 0, q_star = 0;
    int iters = 0;
    if (!skip) {
      FastPoint s;
      {
        const FT au = __ldg((const FT*)d.ua + idx), av = __ldg((const FT*)d.va + idx);
        const FT aT = __ldg((const FT*)d.Ta + idx), ap = __ldg((const FT*)d.pa + idx), aq = __ldg((const FT*)d.qa + idx);
        const FT az = HS ? (FT)d.surface_layer_height.value : slot_at<FT>(d.surface_layer_height, idx);
        FT du = au, dv = av;
        if (relative) {
          du -= d.uo.ptr ? (slot_at<FT>(d.uo, idx) + slot_at<FT>(d.uo, idx + 1)) / 2 : (FT)d.uo.value;
          dv -= d.vo.ptr ? (slot_at<FT>(d.vo, idx) + slot_at<FT>(d.vo, idx + L.sx)) / 2 : (FT)d.vo.value;
        }
        FT To = slot_at<FT>(d.To, idx);
        if (celsius) To = To + 273.15;
        const FT qs = surface_specific_humidity<FT, CT>(d.properties, th, ap, To, slot_at<FT>(d.So, idx));
        const FT Tv = th.virtual_temperature(To, qs);
        s.gTv = P.g / Tv;
        s.c1 = 1 + th.delta * qs;
        s.c2 = th.delta * Tv;
        s.dudv2 = du * du + dv * dv;
        s.h_bl = HS ? (FT)d.boundary_layer_height.value : slot_at<FT>(d.boundary_layer_height, idx);
        s.hd = az - P.d_zero;
        s.log_hd = HS ? T.log_hd : log(s.hd);
        s.dtheta = (aT + P.g * az / th.cp_m(aq)) - To;
        s.dq = aq - qs;
        s.ustar = s.theta_star = s.q_star = 1e-4;
      }
      iters = tab_solve(P, T, tab, s);
      ustar = s.ustar; theta_star = s.theta_star; q_star = s.q_star;
    }
    // epilogue (atmosphere_ocean_fluxes.jl:160-196): atmosphere state re-read
    AtmosState<FT> a;
    a.u = __ldg((const FT*)d.ua + idx);
    a.v = __ldg((const FT*)d.va + idx);
    a.T = __ldg((const FT*)d.Ta + idx);
    a.p = __ldg((const FT*)d.pa + idx);
    a.q = __ldg((const FT*)d.qa + idx);
    FT du = a.u, dv = a.v, Ts;
    if (not_water) {  // zero_interface_state (interface_states.jl:800-803): Δu = uₐ − 0
      ustar = 0; theta_star = 0; q_star = 0; Ts = 273.15;
    } else {
      if (relative) {
        du -= d.uo.ptr ? (slot_at<FT>(d.uo, idx) + slot_at<FT>(d.uo, idx + 1)) / 2 : (FT)d.uo.value;
        dv -= d.vo.ptr ? (slot_at<FT>(d.vo, idx) + slot_at<FT>(d.vo, idx + L.sx)) / 2 : (FT)d.vo.value;
      }
      Ts = slot_at<FT>(d.To, idx);
      if (celsius) Ts = Ts + 273.15;
    }
    FluxEpilogue<FT, CT> e(th, a, ustar, theta_star, q_star, du, dv, false);
    ((FT*)d.latent_heat)[idx] = e.Qv;
    ((FT*)d.sensible_heat)[idx] = e.Qc;
    ((FT*)d.water_vapor)[idx] = e.Jv;
    ((FT*)d.x_momentum)[idx] = e.tx;
    ((FT*)d.y_momentum)[idx] = e.ty;
    ((FT*)d.interface_temperature)[idx] = celsius ? Ts - 273.15 : Ts;
    ((FT*)d.friction_velocity)[idx] = ustar;
    ((FT*)d.temperature_scale)[idx] = theta_star;
    ((FT*)d.water_vapor_scale)[idx] = q_star;
    if (d.iterations) d.iterations[idx] = iters;
  }
}

// Device-resident solver tables, built once per (device, ψ parameter set) and kept for the life of the
// process (14 KB each).  The first call for a parameter set allocates and copies synchronously, so it
// must happen outside CUDA-graph capture; later calls only enqueue the kernel.
struct SolverTables {
  int device;
  double key[26];
  double* dptr;
  TabParams T;
  double fit_error;
};
static std::mutex g_tab_mutex;
static std::vector<SolverTables> g_tabs;

static const SolverTables* solver_tables(const NeFluxFormulation& f) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  double key[26];
  for (int k = 0; k < 12; ++k) { key[k] = f.psi_momentum.a.p[k]; key[12 + k] = f.psi_temperature.a.p[k]; }
  key[24] = f.subgrid_velocities.gustiness_parameter;
  key[25] = f.subgrid_velocities.minimum_gustiness;
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  for (const SolverTables& t : g_tabs)
    if (t.device == dev && std::memcmp(t.key, key, sizeof(key)) == 0) return t.dptr ? &t : nullptr;
  SolverTables t;
  t.device = dev;
  std::memcpy(t.key, key, sizeof(key));
  t.dptr = nullptr;
  std::vector<double> host(fm::TAB_SIZE);
  t.fit_error = build_solver_tables(f, host.data(), t.T);
  if (t.fit_error <= 2e-15) {   // else: ψ parameters the polynomials cannot represent → closed-form kernel
    if (cudaMalloc(&t.dptr, sizeof(double) * fm::TAB_SIZE) != cudaSuccess ||
        cudaMemcpy(t.dptr, host.data(), sizeof(double) * fm::TAB_SIZE, cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaGetLastError();
      if (t.dptr) cudaFree(t.dptr);
      t.dptr = nullptr;
    }
  }
  g_tabs.push_back(t);
  return g_tabs.back().dptr ? &g_tabs.back() : nullptr;
}

// ---- host-side validation + dispatch -----------------------------------------------------------------
static bool stability_kind_ok(const NeStabilityFn& f) { return f.kind >= NE_PSI_ZERO && f.kind <= NE_PSI_LINEAR_STABLE; }
static bool profile_ok(const NeStabilityProfile& p) { return stability_kind_ok(p.a) && (!p.split || stability_kind_ok(p.b)); }
static bool rough_ok(const NeRoughnessLength& r) {
  if (r.kind < NE_ROUGH_CONSTANT || r.kind > NE_ROUGH_SCALAR) return false;
  if (r.wave_kind != NE_WAVE_CONSTANT && r.wave_kind != NE_WAVE_WIND_DEPENDENT) return false;
  if (r.visc_kind != NE_VISC_CONSTANT && r.visc_kind != NE_VISC_TEMPERATURE_DEPENDENT) return false;
  return true;
}

static int validate_formulation(const NeFluxFormulation& f, const NeInterfaceProperties& ip, bool ice) {
  if (f.kind == NE_FLUX_SIMILARITY_THEORY) {
    if (!profile_ok(f.psi_momentum) || !profile_ok(f.psi_temperature) || !profile_ok(f.psi_water_vapor))
      NE_NO_VARIANT("stability function with no kernel variant (user closures are not supported; there is no CPU fallback)");
    if (!rough_ok(f.ell_momentum) || !rough_ok(f.ell_temperature) || !rough_ok(f.ell_water_vapor))
      NE_NO_VARIANT("roughness length with no kernel variant");
    if (f.ell_momentum.kind == NE_ROUGH_SCALAR || f.ell_temperature.kind == NE_ROUGH_MOMENTUM ||
        f.ell_water_vapor.kind == NE_ROUGH_MOMENTUM)
      NE_NO_VARIANT("roughness length kind not valid in this slot");
    if (f.similarity_form != NE_PROFILE_LOGARITHMIC && f.similarity_form != NE_PROFILE_COARE)
      NE_NO_VARIANT("similarity profile with no kernel variant");
    const NeSubgridVelocity& s = f.subgrid_velocities;
    if (s.convective_kind < NE_SGS_NONE || s.convective_kind > NE_SGS_CONVECTIVE ||
        (s.composite && (s.mesoscale_kind < NE_SGS_NONE || s.mesoscale_kind > NE_SGS_CONVECTIVE)))
      NE_NO_VARIANT("subgrid velocity formulation with no kernel variant");
  } else if (f.kind == NE_FLUX_COEFFICIENT_BASED) {
    for (int k = 0; k < 3; ++k)
      if (f.coefficients[k].kind != NE_COEFF_CONSTANT && f.coefficients[k].kind != NE_COEFF_POLYNOMIAL_DRAG)
        NE_NO_VARIANT("transfer coefficient with no kernel variant (Function-valued coefficients are not supported)");
  } else if (f.kind == NE_FLUX_LARGE_YEAGER) {
    if (!profile_ok(f.large_yeager.psi_momentum) || !profile_ok(f.large_yeager.psi_temperature))
      NE_NO_VARIANT("Large-Yeager stability function with no kernel variant");
  } else {
    NE_NO_VARIANT("flux formulation kind %d has no kernel variant", f.kind);
  }
  if (f.stop.kind != NE_STOP_CONVERGENCE && f.stop.kind != NE_STOP_FIXED_ITERATIONS)
    NE_NO_VARIANT("solver stop criteria with no kernel variant");
  if (ip.phase != NE_PHASE_LIQUID && ip.phase != NE_PHASE_ICE) NE_NO_VARIANT("unknown thermodynamic phase");
  if (ip.x_h2o_kind < NE_XH2O_ONE || ip.x_h2o_kind > NE_XH2O_SALINITY) NE_NO_VARIANT("water mole fraction with no kernel variant");
  if (ip.velocity_formulation != NE_VEL_RELATIVE && ip.velocity_formulation != NE_VEL_WIND)
    NE_NO_VARIANT("velocity formulation with no kernel variant");
  const int tf = ip.temperature_formulation;
  if (tf < NE_TEMP_BULK || tf > NE_TEMP_SKIN_ICE_SNOW) NE_NO_VARIANT("temperature formulation with no kernel variant");
  if (!ice && (tf == NE_TEMP_SKIN_CONDUCTIVE || tf == NE_TEMP_SKIN_ICE_SNOW))
    NE_NO_VARIANT("conductive skin temperature needs a sea-ice interior state");
  if (ice && (tf == NE_TEMP_SKIN_DIFFUSIVE || tf == NE_TEMP_SKIN_DIFFUSIVE_INTERIOR))
    NE_NO_VARIANT("diffusive skin temperature is an ocean formulation");
  return NE_OK;
}

static bool viscosity_is_f64_literal(const NeFluxFormulation& f) {
  const NeRoughnessLength* r[3] = {&f.ell_momentum, &f.ell_temperature, &f.ell_water_vapor};
  for (auto p : r)
    if (p->kind != NE_ROUGH_CONSTANT && p->visc_kind == NE_VISC_CONSTANT && p->visc_dtype == NE_F64) return true;
  return false;
}

template <class FT>
static int ao_entry(const NeAtmosOceanDesc* d, void* stream) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  NE_REQUIRE(grid_ok(d->grid, 0, 1), "atmosphere-ocean: launch range + (i+1, j+1) stencil leaves the parent array");
  NE_REQUIRE(d->ua && d->va && d->Ta && d->pa && d->qa, "atmosphere-ocean: null atmosphere state array");
  NE_REQUIRE(d->latent_heat && d->sensible_heat && d->water_vapor && d->x_momentum && d->y_momentum &&
             d->interface_temperature && d->friction_velocity && d->temperature_scale && d->water_vapor_scale,
             "atmosphere-ocean: null output array");
  int rc = validate_formulation(d->flux, d->properties, false);
  if (rc != NE_OK) return rc;
  if (d->properties.temperature_formulation == NE_TEMP_SKIN_DIFFUSIVE_INTERIOR)
    NE_REQUIRE(d->kappa != nullptr, "InteriorDiffusivity needs the kappa array");
  if (d->radiation.enabled) NE_REQUIRE(d->radiation.downwelling_shortwave && d->radiation.downwelling_longwave, "radiation enabled without SW/LW arrays");
  cudaStream_t s = (cudaStream_t)stream;
  const bool ct64 = d->thermo.dtype == NE_F64;
  const bool v64 = std::is_same<FT, double>::value || viscosity_is_f64_literal(d->flux);
  if (std::is_same<FT, double>::value) {
    // NE_B200_FORCE_GENERIC=1 routes the default tree through the generic kernel (used by the parity tests)
    const char* force = std::getenv("NE_B200_FORCE_GENERIC");
    if (fast_path_eligible(d->flux, d->properties, d->thermo) && !(force && force[0] == '1')) {
      Layout L = make_layout(d->grid);
      FastParams P = make_fast_params(d->flux, d->gravitational_acceleration);
      const int64_t n = (int64_t)L.ni * L.nj;
      // NE_B200_CLOSED_FORM_PSI=1 keeps the libdevice closed-form iteration (parity-tested both ways)
      const char* closed = std::getenv("NE_B200_CLOSED_FORM_PSI");
      const SolverTables* tabs = (closed && closed[0] == '1') || !tab_path_eligible(d->flux) ? nullptr : solver_tables(d->flux);
      if (tabs) {
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const char* tm = std::getenv("NE_B200_TAB_MINB");
        const int tminb = tm ? std::atoi(tm) : 4;
        const char* tw = std::getenv("NE_B200_TAB_WAVES");
        const int waves = tw ? std::atoi(tw) : 8;
        const int64_t tiles = (n + 255) / 256;
        const unsigned tb = (unsigned)std::min<int64_t>(tiles, (int64_t)sms * tminb * waves);
        const bool hs = !d->surface_layer_height.ptr && !d->boundary_layer_height.ptr;
        TabParams TP = tabs->T;
        TP.log_hd = std::log(d->surface_layer_height.value - P.d_zero);
#define NE_LAUNCH_TAB2(MB, HS)                                                                                          \
  do {                                                                                                                  \
    if (ct64) ao_flux_tab_kernel<double, MB, HS><<<tb, 256, 0, s>>>(*d, L, Thermo<double>::make(d->thermo), P, TP, tabs->dptr); \
    else ao_flux_tab_kernel<float, MB, HS><<<tb, 256, 0, s>>>(*d, L, Thermo<float>::make(d->thermo), P, TP, tabs->dptr);        \
  } while (0)
#define NE_LAUNCH_TAB(MB)          \
  do {                             \
    if (hs) NE_LAUNCH_TAB2(MB, true); \
    else NE_LAUNCH_TAB2(MB, false);   \
  } while (0)
        if (tminb == 2) NE_LAUNCH_TAB(2);
        else if (tminb == 3) NE_LAUNCH_TAB(3);
        else NE_LAUNCH_TAB(4);
#undef NE_LAUNCH_TAB
#undef NE_LAUNCH_TAB2
        NE_CUDA_CHECK_LAUNCH("ne_atmosphere_ocean_fluxes(tab)");
        return NE_OK;
      }
      const char* mb = std::getenv("NE_B200_FAST_MINB");   // occupancy experiment knob
      const int minb = mb ? std::atoi(mb) : 8;   // 64 registers/thread measured fastest on B200 (profiles/r01_notes.md)
      const unsigned nb = (unsigned)((n + 127) / 128);
#define NE_LAUNCH_FAST(MB)                                                                                   \
  do {                                                                                                       \
    if (ct64) ao_flux_fast_kernel<double, MB><<<nb, 128, 0, s>>>(*d, L, Thermo<double>::make(d->thermo), P); \
    else ao_flux_fast_kernel<float, MB><<<nb, 128, 0, s>>>(*d, L, Thermo<float>::make(d->thermo), P);        \
  } while (0)
      if (minb == 4) NE_LAUNCH_FAST(4);
      else if (minb == 6) NE_LAUNCH_FAST(6);
      else NE_LAUNCH_FAST(8);
#undef NE_LAUNCH_FAST
      NE_CUDA_CHECK_LAUNCH("ne_atmosphere_ocean_fluxes(fast)");
      return NE_OK;
    }
    return ct64 ? launch_ao<double, double, double>(*d, s) : launch_ao<double, float, double>(*d, s);
  } else {
    if (ct64) return v64 ? launch_ao<float, double, double>(*d, s) : launch_ao<float, double, float>(*d, s);
    return v64 ? launch_ao<float, float, double>(*d, s) : launch_ao<float, float, float>(*d, s);
  }
}

template <class FT>
static int asi_entry(const NeAtmosSeaIceDesc* d, void* stream) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  NE_REQUIRE(grid_ok(d->grid, 0, 0), "atmosphere-sea-ice: launch range leaves the parent array");
  NE_REQUIRE(d->ua && d->va && d->Ta && d->pa && d->qa, "atmosphere-sea-ice: null atmosphere state array");
  NE_REQUIRE(d->latent_heat && d->sensible_heat && d->water_vapor && d->x_momentum && d->y_momentum && d->interface_temperature,
             "atmosphere-sea-ice: null output array");
  int rc = validate_formulation(d->flux, d->properties, true);
  if (rc != NE_OK) return rc;
  if (d->radiation.enabled) NE_REQUIRE(d->radiation.downwelling_shortwave && d->radiation.downwelling_longwave, "radiation enabled without SW/LW arrays");
  cudaStream_t s = (cudaStream_t)stream;
  const bool ct64 = d->thermo.dtype == NE_F64;
  const bool v64 = std::is_same<FT, double>::value || viscosity_is_f64_literal(d->flux);
  if (std::is_same<FT, double>::value) {
    return ct64 ? launch_asi<double, double, double>(*d, s) : launch_asi<double, float, double>(*d, s);
  } else {
    if (ct64) return v64 ? launch_asi<float, double, double>(*d, s) : launch_asi<float, double, float>(*d, s);
    return v64 ? launch_asi<float, float, double>(*d, s) : launch_asi<float, float, float>(*d, s);
  }
}

}  // namespace ne

extern "C" {
int ne_atmosphere_ocean_fluxes_f64(const NeAtmosOceanDesc* d, void* stream) { return ne::ao_entry<double>(d, stream); }
int ne_atmosphere_ocean_fluxes_f32(const NeAtmosOceanDesc* d, void* stream) { return ne::ao_entry<float>(d, stream); }
int ne_atmosphere_sea_ice_fluxes_f64(const NeAtmosSeaIceDesc* d, void* stream) { return ne::asi_entry<double>(d, stream); }
int ne_atmosphere_sea_ice_fluxes_f32(const NeAtmosSeaIceDesc* d, void* stream) { return ne::asi_entry<float>(d, stream); }
}
