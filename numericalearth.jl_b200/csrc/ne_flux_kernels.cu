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

#include <deque>
#include <mutex>
#include <vector>

#include "ne_flux_fast.cuh"
#include "ne_flux_tab.cuh"
#include "ne_flux_queue.cuh"
#include "ne_flux_asi_fast.cuh"
#include "ne_flux_land_fast.cuh"
#include "ne_queue_host.cuh"
#include "ne_interp_device.cuh"
#include "ne_physics.cuh"

namespace ne {

// generic kernels: defined in ne_flux_generic.cuh, instantiated in ne_flux_generic_*.cu
template <class FT, class CT, class VT> int launch_ao(const NeAtmosOceanDesc& d, cudaStream_t stream);
template <class FT, class CT, class VT> int launch_asi(const NeAtmosSeaIceDesc& d, cudaStream_t stream);
template <class FT, class CT, class VT> int launch_al(const NeAtmosLandDesc& d, cudaStream_t stream);
// round-2 a–o solve of the default tree (ne_flux_tab2.cu)
bool tab2_eligible(const NeAtmosOceanDesc& d, const TabParams& TP);
int launch_tab2(const NeAtmosOceanDesc& d, const Layout& L, const FastParams& P, const TabParams& TP, const double* tab,
                const double* host_tab, cudaStream_t s, unsigned long long* counts);

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
template <class CT, int MINB, bool HS, bool EXT>
__global__ void __launch_bounds__(256, MINB)
ao_flux_tab_kernel(const __grid_constant__ NeAtmosOceanDesc d, const __grid_constant__ Layout L,
                   const __grid_constant__ Thermo<CT> th, const __grid_constant__ FastParams P,
                   const __grid_constant__ TabParams T, const double* __restrict__ gtab) {
  extern __shared__ __align__(16) double tab[];   // fm::TAB_SIZE doubles (dynamic: see allow_table_smem)
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
      iters = tab_solve<EXT>(P, T, tab, s, (EXT && T.general_psi) ? &d.flux : nullptr);   // non-Edson tables: EXT build only
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

// ---- fused interpolation + atmosphere–ocean solve (update_state! phases 1 and 2 in one pass) ----------
// Per point: fractional indices → gathers of the 7 atmosphere series (+ 2 radiation series) at the two
// bracketing time levels from the L2-resident source → optional stores of the interpolated state (an
// output pointer left NULL is never materialised in HBM) → table-driven solve → 9 flux outputs.  The
// interpolated (T, p, q, Δu, Δv) the flux epilogue needs are parked in shared memory during the solve so
// they do not occupy registers across the iteration.  The solve never reads the interpolated state back
// from HBM: 5 field reads per point fewer than the unfused sequence, and one launch instead of three.
template <class CT, class AT, class TT, int MINB, bool HS>
__global__ void __launch_bounds__(256, MINB)
ao_fused_tab_kernel(const __grid_constant__ NeInterpDesc atm, const __grid_constant__ NeInterpDesc rad,
                    const __grid_constant__ NeAtmosOceanDesc d, const __grid_constant__ Layout L,
                    const __grid_constant__ InterpSource Sa, const __grid_constant__ InterpSource Sr,
                    const __grid_constant__ Thermo<CT> th, const __grid_constant__ FastParams P,
                    const __grid_constant__ TabParams T, const double* __restrict__ gtab) {
  extern __shared__ __align__(16) double tab[];   // fm::TAB_SIZE doubles (dynamic: see allow_table_smem)
  __shared__ double park[5][256];
  for (int k = threadIdx.x; k < fm::TAB_SIZE / 2; k += 256)
    reinterpret_cast<double2*>(tab)[k] = __ldg(reinterpret_cast<const double2*>(gtab) + k);
  __syncthreads();
  using FT = double;
  const int64_t n = (int64_t)L.ni * L.nj;
  const bool celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  const bool relative = d.properties.velocity_formulation == NE_VEL_RELATIVE;
  const TT nta = (TT)atm.time.frac, ntr = (TT)rad.time.frac;
  const bool same_a = atm.time.same != 0, same_r = rad.time.same != 0;
  const int tid = threadIdx.x;
  const bool has_rad = rad.n_fields > 0;
  const int64_t stride = (int64_t)gridDim.x * 256;
  auto index_of = [&](int64_t t) {
    const int32_t jj = (int32_t)(t / L.ni);
    return L.at(L.i_lo + (int32_t)(t - (int64_t)jj * L.ni), L.j_lo + jj);
  };
  int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  // the fractional indices of the NEXT tile are requested before the current solve starts, so their
  // DRAM latency is hidden behind ~10^4 cycles of arithmetic
  FracPair<AT> fa = {0, 0}, fr = {0, 0};
  if (t < n) {
    const int64_t idx0 = index_of(t);
    fa = load_frac<AT>(atm.frac_i, atm.frac_j, idx0);
    if (has_rad) fr = load_frac<AT>(rad.frac_i, rad.frac_j, idx0);
  }
  for (; t < n; t += stride) {
    const int64_t idx = index_of(t);
    const bool not_water = d.inactive ? (d.inactive[idx] != 0) : false;
    const bool skip = not_water && !P.fixed;   // needs_to_converge && not_water (:144)

    FT ustar = 0, theta_star = 0, q_star = 0;
    int iters = 0;
    {
      // ---- phase 1: interpolation (interpolate_atmospheric_state.jl:91-137, interpolate_radiation_state.jl:43-69)
      // all gathers of the point are issued before any of them is consumed
      const InterpPoint<AT> pa = interp_point<AT>(atm.frac_i != nullptr, atm.frac_j != nullptr, fa, Sa);
      const InterpPoint<AT> pr = interp_point<AT>(rad.frac_i != nullptr, rad.frac_j != nullptr, fr, Sr);
      if (t + stride < n) {
        const int64_t idxn = index_of(t + stride);
        fa = load_frac<AT>(atm.frac_i, atm.frac_j, idxn);
        if (has_rad) fr = load_frac<AT>(rad.frac_i, rad.frac_j, idxn);
      }
      Corners8<AT> c[5];
#pragma unroll
      for (int f = 0; f < 5; ++f) c[f] = gather8<AT>((const AT*)atm.series[f][0].data, pa, Sa, same_a);   // u v T q p
      Corners8<AT> cr[2], cp[2];
      bool simple_r[2], simple_p[2];
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        simple_r[f] = has_rad && f < rad.n_fields && rad.out[f] && rad.n_summands[f] == 1 && rad.series[f][0].data;
        if (simple_r[f]) cr[f] = gather8<AT>((const AT*)rad.series[f][0].data, pr, Sr, same_r);
        simple_p[f] = 5 + f < atm.n_fields && atm.out[5 + f] && atm.n_summands[5 + f] == 1 && atm.series[5 + f][0].data;
        if (simple_p[f]) cp[f] = gather8<AT>((const AT*)atm.series[5 + f][0].data, pa, Sa, same_a);
      }
      FT st[5];
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        st[f] = (FT)blend8<AT, TT>(c[f], pa, nta, same_a);
        if (atm.out[f]) ((FT*)atm.out[f])[idx] = st[f];
      }
      if (atm.potential) {
        const int pf = atm.potential_from;
        const FT v = pf == 0 ? st[0] : pf == 1 ? st[1] : pf == 2 ? st[2] : pf == 3 ? st[3] : st[4];
        ((FT*)atm.potential)[idx] = div_rn(v, (FT)atm.ocean_reference_density);
      }
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        if (simple_r[f]) ((FT*)rad.out[f])[idx] = (FT)blend8<AT, TT>(cr[f], pr, ntr, same_r);
        else if (has_rad && f < rad.n_fields && rad.out[f]) ((FT*)rad.out[f])[idx] = (FT)interp_field<AT, TT>(rad, f, pr, Sr, ntr, same_r);
        if (simple_p[f]) ((FT*)atm.out[5 + f])[idx] = (FT)blend8<AT, TT>(cp[f], pa, nta, same_a);   // rain, snow: only written
        else if (5 + f < atm.n_fields && atm.out[5 + f]) ((FT*)atm.out[5 + f])[idx] = (FT)interp_field<AT, TT>(atm, 5 + f, pa, Sa, nta, same_a);
      }
      const FT au = st[0], av = st[1], aT = st[2], aq = st[3], ap = st[4];
      // ---- phase 2: the solve (atmosphere_ocean_fluxes.jl:80-197)
      FT du = au, dv = av;
      if (relative && !not_water) {
        du -= d.uo.ptr ? (slot_at<FT>(d.uo, idx) + slot_at<FT>(d.uo, idx + 1)) / 2 : (FT)d.uo.value;
        dv -= d.vo.ptr ? (slot_at<FT>(d.vo, idx) + slot_at<FT>(d.vo, idx + L.sx)) / 2 : (FT)d.vo.value;
      }
      park[0][tid] = du; park[1][tid] = dv; park[2][tid] = aT; park[3][tid] = ap; park[4][tid] = aq;
      if (!skip) {
        FastPoint s;
        if (relative && not_water) {   // FixedIterations solves inactive cells too, with the ocean velocity
          du -= d.uo.ptr ? (slot_at<FT>(d.uo, idx) + slot_at<FT>(d.uo, idx + 1)) / 2 : (FT)d.uo.value;
          dv -= d.vo.ptr ? (slot_at<FT>(d.vo, idx) + slot_at<FT>(d.vo, idx + L.sx)) / 2 : (FT)d.vo.value;
        }
        const FT az = HS ? (FT)d.surface_layer_height.value : slot_at<FT>(d.surface_layer_height, idx);
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
        iters = tab_solve<false>(P, T, tab, s);
        ustar = s.ustar; theta_star = s.theta_star; q_star = s.q_star;
      }
    }
    // ---- epilogue (atmosphere_ocean_fluxes.jl:160-196)
    AtmosState<FT> a;
    a.u = 0; a.v = 0; a.z = 0; a.h_bl = 0;
    const FT du = park[0][tid], dv = park[1][tid];
    a.T = park[2][tid]; a.p = park[3][tid]; a.q = park[4][tid];
    FT Ts;
    if (not_water) {  // zero_interface_state (interface_states.jl:800-803): Δu = uₐ − 0 (parked as such)
      ustar = 0; theta_star = 0; q_star = 0; Ts = 273.15;
    } else {
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
struct SolverTableKey {   // everything build_solver_tables reads
  NeStabilityProfile psi_momentum, psi_temperature, psi_water_vapor;
  double gustiness_parameter, minimum_gustiness;
  int64_t f32;
};
struct SolverTables {
  int device;
  SolverTableKey key;
  std::vector<double> host;   // host copy (the |ζ| < 2^-12 records also travel as kernel parameters)
  double* dptr;
  TabParams T;
  double fit_error;
};
static std::mutex g_tab_mutex;
static std::deque<SolverTables> g_tabs;   // a deque never moves its elements: the returned pointers stay valid

static const SolverTables* solver_tables(const NeFluxFormulation& f, bool f32 = false) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  SolverTableKey key;
  std::memset(&key, 0, sizeof(key));
  std::memcpy(&key.psi_momentum, &f.psi_momentum, sizeof(NeStabilityProfile));
  std::memcpy(&key.psi_temperature, &f.psi_temperature, sizeof(NeStabilityProfile));
  std::memcpy(&key.psi_water_vapor, &f.psi_water_vapor, sizeof(NeStabilityProfile));
  key.gustiness_parameter = f.subgrid_velocities.gustiness_parameter;
  key.minimum_gustiness = f.subgrid_velocities.minimum_gustiness;
  key.f32 = f32 ? 1 : 0;
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  for (const SolverTables& t : g_tabs)
    if (t.device == dev && std::memcmp(&t.key, &key, sizeof(key)) == 0) return t.dptr ? &t : nullptr;
  SolverTables t;
  t.device = dev;
  std::memcpy(&t.key, &key, sizeof(key));
  t.dptr = nullptr;
  std::vector<double>& host = t.host;
  host.assign(fm::TAB_SIZE, 0.0);
  t.fit_error = build_solver_tables(f, host.data(), t.T, f32);
  if (t.fit_error <= 2e-15) {   // else: ψ parameters the polynomials cannot represent → closed-form / generic kernel
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
  if (r.kind < NE_ROUGH_CONSTANT || r.kind > NE_ROUGH_LAND) return false;
  if (r.wave_kind != NE_WAVE_CONSTANT && r.wave_kind != NE_WAVE_WIND_DEPENDENT) return false;
  if (r.visc_kind != NE_VISC_CONSTANT && r.visc_kind != NE_VISC_TEMPERATURE_DEPENDENT) return false;
  return true;
}

static int validate_formulation(const NeFluxFormulation& f, const NeInterfaceProperties& ip, bool ice, bool land = false) {
  if (f.kind == NE_FLUX_SIMILARITY_THEORY) {
    // LandRoughnessLength / LandZeroPlaneDisplacement read the land model's per-cell properties; over the ocean and sea ice
    // `interior_properties` has no such fields and the markers collapse to constants, which the binding passes as such
    if (!land && (f.ell_momentum.kind == NE_ROUGH_LAND || f.ell_temperature.kind == NE_ROUGH_LAND ||
                  f.ell_water_vapor.kind == NE_ROUGH_LAND || f.zero_plane_displacement_kind != NE_DISPLACEMENT_CONSTANT))
      NE_NO_VARIANT("per-cell land roughness / displacement markers belong to the atmosphere-land interface (pass the resolved constants here)");
    if (f.zero_plane_displacement_kind != NE_DISPLACEMENT_CONSTANT && f.zero_plane_displacement_kind != NE_DISPLACEMENT_LAND)
      NE_NO_VARIANT("zero-plane displacement with no kernel variant");
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

static int validate_ao(const NeAtmosOceanDesc* d, bool need_atmosphere_arrays) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  NE_REQUIRE(grid_ok(d->grid, 0, 1), "atmosphere-ocean: launch range + (i+1, j+1) stencil leaves the parent array");
  if (need_atmosphere_arrays)
    NE_REQUIRE(d->ua && d->va && d->Ta && d->pa && d->qa, "atmosphere-ocean: null atmosphere state array");
  NE_REQUIRE(d->latent_heat && d->sensible_heat && d->water_vapor && d->x_momentum && d->y_momentum &&
             d->interface_temperature && d->friction_velocity && d->temperature_scale && d->water_vapor_scale,
             "atmosphere-ocean: null output array");
  int rc = validate_formulation(d->flux, d->properties, false);
  if (rc != NE_OK) return rc;
  if (d->properties.temperature_formulation == NE_TEMP_SKIN_DIFFUSIVE_INTERIOR)
    NE_REQUIRE(d->kappa != nullptr, "InteriorDiffusivity needs the kappa array");
  if (d->radiation.enabled) NE_REQUIRE(d->radiation.downwelling_shortwave && d->radiation.downwelling_longwave, "radiation enabled without SW/LW arrays");
  return NE_OK;
}

// grid size of the table-driven kernels: enough 256-thread CTAs for `waves` rounds of full residency
static unsigned tab_grid(int64_t n, int minb) {
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const char* tw = std::getenv("NE_B200_TAB_WAVES");
  const int waves = tw ? std::atoi(tw) : 8;
  const int64_t tiles = (n + 255) / 256;
  return (unsigned)std::min<int64_t>(tiles, (int64_t)sms * minb * waves);
}

template <class FT>
static int ao_entry(const NeAtmosOceanDesc* d, void* stream) {
  int rc = validate_ao(d, true);
  if (rc != NE_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const bool ct64 = d->thermo.dtype == NE_F64;
  const bool v64 = std::is_same<FT, double>::value || viscosity_is_f64_literal(d->flux);
  if (std::is_same<FT, double>::value) {
    // NE_B200_FORCE_GENERIC=1 routes the default tree through the generic kernel (used by the parity tests)
    // the default tree with the Edson functions, or with any other pair of the shipped stability functions whose
    // tables verify (Large–Yeager: Split(LinearStable, Paulson); SHEBA; …) — the latter only on the table kernel
    const bool edson = fast_path_eligible(d->flux, d->properties, d->thermo);
    const bool other_psi = !edson && default_roughness_gustiness(d->flux) && tab_path_eligible(d->flux) &&
                           d->properties.temperature_formulation == NE_TEMP_BULK && !env_flag("NE_B200_CLOSED_FORM_PSI");
    const SolverTables* other_tabs = other_psi && !env_flag("NE_B200_FORCE_GENERIC") ? solver_tables(d->flux) : nullptr;
    if ((edson || other_tabs) && !env_flag("NE_B200_FORCE_GENERIC")) {
      Layout L = make_layout(d->grid);
      FastParams P = make_fast_params(d->flux, d->gravitational_acceleration);
      const int64_t n = (int64_t)L.ni * L.nj;
      // NE_B200_CLOSED_FORM_PSI=1 keeps the libdevice closed-form iteration (parity-tested both ways)
      const SolverTables* tabs = other_tabs ? other_tabs
          : (env_flag("NE_B200_CLOSED_FORM_PSI") || !tab_path_eligible(d->flux) ? nullptr : solver_tables(d->flux));
      // Float64: the work-queue kernel (opt-in, NE_B200_QUEUE=1) does not beat the one-thread-per-point kernel on
      // B200 (1.78 vs 1.73 ms on C4: the trip counts only spread 7–24 and the desynchronised warps cost more in
      // instruction-cache misses and exposed load latency than the denser rounds save, profiles/r01_notes.md)
      const bool ext = !strict_default_options(d->flux) || (tabs && tabs->T.general_psi);
      if (tabs && env_flag("NE_B200_QUEUE") && queue_path_ok(d->grid) && !ext)
        return ct64 ? launch_queue<double, double>(*d, tabs->T, tabs->dptr, s) : launch_queue<double, float>(*d, tabs->T, tabs->dptr, s);
      if (tabs) {
        const unsigned tb = tab_grid(n, 3);
        const bool hs = !d->surface_layer_height.ptr && !d->boundary_layer_height.ptr;
        TabParams TP = tabs->T;
        TP.far_fm = !TP.general_psi && far_unstable_fm_ok(P);
        TP.log_hd = std::log(d->surface_layer_height.value - P.d_zero);
        // the strict default tree with scalar heights: the round-2 kernel (NE_B200_TAB_V1=1 keeps the round-1 one)
        if (!ext && tab2_eligible(*d, TP)) return launch_tab2(*d, L, P, TP, tabs->dptr, tabs->host.data(), s, nullptr);
#define NE_LAUNCH_TAB3(MB, HS, EXT)                                                                                          \
  do {                                                                                                                  \
    if (ct64) { allow_table_smem<ao_flux_tab_kernel<double, MB, HS, EXT>>(); ao_flux_tab_kernel<double, MB, HS, EXT><<<tb, 256, TAB_SMEM_BYTES, s>>>(*d, L, Thermo<double>::make(d->thermo), P, TP, tabs->dptr); } \
    else { allow_table_smem<ao_flux_tab_kernel<float, MB, HS, EXT>>(); ao_flux_tab_kernel<float, MB, HS, EXT><<<tb, 256, TAB_SMEM_BYTES, s>>>(*d, L, Thermo<float>::make(d->thermo), P, TP, tabs->dptr); }        \
  } while (0)
#define NE_LAUNCH_TAB2(MB, HS)          \
  do {                                  \
    if (ext) NE_LAUNCH_TAB3(MB, HS, true); \
    else NE_LAUNCH_TAB3(MB, HS, false);    \
  } while (0)
#define NE_LAUNCH_TAB(MB)          \
  do {                             \
    if (hs) NE_LAUNCH_TAB2(MB, true); \
    else NE_LAUNCH_TAB2(MB, false);   \
  } while (0)
        NE_LAUNCH_TAB(3);   // 80 registers, 3 CTAs per SM: the measured optimum (profiles/r01_notes.md: 64 regs 1.89-1.99 ms, 128 regs 1.87 ms)
#undef NE_LAUNCH_TAB
#undef NE_LAUNCH_TAB2
#undef NE_LAUNCH_TAB3
        NE_CUDA_CHECK_LAUNCH("ne_atmosphere_ocean_fluxes(tab)");
        return NE_OK;
      }
      add_small_zeta_poly(P, d->flux);   // closed-form kernel only
      const unsigned nb = (unsigned)((n + 127) / 128);   // 64 registers/thread measured fastest on B200 (profiles/r01_notes.md)
#define NE_LAUNCH_FAST(MB)                                                                                   \
  do {                                                                                                       \
    if (ct64) ao_flux_fast_kernel<double, MB><<<nb, 128, 0, s>>>(*d, L, Thermo<double>::make(d->thermo), P); \
    else ao_flux_fast_kernel<float, MB><<<nb, 128, 0, s>>>(*d, L, Thermo<float>::make(d->thermo), P);        \
  } while (0)
      NE_LAUNCH_FAST(8);
#undef NE_LAUNCH_FAST
      NE_CUDA_CHECK_LAUNCH("ne_atmosphere_ocean_fluxes(fast)");
      return NE_OK;
    }
    return ct64 ? launch_ao<double, double, double>(*d, s) : launch_ao<double, float, double>(*d, s);
  } else {
    // Float32 model, default tree, Float32 thermodynamics, Float64-literal viscosity: the mixed-precision table
    // iteration on the work-queue kernel; everything else (and NE_B200_FORCE_GENERIC=1) through the generic kernel
    if (!ct64 && viscosity_is_f64_literal(d->flux) && fast_path_eligible(d->flux, d->properties, d->thermo) &&
        tab_path_eligible(d->flux) && queue_path_ok(d->grid) && !env_flag("NE_B200_FORCE_GENERIC") &&
        !env_flag("NE_B200_CLOSED_FORM_PSI")) {
      const SolverTables* tabs = solver_tables(d->flux, true);
      if (tabs) return launch_queue<float, float>(*d, tabs->T, tabs->dptr, s);
    }
    if (ct64) return v64 ? launch_ao<float, double, double>(*d, s) : launch_ao<float, double, float>(*d, s);
    return v64 ? launch_ao<float, float, double>(*d, s) : launch_ao<float, float, float>(*d, s);
  }
}

static bool same_launch(const NeExchangeGrid& a, const NeExchangeGrid& b) {
  return a.nx == b.nx && a.ny == b.ny && a.hx == b.hx && a.hy == b.hy && a.i_lo == b.i_lo && a.i_hi == b.i_hi &&
         a.j_lo == b.j_lo && a.j_hi == b.j_hi;
}

// Fused interpolation + a–o solve for the default plugin tree in Float64.  Returns NE_OK when the fused kernel
// was enqueued, +1 when this step does not qualify (the caller then enqueues the component kernels), < 0 on error.
int fused_interp_ao_f64(const NeInterpDesc* atm, const NeInterpDesc* rad, const NeAtmosOceanDesc* d, void* stream) {
  // Opt-in (NE_B200_FUSE_INTERP=1): measured on B200 the single-pass kernel is slower than the component
  // kernels (2.85 vs 2.16 ms on C4): warps parked in the gather phase lower the occupancy the FP64-latency-bound
  // iteration needs, and the larger code thrashes the instruction cache (profiles/r01_notes.md).
  if (!env_flag("NE_B200_FUSE_INTERP") || env_flag("NE_B200_FORCE_GENERIC") || env_flag("NE_B200_CLOSED_FORM_PSI")) return 1;
  if (!atm || !d || !fast_path_eligible(d->flux, d->properties, d->thermo) || !tab_path_eligible(d->flux)) return 1;
  const bool has_rad = rad && rad->n_fields > 0;
  if (atm->n_fields < 5 || atm->n_fields > 7 || !same_launch(atm->grid, d->grid)) return 1;
  if (atm->rotation_cos || atm->rotation_sin) return 1;   // rotated exchange grids: component kernels
  for (int f = 0; f < 5; ++f)
    if (atm->n_summands[f] != 1 || !atm->series[f][0].data) return 1;
  if (atm->potential && (atm->potential_from < 0 || atm->potential_from > 4)) return 1;
  if (has_rad && (rad->n_fields > 2 || !same_launch(rad->grid, d->grid) || rad->src_dtype != atm->src_dtype ||
                  rad->time.frac_dtype != atm->time.frac_dtype)) return 1;
  int rc = validate_ao(d, false);
  if (rc != NE_OK) return rc;
  for (const NeInterpDesc* x : {atm, has_rad ? rad : atm}) {
    NE_REQUIRE(x->src_nx > 0 && x->src_ny > 0 && x->src_nt > 0, "interp: bad source extents");
    NE_REQUIRE((x->src_nx + 2 * x->src_hx) * (x->src_ny + 2 * x->src_hy) < (int64_t)1 << 31, "interp: source plane exceeds 2^31 elements");
    NE_REQUIRE(x->time.m1 >= 1 && x->time.m1 <= x->src_nt && x->time.m2 >= 1 && x->time.m2 <= x->src_nt,
               "interp: time memory slots out of range");
    for (int f = 0; f < x->n_fields; ++f)
      NE_REQUIRE(x->n_summands[f] >= 0 && x->n_summands[f] <= NE_MAX_SUMMANDS, "interp: too many summands");
  }
  const SolverTables* tabs = solver_tables(d->flux);
  if (!tabs) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  Layout L = make_layout(d->grid);
  FastParams P = make_fast_params(d->flux, d->gravitational_acceleration);
  const int64_t n = (int64_t)L.ni * L.nj;
  const unsigned tb = tab_grid(n, 3);
  const bool hs = !d->surface_layer_height.ptr && !d->boundary_layer_height.ptr;
  TabParams TP = tabs->T;
  TP.far_fm = !TP.general_psi && far_unstable_fm_ok(P);
  TP.log_hd = std::log(d->surface_layer_height.value - P.d_zero);
  NeInterpDesc norad;
  std::memset(&norad, 0, sizeof(norad));
  const NeInterpDesc& r = has_rad ? *rad : norad;
  const InterpSource Sa = make_interp_source(*atm);
  InterpSource Sr = Sa;
  if (has_rad) Sr = make_interp_source(*rad);
  const bool ct64 = d->thermo.dtype == NE_F64, a64 = atm->src_dtype == NE_F64, t64 = atm->time.frac_dtype == NE_F64;
#define NE_FUSED4(CT, AT, TT, MB, HS)                                  \
  do {                                                                 \
    allow_table_smem<ao_fused_tab_kernel<CT, AT, TT, MB, HS>>();        \
    ao_fused_tab_kernel<CT, AT, TT, MB, HS><<<tb, 256, TAB_SMEM_BYTES, s>>>(*atm, r, *d, L, Sa, Sr, Thermo<CT>::make(d->thermo), P, TP, tabs->dptr); \
  } while (0)
#define NE_FUSED3(CT, AT, TT)                     \
  do {                                            \
    if (hs) NE_FUSED4(CT, AT, TT, 3, true);       \
    else NE_FUSED4(CT, AT, TT, 3, false);         \
  } while (0)
  // thermodynamics default to the atmosphere's element type (prescribed_atmosphere.jl:224): the mixed
  // (CT ≠ AT) combinations and a time fraction narrower than the data are left to the component kernels
  if (ct64 && a64 && t64) NE_FUSED3(double, double, double);
  else if (!ct64 && !a64 && t64) NE_FUSED3(float, float, double);
  else if (!ct64 && !a64 && !t64) NE_FUSED3(float, float, float);
  else return 1;
#undef NE_FUSED3
#undef NE_FUSED4
  NE_CUDA_CHECK_LAUNCH("ne_fused_interface_step(interp+solve)");
  return NE_OK;
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
    // default tree: work-queue kernel with the table-driven similarity step (NE_B200_FORCE_GENERIC=1: generic kernel)
    if (asi_fast_path_eligible(d->flux, d->properties) && queue_path_ok(d->grid) && !env_flag("NE_B200_FORCE_GENERIC") &&
        !env_flag("NE_B200_CLOSED_FORM_PSI")) {
      const SolverTables* tabs = solver_tables(d->flux, false);
      if (tabs) return ct64 ? launch_asi_queue<double, double>(*d, tabs->T, tabs->dptr, s) : launch_asi_queue<double, float>(*d, tabs->T, tabs->dptr, s);
    }
    return ct64 ? launch_asi<double, double, double>(*d, s) : launch_asi<double, float, double>(*d, s);
  } else {
    // Float32 model: the same kernel with the mixed-precision similarity step (Float32 thermodynamics, Float64-literal
    // viscosity, strict default options)
    if (!ct64 && viscosity_is_f64_literal(d->flux) && asi_fast_path_eligible(d->flux, d->properties) &&
        strict_default_options(d->flux) && queue_path_ok(d->grid) && !env_flag("NE_B200_FORCE_GENERIC") &&
        !env_flag("NE_B200_CLOSED_FORM_PSI")) {
      const SolverTables* tabs = solver_tables(d->flux, true);
      if (tabs) return launch_asi_queue<float, float>(*d, tabs->T, tabs->dptr, s);
    }
    if (ct64) return v64 ? launch_asi<float, double, double>(*d, s) : launch_asi<float, double, float>(*d, s);
    return v64 ? launch_asi<float, float, double>(*d, s) : launch_asi<float, float, float>(*d, s);
  }
}

template <class FT>
static int al_entry(const NeAtmosLandDesc* d, void* stream) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  NE_REQUIRE(grid_ok(d->grid, 0, 0), "atmosphere-land: launch range leaves the parent array");
  NE_REQUIRE(d->ua && d->va && d->Ta && d->pa && d->qa, "atmosphere-land: null atmosphere state array");
  NE_REQUIRE(d->latent_heat && d->sensible_heat && d->water_vapor && d->x_momentum && d->y_momentum &&
             d->interface_temperature && d->friction_velocity && d->temperature_scale && d->water_vapor_scale,
             "atmosphere-land: null output array");
  int rc = validate_formulation(d->flux, d->properties, false, true);
  if (rc != NE_OK) return rc;
  // SkinTemperature's flux balance reads reference_density / heat_capacity of the interior properties (interface_states.jl:
  // 434-457), which atmosphere_land_surface_properties does not provide (atmosphere_land_fluxes.jl:114-115): the reference has
  // no working method for it either
  if (d->properties.temperature_formulation != NE_TEMP_BULK)
    NE_NO_VARIANT("atmosphere-land: only BulkTemperature has a kernel variant (the reference's land properties carry no heat capacity / density for a skin-temperature balance)");
  const NeLandHumidity& h = d->humidity;
  if (h.kind < NE_LANDQ_BULK || h.kind > NE_LANDQ_DRY_LAYER)
    NE_NO_VARIANT("land humidity formulation %d has no kernel variant", h.kind);
  if (h.kind == NE_LANDQ_DRY_LAYER) {
    if (h.tortuosity != NE_TORTUOSITY_CONSTANT && h.tortuosity != NE_TORTUOSITY_POWER_LAW) NE_NO_VARIANT("tortuosity model with no kernel variant");
    NE_REQUIRE(h.dry_layer_onset_saturation > 0 && h.thermal_exchange_depth > 0 && h.minimum_dry_layer_depth > 0 && h.porosity > 0,
               "DryLayerHumidity: onset saturation, thermal exchange depth, minimum dry-layer depth and porosity must be positive");
  }
  if (h.phase != NE_PHASE_LIQUID && h.phase != NE_PHASE_ICE) NE_NO_VARIANT("unknown thermodynamic phase");
  if (h.kind == NE_LANDQ_FRACTIONAL_CRITICAL) NE_REQUIRE(h.critical_saturation > 0, "CriticalSaturation must be positive");
  if (h.kind == NE_LANDQ_SKIN) NE_REQUIRE(h.surface_thickness > 0, "SkinHumidity: surface_thickness must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  const bool ct64 = d->thermo.dtype == NE_F64;
  const bool v64 = std::is_same<FT, double>::value || viscosity_is_f64_literal(d->flux);
  if (std::is_same<FT, double>::value) {
    // the land defaults (constant roughness, tabulatable stability functions, BulkTemperature): work-queue kernel
    if (land_fast_path_eligible(d->flux, d->properties) && queue_path_ok(d->grid) && !env_flag("NE_B200_FORCE_GENERIC") &&
        !env_flag("NE_B200_CLOSED_FORM_PSI")) {
      const SolverTables* tabs = solver_tables(d->flux, false);
      if (tabs) return ct64 ? launch_land_queue<double>(*d, tabs->T, tabs->dptr, s) : launch_land_queue<float>(*d, tabs->T, tabs->dptr, s);
    }
    return ct64 ? launch_al<double, double, double>(*d, s) : launch_al<double, float, double>(*d, s);
  }
  if (ct64) return v64 ? launch_al<float, double, double>(*d, s) : launch_al<float, double, float>(*d, s);
  return v64 ? launch_al<float, float, double>(*d, s) : launch_al<float, float, float>(*d, s);
}

// FP64 instructions one launch of the shipped a–o solve executes on this descriptor (the OpsCount instantiation of
// the same kernel source: thread-level fma / mul / add counts, library-code estimate, thread trips, warp trips).
static int count_solve_ops(const NeAtmosOceanDesc* d, uint64_t* out, void* stream) {
  int rc = validate_ao(d, true);
  if (rc != NE_OK) return rc;
  NE_REQUIRE(out != nullptr, "null output");
  if (!fast_path_eligible(d->flux, d->properties, d->thermo) || !tab_path_eligible(d->flux) || !strict_default_options(d->flux))
    NE_NO_VARIANT("operation counting covers the default plugin tree only");
  const SolverTables* tabs = solver_tables(d->flux);
  NE_REQUIRE(tabs != nullptr, "solver tables unavailable");
  Layout L = make_layout(d->grid);
  FastParams P = make_fast_params(d->flux, d->gravitational_acceleration);
  TabParams TP = tabs->T;
  TP.far_fm = !TP.general_psi && far_unstable_fm_ok(P);
  TP.log_hd = std::log(d->surface_layer_height.value - P.d_zero);
  if (!tab2_eligible(*d, TP)) NE_NO_VARIANT("operation counting needs the round-2 solve kernel (scalar heights, < 2^31 points)");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, 8 * sizeof(unsigned long long));
  if (e != cudaSuccess) return cuda_error(e, "ne_count_solve_ops: cudaMalloc");
  cudaMemsetAsync(dev, 0, 8 * sizeof(unsigned long long), s);
  rc = launch_tab2(*d, L, P, TP, tabs->dptr, tabs->host.data(), s, dev);
  if (rc == NE_OK) {
    unsigned long long host[8];
    e = cudaMemcpyAsync(host, dev, sizeof(host), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = cuda_error(e, "ne_count_solve_ops");
    else for (int k = 0; k < 8; ++k) out[k] = host[k];
  }
  cudaFree(dev);
  return rc;
}

}  // namespace ne

extern "C" {
int ne_count_solve_ops_f64(const NeAtmosOceanDesc* d, uint64_t* out, void* stream) { NE_NVTX(); return ne::count_solve_ops(d, out, stream); }
int ne_atmosphere_land_fluxes_f64(const NeAtmosLandDesc* d, void* stream) { NE_NVTX(); return ne::al_entry<double>(d, stream); }
int ne_atmosphere_land_fluxes_f32(const NeAtmosLandDesc* d, void* stream) { NE_NVTX(); return ne::al_entry<float>(d, stream); }
int ne_atmosphere_ocean_fluxes_f64(const NeAtmosOceanDesc* d, void* stream) { NE_NVTX(); return ne::ao_entry<double>(d, stream); }
int ne_atmosphere_ocean_fluxes_f32(const NeAtmosOceanDesc* d, void* stream) { NE_NVTX(); return ne::ao_entry<float>(d, stream); }
int ne_atmosphere_sea_ice_fluxes_f64(const NeAtmosSeaIceDesc* d, void* stream) { NE_NVTX(); return ne::asi_entry<double>(d, stream); }
int ne_atmosphere_sea_ice_fluxes_f32(const NeAtmosSeaIceDesc* d, void* stream) { NE_NVTX(); return ne::asi_entry<float>(d, stream); }
}
