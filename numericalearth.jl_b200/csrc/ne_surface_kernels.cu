// ne_surface_kernels.cu — the HBM-bound streaming kernels either side of the turbulent solve:
// sea-ice–ocean fluxes (frazil column clamp + interface heat/salt), sea-ice–ocean stress,
// net ocean / sea-ice flux assembly, radiative flux application and the diagnostics reduction.
//
// Replaces (citations relative to /root/reference/src/):
//   _compute_sea_ice_ocean_fluxes!        EarthSystemModels/InterfaceComputations/sea_ice_ocean_fluxes.jl:106-226
//   _compute_sea_ice_ocean_stress!        EarthSystemModels/InterfaceComputations/sea_ice_ocean_fluxes.jl:79-104
//   compute_interface_heat_flux, solve_interface_conditions
//                                         EarthSystemModels/InterfaceComputations/sea_ice_ocean_heat_flux_formulations.jl:176-313
//   get_friction_velocity                 EarthSystemModels/InterfaceComputations/friction_velocity.jl:24-44
//   _freeze_ocean_temperature!            SeaIces/freezing_limited_ocean_temperature.jl:95-118
//   _assemble_net_ocean_fluxes!           Oceans/assemble_net_ocean_fluxes.jl:74-153
//   _assemble_net_sea_ice_fluxes!         SeaIces/assemble_net_sea_ice_fluxes.jl:42-81
//   _apply_air_sea_radiative_fluxes!      Radiations/apply_air_sea_radiative_fluxes.jl:62-111
//   _apply_air_sea_ice_radiative_fluxes!  Radiations/apply_air_sea_ice_radiative_fluxes.jl:55-90
//
// All are one-thread-per-point, consecutive threads along x: every load/store of a warp is one
// fully-used 128 B (f32) / 256 B (f64) segment; constants arrive through NeSlot without a load.
#include <cstdlib>

#include "ne_physics.cuh"

namespace ne {

#define NE_POINT_INDEX()                                                        \
  const int64_t t__ = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;           \
  if (t__ >= (int64_t)L.ni * L.nj) return;                                      \
  const int32_t jj__ = (int32_t)(t__ / L.ni);                                   \
  const int32_t i = L.i_lo + (int32_t)(t__ - (int64_t)jj__ * L.ni);             \
  const int32_t j = L.j_lo + jj__;                                              \
  const int64_t idx = L.at(i, j);                                               \
  (void)i; (void)j

template <class FT>
__global__ void __launch_bounds__(256)
sea_ice_ocean_kernel(const __grid_constant__ NeSeaIceOceanDesc d, const __grid_constant__ Layout L, int64_t plane) {
  NE_POINT_INDEX();
  FT* T = (FT*)d.T;
  const FT* S = (const FT*)d.S;
  const FT* dz = (const FT*)d.dz;
  const FT rho = (FT)d.ocean.reference_density, c = (FT)d.ocean.heat_capacity;
  const FT slope = (FT)d.ocean.liquidus_slope, Tfresh = (FT)d.ocean.liquidus_freshwater_melting_temperature;
  const FT dt = (FT)d.dt;
  FT dQ = 0;
  FT TN = 0, SN = 0;
  for (int64_t k = d.nz; k >= 1; --k) {  // frazil clamp of the whole column (:155-178)
    const int64_t a = idx + (k + d.hz - 1) * plane;
    FT Tk = T[a], Sk = __ldg(S + a);
    FT Tm = Tfresh - slope * Sk;  // melting_temperature(LinearLiquidus, S)
    bool freezing = Tk < Tm;
    FT dE = freezing ? rho * c * (Tm - Tk) : (FT)0;  // Bool * x, false is a strong zero
    FT Tnew = freezing ? Tm : Tk;
    if (freezing) T[a] = Tnew;
    if (k == d.nz) { TN = Tnew; SN = Sk; }
    dQ -= dE * __ldg(dz + (k - 1)) / dt;
  }
  ((FT*)d.frazil_heat)[idx] = dQ;
  if (d.formulation == NE_SIO_FREEZE_ONLY) return;

  const FT E = (FT)d.latent_heat;
  FT Si = slot_at<FT>(d.ice_salinity, idx);
  FT hi = slot_at<FT>(d.hi, idx), conc = slot_at<FT>(d.concentration, idx), hc = slot_at<FT>(d.hc, idx);
  FT ustar;
  if (d.friction_velocity_kind == NE_USTAR_CONSTANT) ustar = (FT)d.friction_velocity;
  else {  // friction_velocity.jl:26-27,44
    const FT* tx = (const FT*)d.x_momentum_in;
    const FT* ty = (const FT*)d.y_momentum_in;
    FT ax = (sq(__ldg(tx + idx)) + sq(__ldg(tx + idx + 1))) / 2;
    FT ay = (sq(__ldg(ty + idx)) + sq(__ldg(ty + idx + L.sx))) / 2;
    ustar = m_sqrt(m_sqrt(ax + ay) / rho);
  }
  FT Q;
  if (d.formulation == NE_SIO_ICE_BATH) {  // heat_flux_formulations.jl:176-195
    FT Tm = Tfresh - slope * SN;
    Q = rho * c * (FT)d.heat_transfer_coefficient * ustar * (TN - Tm) * conc;
  } else {  // three-equation :224-313
    FT ah = (FT)d.heat_transfer_coefficient, as = (FT)d.salt_transfer_coefficient;
    FT kap = 0, Tsi = 0;
    if (d.has_conductive_flux) {  // :251-260
      kap = (hi >= hc) ? (FT)d.conductivity / (hi * E) : (FT)0;
      Tsi = __ldg((const FT*)d.internal_temperature + idx);
    }
    FT l1 = -slope, l2 = Tfresh;
    FT eta = rho * c * ah * ustar / E;
    FT gam = rho * as * ustar;
    FT th = eta + kap;
    FT qa = th * l1;
    FT qb = -gam - eta * TN - kap * Tsi + th * (l2 - l1 * Si);
    FT qc = gam * SN + (eta * TN + kap * Tsi - th * l2) * Si;
    FT xi = (qa == 0) ? (FT)0 : 1 / (2 * qa);
    FT Dl = mx(sq(qb) - 4 * qa * qc, (FT)0);
    FT rD = m_sqrt(Dl);
    FT Ss = (-qb - rD) * xi;
    Ss = (Ss < 0) ? (-qb + rD) * xi : Ss;
    FT Ts = Tfresh - slope * Ss;
    FT q = eta * (TN - Ts) + kap * (Tsi - Ts);
    Q = E * q * conc;
    ((FT*)d.interface_temperature)[idx] = Ts;
    ((FT*)d.interface_salinity)[idx] = Ss;
  }
  ((FT*)d.interface_heat)[idx] = Q;
  FT Ei = slot_at<FT>(d.ice_mass_flux, idx), Es = slot_at<FT>(d.snow_mass_flux, idx);
  ((FT*)d.freshwater)[idx] = -(Ei + Es) / rho;
  ((FT*)d.salt)[idx] = Ei * Si / rho;
}

template <class FT>
__global__ void __launch_bounds__(256)
sea_ice_ocean_stress_kernel(const __grid_constant__ NeSeaIceOceanStressDesc d, const __grid_constant__ Layout L) {
  NE_POINT_INDEX();
  const FT* ui = (const FT*)d.ui; const FT* vi = (const FT*)d.vi;
  const FT* uo = (const FT*)d.uo; const FT* vo = (const FT*)d.vo;
  const FT rho = (FT)d.ocean_density, Cd = (FT)d.drag_coefficient;
  const int64_t sx = L.sx;
  auto dvv = [&](int64_t a) { return __ldg(vo + a) - __ldg(vi + a); };  // Δu = uₑ − uᵢ (test/test_surface_fluxes.jl:334-335)
  auto duu = [&](int64_t a) { return __ldg(uo + a) - __ldg(ui + a); };
  FT du = duu(idx);
  FT dv4 = (dvv(idx) + dvv(idx - 1) + dvv(idx + sx) + dvv(idx - 1 + sx)) / 4;
  ((FT*)d.x_momentum)[idx] = rho * Cd * m_sqrt(sq(du) + sq(dv4)) * du;
  FT dv = dvv(idx);
  FT du4 = (duu(idx) + duu(idx + 1) + duu(idx - sx) + duu(idx + 1 - sx)) / 4;
  ((FT*)d.y_momentum)[idx] = rho * Cd * m_sqrt(sq(du4) + sq(dv)) * dv;
}

// one point of _assemble_net_ocean_fluxes! (Oceans/assemble_net_ocean_fluxes.jl:74-153)
template <class FT>
__device__ __forceinline__ void assemble_ocean_point(const NeAssembleOceanDesc& d, const Layout& L, int64_t idx) {
  const FT rho_inv = 1 / (FT)d.ocean.reference_density;
  const FT c_inv = 1 / (FT)d.ocean.heat_capacity;
  const int64_t sx = L.sx;
  FT conc = slot_at<FT>(d.concentration, idx);
  FT conc_w = slot_at<FT>(d.concentration, idx - 1), conc_s = slot_at<FT>(d.concentration, idx - sx);
  FT To = slot_at<FT>(d.ocean_surface_temperature, idx);
  FT Jrn = slot_at<FT>(d.rainfall, idx), Jsn = slot_at<FT>(d.snowfall, idx);
  FT Psn = slot_at<FT>(d.intercepted_snowfall, idx), Jln = slot_at<FT>(d.land_freshwater, idx);
  FT QT = slot_at<FT>(d.sensible_heat, idx), Qv = slot_at<FT>(d.latent_heat, idx), Jv = slot_at<FT>(d.water_vapor, idx);
  FT SQ = (QT + Qv) * (1 - conc);
  FT SF = -(Jrn + Jln + Jsn - Psn) * rho_inv + (1 - conc) * Jv * rho_inv;
  FT Jw_ao = -SF;
  const bool inactive = d.inactive ? (d.inactive[idx] != 0) : false;
  FT Qin = slot_at<FT>(d.interface_heat, idx), Js_io = slot_at<FT>(d.salt_io, idx), Jw_io = slot_at<FT>(d.freshwater_io, idx);
  FT JT_ao = SQ * rho_inv * c_inv;
  FT JT_io = Qin * rho_inv * c_inv;
  // τᶜᶜᶜ = ρ⁻¹ (1 - ℵ) ρτ averaged to the faces (:136-139)
  FT txc = rho_inv * (1 - conc) * slot_at<FT>(d.x_momentum_ao, idx);
  FT txw = rho_inv * (1 - conc_w) * slot_at<FT>(d.x_momentum_ao, idx - 1);
  FT tyc = rho_inv * (1 - conc) * slot_at<FT>(d.y_momentum_ao, idx);
  FT tys = rho_inv * (1 - conc_s) * slot_at<FT>(d.y_momentum_ao, idx - sx);
  FT tx_ao = (txw + txc) / 2;
  FT ty_ao = (tys + tyc) / 2;
  FT tx_io = slot_at<FT>(d.x_momentum_io, idx) * rho_inv * ((conc_w + conc) / 2);
  FT ty_io = slot_at<FT>(d.y_momentum_io, idx) * rho_inv * ((conc_s + conc) / 2);
  ((FT*)d.tau_x)[idx] = inactive ? (FT)0 : tx_ao + tx_io;
  ((FT*)d.tau_y)[idx] = inactive ? (FT)0 : ty_ao + ty_io;
  ((FT*)d.JT)[idx] = inactive ? (FT)0 : JT_ao + JT_io;
  ((FT*)d.JS)[idx] = inactive ? (FT)0 : Js_io;
  ((FT*)d.Jw)[idx] = inactive ? (FT)0 : Jw_ao + Jw_io;
  ((FT*)d.JH)[idx] = inactive ? (FT)0 : To * Jw_ao;
}

template <class FT>
__global__ void __launch_bounds__(256)
assemble_ocean_kernel(const __grid_constant__ NeAssembleOceanDesc d, const __grid_constant__ Layout L) {
  NE_POINT_INDEX();
  assemble_ocean_point<FT>(d, L, idx);
}

template <class FT>
__global__ void __launch_bounds__(256)
assemble_sea_ice_kernel(const __grid_constant__ NeAssembleSeaIceDesc d, const __grid_constant__ Layout L) {
  NE_POINT_INDEX();
  FT conc = slot_at<FT>(d.concentration, idx);
  FT QT = slot_at<FT>(d.sensible_heat, idx), Qv = slot_at<FT>(d.latent_heat, idx);
  FT Qf = slot_at<FT>(d.frazil_heat, idx), Qi = slot_at<FT>(d.interface_heat, idx);
  FT Jsn = slot_at<FT>(d.snowfall, idx);
  const bool inactive = d.inactive ? (d.inactive[idx] != 0) : false;
  FT tu = (slot_at<FT>(d.x_momentum, idx - 1) + slot_at<FT>(d.x_momentum, idx)) / 2;
  FT tv = (slot_at<FT>(d.y_momentum, idx - L.sx) + slot_at<FT>(d.y_momentum, idx)) / 2;
  ((FT*)d.top_heat)[idx] = inactive ? (FT)0 : (QT + Qv) * conc;
  ((FT*)d.top_snowfall)[idx] = inactive ? (FT)0 : Jsn;
  ((FT*)d.top_u)[idx] = inactive ? (FT)0 : tu;
  ((FT*)d.top_v)[idx] = inactive ? (FT)0 : tv;
  ((FT*)d.bottom_heat)[idx] = inactive ? (FT)0 : Qf + Qi;
}

// one point of _apply_air_sea[_ice]_radiative_fluxes! (Radiations/apply_air_sea_radiative_fluxes.jl:62-111,
// apply_air_sea_ice_radiative_fluxes.jl:55-90)
template <class FT>
__device__ __forceinline__ void apply_radiation_point(const NeApplyRadiationDesc& d, const Layout& L, int64_t idx, int32_t j) {
  FT conc = slot_at<FT>(d.concentration, idx);
  FT Ts = __ldg((const FT*)d.surface_temperature + idx);
  if (d.medium.temperature_units == NE_DEGREES_CELSIUS) Ts = Ts + (FT)273.15;
  RadState<FT> rs = radiation_state<FT>(d.radiation, L, idx, j);
  FT up = rs.sigma * rs.eps * pow4(Ts);   // radiation_kernels.jl:3
  FT ab = -rs.eps * rs.lw;                // :4
  FT tr = -(1 - rs.alpha) * rs.sw;        // :5
  const bool inactive = d.inactive ? (d.inactive[idx] != 0) : false;
  FT* H = (FT*)d.heat_flux;
  if (!d.over_sea_ice) {
    ab *= (1 - conc);
    tr *= (1 - conc);
    FT up_o = up * (1 - conc);
    FT Qss = tr;
    if (d.two_color) {  // Oceans/radiative_forcing.jl:84-91
      ((FT*)d.two_color_surface_flux)[idx] = -tr / ((FT)d.medium.reference_density * (FT)d.medium.heat_capacity);
      Qss = 0;
    }
    FT SQ = up_o + ab + Qss;
    FT JT = SQ * (1 / (FT)d.medium.reference_density) * (1 / (FT)d.medium.heat_capacity);
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

template <class FT>
__global__ void __launch_bounds__(256)
apply_radiation_kernel(const __grid_constant__ NeApplyRadiationDesc d, const __grid_constant__ Layout L) {
  NE_POINT_INDEX();
  apply_radiation_point<FT>(d, L, idx, j);
}

// ---- ElevationCorrection: _correct_atmosphere_elevation! (atmosphere_state_correction.jl:133-146) -------
template <class FT>
__global__ void __launch_bounds__(256)
elevation_correction_kernel(const __grid_constant__ NeElevationCorrectionDesc d, const __grid_constant__ Layout L) {
  NE_POINT_INDEX();
  (void)j;
  const FT dz = __ldg((const FT*)d.elevation_difference + idx);
  const FT dT = (FT)d.lapse_rate * dz;
  FT* T = (FT*)d.T;
  FT* p = (FT*)d.p;
  const FT T0 = T[idx];
  const FT Tbar = T0 - dT / 2;   // layer-mean temperature for the hydrostatic integral
  p[idx] = p[idx] * m_exp(-(FT)d.gravitational_acceleration * dz / ((FT)d.dry_air_gas_constant * Tbar));
  T[idx] = T0 - dT;              // lapse-rate shift; q is conserved
}

// ---- diagnostics: deterministic two-stage area-weighted sums (FP64 accumulation) ---------------------
// block-level stage: warp shuffle tree, then thread f adds the 8 warp sums of field f in warp order
template <int NF>
__device__ __forceinline__ void diag_block_reduce(const NeDiagDesc& d, const double (&acc)[NF], double (&sm)[8][NF]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    double v = acc[f];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[warp][f] = v;
  }
  __syncthreads();
  if (threadIdx.x < d.n_fields) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
    d.partial[(int64_t)blockIdx.x * d.n_fields + threadIdx.x] = v;
  }
}

// the per-point part of the first stage (shared by diag_partial_kernel and post_solve_kernel: same arithmetic, same
// point -> thread assignment, hence bit-identical sums)
template <class FT, int NF>
__device__ __forceinline__ void diag_accumulate(const NeDiagDesc& d, int64_t idx, double (&acc)[NF]) {
  const bool active = !(d.inactive && d.inactive[idx]);
  // every load is issued unconditionally (predicated on the field count only), then masked: the loads of
  // a point do not wait for its mask byte
  const double w = d.area ? (double)__ldg((const FT*)d.area + idx) : 1.0;
  double x[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) x[f] = f < d.n_fields ? (double)__ldcg((const FT*)d.fields[f] + idx) : 0.0;
  if (active) {
#pragma unroll
    for (int f = 0; f < NF; ++f)
      if (f < d.n_fields) acc[f] += w * x[f];
  }
}

// NF: compile-time bound on the field count (4, 8 or 16) so the accumulators of unused slots cost no registers
template <class FT, int NF>
__global__ void __launch_bounds__(256)
diag_partial_kernel(const __grid_constant__ NeDiagDesc d, const __grid_constant__ Layout L) {
  __shared__ double sm[8][NF];
  const int64_t n = (int64_t)L.ni * L.nj;
  double acc[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) acc[f] = 0;
  // fixed assignment of points to blocks/threads => run-to-run and rank-count independent order
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int32_t jj = (int32_t)(t / L.ni);
    const int64_t idx = L.at(L.i_lo + (int32_t)(t - (int64_t)jj * L.ni), L.j_lo + jj);
    diag_accumulate<FT, NF>(d, idx, acc);
  }
  diag_block_reduce<NF>(d, acc, sm);
}

// one warp per field: lane l sums blocks l, l+32, … in order, then a fixed shuffle tree — the order depends
// only on n_blocks, never on scheduling
__global__ void __launch_bounds__(32 * NE_DIAG_MAX_FIELDS)
diag_final_kernel(const double* __restrict__ partial, int64_t n_blocks, int n_fields, double* __restrict__ result) {
  const int f = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (f >= n_fields) return;
  double v = 0;
  for (int64_t b = lane; b < n_blocks; b += 32) v += partial[b * n_fields + f];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) result[f] = v;
}

// ---- phases 3-4 of update_state! in one pass: net ocean flux assembly, radiative flux application and (optionally)
// the first stage of the diagnostics sums.  Grid-stride over the launch range with the diagnostics' fixed block
// count, so a point is summed by the same thread, in the same order, as in diag_partial_kernel.  The fields the
// diagnostics read were written (or read) by this very thread a few instructions earlier: they come from L1/L2,
// not from HBM — the separate reduction pass over 7 fields x 58 MB (1/12 degree) disappears.
template <class FT, int NF, bool RAD, bool DIAG, int MINB>
__global__ void __launch_bounds__(256, MINB)
post_solve_kernel(const __grid_constant__ NeAssembleOceanDesc da, const __grid_constant__ NeApplyRadiationDesc dr,
                  const __grid_constant__ NeDiagDesc dd, const __grid_constant__ Layout L) {
  __shared__ double sm[8][NF];
  const int64_t n = (int64_t)L.ni * L.nj;
  double acc[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) acc[f] = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int32_t jj = (int32_t)(t / L.ni);
    const int32_t j = L.j_lo + jj;
    const int64_t idx = L.at(L.i_lo + (int32_t)(t - (int64_t)jj * L.ni), j);
    assemble_ocean_point<FT>(da, L, idx);
    if (RAD) apply_radiation_point<FT>(dr, L, idx, j);
    if (DIAG) diag_accumulate<FT, NF>(dd, idx, acc);
  }
  if (DIAG) diag_block_reduce<NF>(dd, acc, sm);
}

template <class K, class D>
static int launch_points(K kernel, const D& d, int stencil_lo, int stencil_hi, cudaStream_t s, const char* name) {
  NE_REQUIRE(grid_ok(d.grid, stencil_lo, stencil_hi), "%s: launch range (+stencil) leaves the parent array", name);
  Layout L = make_layout(d.grid);
  const int64_t n = (int64_t)L.ni * L.nj;
  kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d, L);
  NE_CUDA_CHECK_LAUNCH(name);
  return NE_OK;
}

template <class FT>
static int sio_entry(const NeSeaIceOceanDesc* d, void* stream) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  NE_REQUIRE(grid_ok(d->grid, 0, d->friction_velocity_kind == NE_USTAR_MOMENTUM_BASED ? 1 : 0), "sea-ice-ocean: launch range leaves the parent array");
  NE_REQUIRE(d->T && d->S && d->dz && d->nz >= 1 && d->hz >= 0, "sea-ice-ocean: null column arrays");
  NE_REQUIRE(d->frazil_heat != nullptr, "sea-ice-ocean: null frazil_heat");
  if (d->formulation != NE_SIO_FREEZE_ONLY) {
    if (d->formulation != NE_SIO_ICE_BATH && d->formulation != NE_SIO_THREE_EQUATION)
      NE_NO_VARIANT("sea-ice-ocean heat flux formulation %d has no kernel variant", d->formulation);
    NE_REQUIRE(d->interface_heat && d->salt && d->freshwater, "sea-ice-ocean: null flux outputs");
    if (d->formulation == NE_SIO_THREE_EQUATION)
      NE_REQUIRE(d->interface_temperature && d->interface_salinity, "three-equation: null interface T/S");
    if (d->friction_velocity_kind == NE_USTAR_MOMENTUM_BASED)
      NE_REQUIRE(d->x_momentum_in && d->y_momentum_in, "momentum-based friction velocity needs the stresses");
    else if (d->friction_velocity_kind != NE_USTAR_CONSTANT) NE_NO_VARIANT("friction velocity formulation with no kernel variant");
    if (d->has_conductive_flux) NE_REQUIRE(d->internal_temperature != nullptr, "conductive flux needs internal_temperature");
  }
  Layout L = make_layout(d->grid);
  const int64_t n = (int64_t)L.ni * L.nj;
  const int64_t plane = L.sx * (d->grid.ny + 2 * d->grid.hy);
  sea_ice_ocean_kernel<FT><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*d, L, plane);
  NE_CUDA_CHECK_LAUNCH("ne_sea_ice_ocean_fluxes");
  return NE_OK;
}

template <class FT>
static int diag_entry(const NeDiagDesc* d, void* stream) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  NE_REQUIRE(grid_ok(d->grid, 0, 0), "diag: launch range leaves the parent array");
  NE_REQUIRE(d->n_fields >= 1 && d->n_fields <= NE_DIAG_MAX_FIELDS, "diag: n_fields out of range");
  NE_REQUIRE(d->partial && d->result && d->n_blocks >= 1, "diag: null scratch/result");
  for (int f = 0; f < d->n_fields; ++f) NE_REQUIRE(d->fields[f] != nullptr, "diag: null field");
  Layout L = make_layout(d->grid);
  if (d->n_fields <= 4) diag_partial_kernel<FT, 4><<<(unsigned)d->n_blocks, 256, 0, (cudaStream_t)stream>>>(*d, L);
  else if (d->n_fields <= 8) diag_partial_kernel<FT, 8><<<(unsigned)d->n_blocks, 256, 0, (cudaStream_t)stream>>>(*d, L);
  else diag_partial_kernel<FT, NE_DIAG_MAX_FIELDS><<<(unsigned)d->n_blocks, 256, 0, (cudaStream_t)stream>>>(*d, L);
  NE_CUDA_CHECK_LAUNCH("ne_diag_reduce(partial)");
  diag_final_kernel<<<1, 32 * NE_DIAG_MAX_FIELDS, 0, (cudaStream_t)stream>>>(d->partial, d->n_blocks, d->n_fields, d->result);
  NE_CUDA_CHECK_LAUNCH("ne_diag_reduce(final)");
  return NE_OK;
}

static bool same_range(const NeExchangeGrid& a, const NeExchangeGrid& b) {
  return a.nx == b.nx && a.ny == b.ny && a.hx == b.hx && a.hy == b.hy && a.i_lo == b.i_lo && a.i_hi == b.i_hi &&
         a.j_lo == b.j_lo && a.j_hi == b.j_hi;
}

// Phases 3-4 (+ diagnostics) of the fused step as one kernel.  Returns NE_OK when enqueued, +1 when the descriptors
// do not qualify (the caller then enqueues the component kernels), < 0 on error.
template <class FT>
static int post_solve(const NeFusedStepDesc* d, void* stream) {
  const char* off = std::getenv("NE_B200_NO_POST_SOLVE_FUSION");
  if (off && off[0] == '1') return 1;
  const NeAssembleOceanDesc& a = d->assemble;
  const NeApplyRadiationDesc& r = d->apply_radiation;
  const NeDiagDesc& g = d->diag;
  const bool rad = r.radiation.enabled != 0, diag = g.n_fields > 0;
  if (!rad && !diag) return 1;
  if (rad && (!same_range(a.grid, r.grid) || r.over_sea_ice || r.heat_flux != a.JT)) return 1;
  if (diag && !same_range(a.grid, g.grid)) return 1;
  NE_REQUIRE(a.tau_x && a.tau_y && a.JT && a.JS && a.Jw && a.JH, "assemble ocean: null output");
  NE_REQUIRE(grid_ok(a.grid, 1, 0), "assemble ocean: launch range (+stencil) leaves the parent array");
  if (rad) {
    NE_REQUIRE(r.surface_temperature && r.heat_flux && r.upwelling_longwave && r.downwelling_longwave && r.downwelling_shortwave, "apply radiation: null array");
    NE_REQUIRE(!r.two_color || r.two_color_surface_flux, "apply radiation: two_color without surface_flux array");
  }
  if (diag) {
    NE_REQUIRE(g.n_fields <= NE_DIAG_MAX_FIELDS, "diag: n_fields out of range");
    NE_REQUIRE(g.partial && g.result && g.n_blocks >= 1, "diag: null scratch/result");
    for (int f = 0; f < g.n_fields; ++f) NE_REQUIRE(g.fields[f] != nullptr, "diag: null field");
  }
  Layout L = make_layout(a.grid);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = (int64_t)L.ni * L.nj;
  const unsigned blocks = diag ? (unsigned)g.n_blocks : (unsigned)((n + 255) / 256);
  // 64 registers (4 CTAs per SM): 0.29 ms; 80 regs 0.35, 122 regs 0.39, 48 regs (spills) 0.33 ms (profiles/r01_notes.md)
#define NE_POST(NF, RAD, DIAG) post_solve_kernel<FT, NF, RAD, DIAG, 4><<<blocks, 256, 0, s>>>(a, r, g, L)
  if (!diag) NE_POST(4, true, false);
  else if (g.n_fields <= 4) { if (rad) NE_POST(4, true, true); else NE_POST(4, false, true); }
  else if (g.n_fields <= 8) { if (rad) NE_POST(8, true, true); else NE_POST(8, false, true); }
  else { if (rad) NE_POST(NE_DIAG_MAX_FIELDS, true, true); else NE_POST(NE_DIAG_MAX_FIELDS, false, true); }
#undef NE_POST
  NE_CUDA_CHECK_LAUNCH("ne_fused_interface_step(assembly+radiation+diagnostics)");
  if (diag) {
    diag_final_kernel<<<1, 32 * NE_DIAG_MAX_FIELDS, 0, s>>>(g.partial, g.n_blocks, g.n_fields, g.result);
    NE_CUDA_CHECK_LAUNCH("ne_fused_interface_step(diagnostics final)");
  }
  return NE_OK;
}

int post_solve_f64(const NeFusedStepDesc* d, void* stream) { return post_solve<double>(d, stream); }
int post_solve_f32(const NeFusedStepDesc* d, void* stream) { return post_solve<float>(d, stream); }

}  // namespace ne

extern "C" {
int ne_sea_ice_ocean_fluxes_f64(const NeSeaIceOceanDesc* d, void* s) { NE_NVTX(); return ne::sio_entry<double>(d, s); }
int ne_sea_ice_ocean_fluxes_f32(const NeSeaIceOceanDesc* d, void* s) { NE_NVTX(); return ne::sio_entry<float>(d, s); }

int ne_sea_ice_ocean_stress_f64(const NeSeaIceOceanStressDesc* d, void* s) { NE_NVTX();
  NE_REQUIRE(d && d->ui && d->vi && d->uo && d->vo && d->x_momentum && d->y_momentum, "stress: null array");
  return ne::launch_points(ne::sea_ice_ocean_stress_kernel<double>, *d, 1, 1, (cudaStream_t)s, "ne_sea_ice_ocean_stress");
}
int ne_sea_ice_ocean_stress_f32(const NeSeaIceOceanStressDesc* d, void* s) { NE_NVTX();
  NE_REQUIRE(d && d->ui && d->vi && d->uo && d->vo && d->x_momentum && d->y_momentum, "stress: null array");
  return ne::launch_points(ne::sea_ice_ocean_stress_kernel<float>, *d, 1, 1, (cudaStream_t)s, "ne_sea_ice_ocean_stress");
}
int ne_assemble_net_ocean_fluxes_f64(const NeAssembleOceanDesc* d, void* s) { NE_NVTX();
  NE_REQUIRE(d && d->tau_x && d->tau_y && d->JT && d->JS && d->Jw && d->JH, "assemble ocean: null output");
  return ne::launch_points(ne::assemble_ocean_kernel<double>, *d, 1, 0, (cudaStream_t)s, "ne_assemble_net_ocean_fluxes");
}
int ne_assemble_net_ocean_fluxes_f32(const NeAssembleOceanDesc* d, void* s) { NE_NVTX();
  NE_REQUIRE(d && d->tau_x && d->tau_y && d->JT && d->JS && d->Jw && d->JH, "assemble ocean: null output");
  return ne::launch_points(ne::assemble_ocean_kernel<float>, *d, 1, 0, (cudaStream_t)s, "ne_assemble_net_ocean_fluxes");
}
int ne_assemble_net_sea_ice_fluxes_f64(const NeAssembleSeaIceDesc* d, void* s) { NE_NVTX();
  NE_REQUIRE(d && d->top_heat && d->top_snowfall && d->top_u && d->top_v && d->bottom_heat, "assemble sea ice: null output");
  return ne::launch_points(ne::assemble_sea_ice_kernel<double>, *d, 1, 0, (cudaStream_t)s, "ne_assemble_net_sea_ice_fluxes");
}
int ne_assemble_net_sea_ice_fluxes_f32(const NeAssembleSeaIceDesc* d, void* s) { NE_NVTX();
  NE_REQUIRE(d && d->top_heat && d->top_snowfall && d->top_u && d->top_v && d->bottom_heat, "assemble sea ice: null output");
  return ne::launch_points(ne::assemble_sea_ice_kernel<float>, *d, 1, 0, (cudaStream_t)s, "ne_assemble_net_sea_ice_fluxes");
}
int ne_apply_radiative_fluxes_f64(const NeApplyRadiationDesc* d, void* s) { NE_NVTX();
  NE_REQUIRE(d && d->surface_temperature && d->heat_flux && d->upwelling_longwave && d->downwelling_longwave && d->downwelling_shortwave, "apply radiation: null array");
  NE_REQUIRE(!d->two_color || d->two_color_surface_flux, "apply radiation: two_color without surface_flux array");
  return ne::launch_points(ne::apply_radiation_kernel<double>, *d, 0, 0, (cudaStream_t)s, "ne_apply_radiative_fluxes");
}
int ne_apply_radiative_fluxes_f32(const NeApplyRadiationDesc* d, void* s) { NE_NVTX();
  NE_REQUIRE(d && d->surface_temperature && d->heat_flux && d->upwelling_longwave && d->downwelling_longwave && d->downwelling_shortwave, "apply radiation: null array");
  NE_REQUIRE(!d->two_color || d->two_color_surface_flux, "apply radiation: two_color without surface_flux array");
  return ne::launch_points(ne::apply_radiation_kernel<float>, *d, 0, 0, (cudaStream_t)s, "ne_apply_radiative_fluxes");
}
int ne_correct_atmosphere_elevation_f64(const NeElevationCorrectionDesc* d, void* s) { NE_NVTX();
  NE_REQUIRE(d && d->T && d->p && d->elevation_difference, "elevation correction: null array");
  return ne::launch_points(ne::elevation_correction_kernel<double>, *d, 0, 0, (cudaStream_t)s, "ne_correct_atmosphere_elevation");
}
int ne_correct_atmosphere_elevation_f32(const NeElevationCorrectionDesc* d, void* s) { NE_NVTX();
  NE_REQUIRE(d && d->T && d->p && d->elevation_difference, "elevation correction: null array");
  return ne::launch_points(ne::elevation_correction_kernel<float>, *d, 0, 0, (cudaStream_t)s, "ne_correct_atmosphere_elevation");
}
int ne_diag_reduce_f64(const NeDiagDesc* d, void* s) { NE_NVTX(); return ne::diag_entry<double>(d, s); }
int ne_diag_reduce_f32(const NeDiagDesc* d, void* s) { NE_NVTX(); return ne::diag_entry<float>(d, s); }
}
