// ne_flux_queue_land.cu — work-queue kernel, atmosphere–land default tree (Float64): launchers.
#include "ne_flux_land_fast.cuh"
#include "ne_queue_host.cuh"

namespace ne {

template <class CT, bool HS>
static int launch_land_queue_hs(const NeAtmosLandDesc& d, const TabParams& T, const double* tab, cudaStream_t s) {
  using Problem = LandProblem<CT, HS>;
  typename Problem::Params prm;
  prm.d = d;
  prm.L = make_layout(d.grid);
  prm.th = Thermo<CT>::make(d.thermo);
  prm.P = make_fast_params(d.flux, d.gravitational_acceleration, false);
  prm.T = T;
  prm.T.far_fm = !T.general_psi && far_unstable_fm_ok(prm.P);
  prm.T.log_hd = std::log(d.surface_layer_height.value - prm.P.d_zero);
  uint32_t* counters = queue_counters();
  NE_REQUIRE(counters != nullptr, "atmosphere-land: could not allocate the work-queue counters");
  if (cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(uint32_t), s); e != cudaSuccess)   // see queue_counters()
    return cuda_error(e, "work-queue kernel (counter reset)");
  const unsigned grid = queue_grid((int64_t)prm.L.ni * prm.L.nj, 4, 4);
  if (cudaError_t e = allow_table_smem<flux_queue_kernel<Problem, 4, 4>>(); e != cudaSuccess) return cuda_error(e, "work-queue kernel (shared memory opt-in)");
  flux_queue_kernel<Problem, 4, 4><<<grid, 128, TAB_SMEM_BYTES, s>>>(prm, tab, queue_theta(), counters);
  NE_CUDA_CHECK_LAUNCH("ne_atmosphere_land_fluxes(queue)");
  return NE_OK;
}

template <class CT>
int launch_land_queue(const NeAtmosLandDesc& d, const TabParams& T, const double* tab, cudaStream_t s) {
  const bool hs = !d.surface_layer_height.ptr && !d.boundary_layer_height.ptr;
  return hs ? launch_land_queue_hs<CT, true>(d, T, tab, s) : launch_land_queue_hs<CT, false>(d, T, tab, s);
}
template int launch_land_queue<double>(const NeAtmosLandDesc&, const TabParams&, const double*, cudaStream_t);
template int launch_land_queue<float>(const NeAtmosLandDesc&, const TabParams&, const double*, cudaStream_t);

}  // namespace ne
