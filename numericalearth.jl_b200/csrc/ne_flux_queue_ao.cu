// ne_flux_queue_ao.cu — work-queue kernel, atmosphere–ocean default tree (Float64 opt-in, Float32 default): launchers.
#include "ne_flux_queue.cuh"
#include "ne_queue_host.cuh"

namespace ne {

constexpr int QUEUE_SLOTS = 1024;
struct QueueCounters { int device; uint32_t* dptr; unsigned next; };

uint32_t* queue_counters() {
  static std::mutex mutex;
  static std::vector<QueueCounters> pools;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mutex);
  for (QueueCounters& q : pools)
    if (q.device == dev) return q.dptr + 2 * (q.next++ % QUEUE_SLOTS);
  QueueCounters q = {dev, nullptr, 1};
  if (cudaMalloc(&q.dptr, sizeof(uint32_t) * 2 * QUEUE_SLOTS) != cudaSuccess ||
      cudaMemset(q.dptr, 0, sizeof(uint32_t) * 2 * QUEUE_SLOTS) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  pools.push_back(q);
  return q.dptr;
}

template <class FT, class CT, bool HS>
static int launch_queue_hs(const NeAtmosOceanDesc& d, const TabParams& T, const double* tab, cudaStream_t s) {
  using Problem = AoProblem<FT, CT, HS>;
  const bool f32 = std::is_same<FT, float>::value;
  typename Problem::Params prm;
  prm.d = d;
  prm.L = make_layout(d.grid);
  prm.th = Thermo<CT>::make(d.thermo);
  prm.P = make_fast_params(d.flux, d.gravitational_acceleration, f32);
  prm.Q = make_front_f32(d.flux, d.gravitational_acceleration);
  prm.T = T;
  prm.T.far_fm = !T.general_psi && far_unstable_fm_ok(prm.P);
  prm.T.log_hd = f32 ? std::log((double)((float)d.surface_layer_height.value - prm.Q.d_zero))
                     : std::log(d.surface_layer_height.value - prm.P.d_zero);
  uint32_t* counters = queue_counters();
  NE_REQUIRE(counters != nullptr, "atmosphere-ocean: could not allocate the work-queue counters");
  if (cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(uint32_t), s); e != cudaSuccess)   // see queue_counters()
    return cuda_error(e, "work-queue kernel (counter reset)");
  const unsigned grid = queue_grid((int64_t)prm.L.ni * prm.L.nj, 8, 3);
  if (cudaError_t e = allow_table_smem<flux_queue_kernel<Problem, 8, 3>>(); e != cudaSuccess) return cuda_error(e, "work-queue kernel (shared memory opt-in)");
  flux_queue_kernel<Problem, 8, 3><<<grid, 256, TAB_SMEM_BYTES, s>>>(prm, tab, queue_theta(), counters);
  NE_CUDA_CHECK_LAUNCH("ne_atmosphere_ocean_fluxes(queue)");
  return NE_OK;
}

template <class FT, class CT>
int launch_queue(const NeAtmosOceanDesc& d, const TabParams& T, const double* tab, cudaStream_t s) {
  const bool hs = !d.surface_layer_height.ptr && !d.boundary_layer_height.ptr;
  return hs ? launch_queue_hs<FT, CT, true>(d, T, tab, s) : launch_queue_hs<FT, CT, false>(d, T, tab, s);
}
template int launch_queue<double, double>(const NeAtmosOceanDesc&, const TabParams&, const double*, cudaStream_t);
template int launch_queue<double, float>(const NeAtmosOceanDesc&, const TabParams&, const double*, cudaStream_t);
template int launch_queue<float, float>(const NeAtmosOceanDesc&, const TabParams&, const double*, cudaStream_t);

}  // namespace ne
