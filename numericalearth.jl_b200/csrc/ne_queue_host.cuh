// ne_queue_host.cuh — host-side plumbing of the work-queue kernels (shared by the a–o and a–si launch TUs).
#pragma once

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "ne_flux_tab.cuh"

namespace ne {

inline bool env_flag(const char* name) {
  const char* v = std::getenv(name);
  return v && v[0] == '1';
}
inline int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

// The solver table (fm::TAB_SIZE doubles, 44 KB) is DYNAMIC shared memory in every kernel that stages it: together with
// a kernel's static shared memory it exceeds the 48 KB a kernel may use without opting in.  Once per kernel and device.
constexpr size_t TAB_SMEM_BYTES = sizeof(double) * fm::TAB_SIZE;
template <auto Kernel>
inline cudaError_t allow_table_smem(size_t bytes = TAB_SMEM_BYTES) {
  static std::atomic<unsigned long long> done{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_relaxed) & bit) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_relaxed);
  return e;
}

// the work-queue kernel indexes points with 32 bits
inline bool queue_path_ok(const NeExchangeGrid& g) {
  if (env_flag("NE_B200_TAB_CLASSIC")) return false;
  const int64_t parent = (int64_t)(g.nx + 2 * g.hx) * (int64_t)(g.ny + 2 * g.hy);
  return parent < ((int64_t)1 << 31);
}

inline FrontF32 make_front_f32(const NeFluxFormulation& flux, double gravitational_acceleration) {
  FrontF32 Q;
  Q.gmin = (float)flux.subgrid_velocities.minimum_gustiness;
  Q.beta = (float)flux.subgrid_velocities.gustiness_parameter;
  Q.Cg = (float)flux.ell_momentum.wave_constant;
  Q.g_rough = (float)flux.ell_momentum.gravitational_acceleration;
  Q.kappa = (float)flux.von_karman_constant;
  Q.tol = (float)flux.stop.tolerance;
  Q.g = (float)gravitational_acceleration;
  Q.d_zero = (float)flux.zero_plane_displacement;
  return Q;
}

// Tile/CTA counters of the work-queue kernel: a per-device pool of {tile, cta} pairs used round robin.  A kernel leaves
// its pair zeroed (last CTA out) AND every launcher zeroes the pair on its own stream right before the launch (a memset
// node under CUDA-graph capture): a launch that was aborted half-way cannot leave a later one with tile >= n_tiles and no
// outputs.  Two launches share a pair only if QUEUE_SLOTS launches are in flight at once (or a captured graph, pinned to
// the pair it was captured with, replays while the live launch that drew the same pair is still running).
uint32_t* queue_counters();   // ne_flux_queue_ao.cu

// grid of the persistent work-queue kernels: resident CTAs x NE_B200_QUEUE_WAVES, at most one CTA per 32*warps points
inline unsigned queue_grid(int64_t n, int warps, int ctas_per_sm) {
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ctas_needed = (n + 32 * warps - 1) / (32 * warps);
  const int waves = std::max(1, env_int("NE_B200_QUEUE_WAVES", 1));
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ctas_needed, (int64_t)sms * ctas_per_sm * waves));
}
inline int queue_theta() {
  const int theta = env_int("NE_B200_QUEUE_THETA", 12);
  return theta < 0 ? 0 : (theta > 24 ? 24 : theta);
}

// launchers (explicitly instantiated in ne_flux_queue_ao.cu / ne_flux_queue_asi.cu); `tab` = device solver table
template <class FT, class CT> int launch_queue(const NeAtmosOceanDesc& d, const TabParams& T, const double* tab, cudaStream_t s);
template <class CT> int launch_land_queue(const NeAtmosLandDesc& d, const TabParams& T, const double* tab, cudaStream_t s);
template <class FT, class CT> int launch_asi_queue(const NeAtmosSeaIceDesc& d, const TabParams& T, const double* tab, cudaStream_t s);

}  // namespace ne
