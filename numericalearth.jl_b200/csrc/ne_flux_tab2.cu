// ne_flux_tab2.cu — the a–o solve of the default plugin tree in Float64 as shipped in round 2 (sm_100a).
//
// Replaces _compute_atmosphere_ocean_interface_state! (InterfaceComputations/atmosphere_ocean_fluxes.jl:80-197) for
// SimilarityTheoryFluxes defaults + BulkTemperature + scalar surface/boundary-layer heights; every other tree keeps
// the kernels of ne_flux_kernels.cu.  See ne_flux_tab2.cuh for what changed in the iteration.
//
// Launch shape: 256-thread CTAs, 3 per SM (80 registers); a CTA stages the 27 KB solver table in shared memory once
// and walks 1024-point windows of the launch range with a grid stride; inside a window the 32 groups of 32 points
// are dealt to the 8 warps in a snake order (w, 15−w, 16+w, 31−w) and the warps never synchronise again.
//
// Trip-count ordering.  A warp iterates until its slowest lane has converged: with the points in memory order 18 %
// of the lane-trips of C4 are idle lanes (trip counts 7–24, mean 13.6).  The trip count of a point changes little
// from one coupled step to the next, and the kernel stores it (`iterations`), so before the solve
// `trip_order_kernel` counting-sorts every 1024-point window by the PREVIOUS step's count (a permutation of window
// offsets, 2 bytes per point, in a library-owned scratch buffer keyed by the `iterations` pointer) and lane l of
// group g takes the point perm[32 g + l]: idle lane-trips fall to 3 %.  Which lane computes a point does not change
// its arithmetic: results are bit-identical with and without the ordering (tested), and a stale or all-zero
// `iterations` array (first step) only costs the speed-up.  Loads and stores of a group are scattered over the 8 KB
// (per field) of its window instead of 256 contiguous bytes: L1/L2 absorb it, DRAM traffic stays the algorithmic one.
#include <deque>
#include <mutex>

#include "ne_flux_tab2.cuh"
#include "ne_queue_host.cuh"

namespace ne {

constexpr int TAB2_WINDOW = 1024;
constexpr int TAB2_KEYS = 64;

__global__ void __launch_bounds__(256)
trip_order_kernel(const int32_t* __restrict__ iterations, const __grid_constant__ Layout L, const uint32_t n,
                  uint16_t* __restrict__ perm) {
  __shared__ int hist[TAB2_KEYS];
  __shared__ int base[TAB2_KEYS];
  const int tid = threadIdx.x;
  if (tid < TAB2_KEYS) hist[tid] = 0;
  __syncthreads();
  const uint32_t t0 = blockIdx.x * (uint32_t)TAB2_WINDOW;
  int key[4], rank[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t t = t0 + tid + 256 * k;
    int kk = TAB2_KEYS - 1;                    // beyond the launch range: last
    if (t < n) {
      const uint32_t jj = t / (uint32_t)L.ni;
      const int v = iterations[L.at(L.i_lo + (int32_t)(t - jj * (uint32_t)L.ni), L.j_lo + (int32_t)jj)];
      kk = v < 0 ? 0 : (v > TAB2_KEYS - 2 ? TAB2_KEYS - 2 : v);
    }
    key[k] = kk;
    rank[k] = atomicAdd(&hist[kk], 1);
  }
  __syncthreads();
  if (tid < 32) {   // exclusive scan of the 64 bins by one warp
    const int a = hist[2 * tid], b = hist[2 * tid + 1];
    int s = a + b;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, s, off);
      if (tid >= off) s += v;
    }
    base[2 * tid] = s - a - b;
    base[2 * tid + 1] = s - b;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) perm[t0 + base[key[k]] + rank[k]] = (uint16_t)(tid + 256 * k);
}

template <class CT, bool SORT, bool FMPRO, class O>
__global__ void __launch_bounds__(256, 3)
ao_flux_tab2_kernel(const __grid_constant__ NeAtmosOceanDesc d, const __grid_constant__ Layout L,
                    const __grid_constant__ Thermo<CT> th, const __grid_constant__ FastParams P,
                    const __grid_constant__ TabParams T, const double* __restrict__ gtab,
                    const uint16_t* __restrict__ perm, unsigned long long* __restrict__ counts) {
  __shared__ __align__(16) double tab[fm::TAB_SIZE];
  for (int k = threadIdx.x; k < fm::TAB_SIZE / 2; k += 256)
    reinterpret_cast<double2*>(tab)[k] = __ldg(reinterpret_cast<const double2*>(gtab) + k);
  __syncthreads();
  O o;
  const uint32_t n = (uint32_t)L.ni * (uint32_t)L.nj;
  const uint32_t n_windows = (n + TAB2_WINDOW - 1) / TAB2_WINDOW;
  const bool celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  const bool relative = d.properties.velocity_formulation == NE_VEL_RELATIVE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long warp_trips = 0;
  for (uint32_t w = blockIdx.x; w < n_windows; w += gridDim.x) {
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
      // snake: every warp gets one group of each quarter of the (sorted) window
      const int g = (r == 0) ? warp : (r == 1) ? 15 - warp : (r == 2) ? 16 + warp : 31 - warp;
      const uint32_t slot = w * (uint32_t)TAB2_WINDOW + (uint32_t)(g * 32 + lane);
      const uint32_t t = SORT ? w * (uint32_t)TAB2_WINDOW + perm[slot] : slot;
      if (t >= n) continue;
      const uint32_t jj = t / (uint32_t)L.ni;
      const int64_t idx = L.at(L.i_lo + (int32_t)(t - jj * (uint32_t)L.ni), L.j_lo + (int32_t)jj);
      const bool not_water = d.inactive ? (d.inactive[idx] != 0) : false;
      const bool skip = not_water && !P.fixed;   // needs_to_converge && not_water (atmosphere_ocean_fluxes.jl:144)
      double ustar = 0, theta_star = 0, q_star = 0;
      int iters = 0;
      if (!skip) {
        FastPoint s;
        tab2_prologue<O, CT, true, FMPRO>(o, d, L, th, P, T, tab, idx, celsius, relative, s);
        iters = tab2_solve(o, P, T, tab, s, counts);
        ustar = s.ustar; theta_star = s.theta_star; q_star = s.q_star;
        if (!std::is_same<O, fm::OpsPlain>::value) {
          const int mx = __reduce_max_sync(__activemask(), iters);
          if (lane == __ffs(__activemask()) - 1) warp_trips += mx;
        }
      }
      tab2_epilogue<O, CT, FMPRO>(o, d, L, th, idx, celsius, relative, not_water, ustar, theta_star, q_star, iters);
    }
  }
  if (!std::is_same<O, fm::OpsPlain>::value) {
    o.flush(counts);
    if (counts && warp_trips) atomicAdd(counts + 5, warp_trips);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------
// Scratch permutation buffers, one per (device, iterations array, launch extent), allocated on first use (outside
// CUDA-graph capture, like the solver tables) and kept for the life of the process.
struct OrderBuf { int device; const void* key; uint32_t n; uint16_t* perm; };
static std::mutex g_order_mutex;
static std::deque<OrderBuf> g_order;

static uint16_t* order_buffer(const void* key, uint32_t n) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(g_order_mutex);
  for (const OrderBuf& b : g_order)
    if (b.device == dev && b.key == key && b.n == n) return b.perm;
  if (g_order.size() >= 64) {   // a host that keeps re-allocating its `iterations` array: start over
    for (OrderBuf& b : g_order) cudaFree(b.perm);
    g_order.clear();
  }
  OrderBuf b{dev, key, n, nullptr};
  const size_t windows = ((size_t)n + TAB2_WINDOW - 1) / TAB2_WINDOW;
  if (cudaMalloc(&b.perm, windows * TAB2_WINDOW * sizeof(uint16_t)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  g_order.push_back(b);
  return b.perm;
}

static unsigned tab2_grid(uint32_t n_windows) {
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int waves = std::max(1, env_int("NE_B200_TAB_WAVES", 4));
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(n_windows, (int64_t)sms * 3 * waves));
}

bool tab2_eligible(const NeAtmosOceanDesc& d, const TabParams& TP) {
  if (env_flag("NE_B200_TAB_V1")) return false;
  if (TP.general_psi || d.surface_layer_height.ptr || d.boundary_layer_height.ptr) return false;
  const int64_t n = (int64_t)(d.grid.i_hi - d.grid.i_lo + 1) * (int64_t)(d.grid.j_hi - d.grid.j_lo + 1);
  return n > 0 && n < ((int64_t)1 << 31) - TAB2_WINDOW;
}

// counts != nullptr: the counting instantiation (6 device counters: fma, mul, add, other, thread trips, warp trips)
int launch_tab2(const NeAtmosOceanDesc& d, const Layout& L, const FastParams& P, const TabParams& TP, const double* tab,
                cudaStream_t s, unsigned long long* counts) {
  const uint32_t n = (uint32_t)L.ni * (uint32_t)L.nj;
  const uint32_t n_windows = (n + TAB2_WINDOW - 1) / TAB2_WINDOW;
  const bool ct64 = d.thermo.dtype == NE_F64;
  const bool fmpro = !env_flag("NE_B200_TAB2_LIBM_PROLOGUE");
  uint16_t* perm = nullptr;
  if (d.iterations && !env_flag("NE_B200_TAB2_NO_ORDER")) perm = order_buffer(d.iterations, n);
  if (perm) {
    trip_order_kernel<<<n_windows, 256, 0, s>>>(d.iterations, L, n, perm);
    NE_CUDA_CHECK_LAUNCH("ne_atmosphere_ocean_fluxes(trip order)");
  }
  const unsigned grid = tab2_grid(n_windows);
#define NE_TAB2(CT, SORT, FMPRO, O) \
  ao_flux_tab2_kernel<CT, SORT, FMPRO, O><<<grid, 256, 0, s>>>(d, L, Thermo<CT>::make(d.thermo), P, TP, tab, perm, counts)
#define NE_TAB2_O(CT, SORT, FMPRO)                      \
  do {                                                  \
    if (counts) NE_TAB2(CT, SORT, FMPRO, fm::OpsCount); \
    else NE_TAB2(CT, SORT, FMPRO, fm::OpsPlain);        \
  } while (0)
#define NE_TAB2_S(CT, FMPRO)                   \
  do {                                         \
    if (perm) NE_TAB2_O(CT, true, FMPRO);      \
    else NE_TAB2_O(CT, false, FMPRO);          \
  } while (0)
  if (fmpro) {
    if (ct64) NE_TAB2_S(double, true); else NE_TAB2_S(float, true);
  } else {
    if (counts) { set_error("the counting build covers the shipped prologue only"); return NE_E_INVALID; }
    if (ct64) { if (perm) NE_TAB2(double, true, false, fm::OpsPlain); else NE_TAB2(double, false, false, fm::OpsPlain); }
    else { if (perm) NE_TAB2(float, true, false, fm::OpsPlain); else NE_TAB2(float, false, false, fm::OpsPlain); }
  }
#undef NE_TAB2_S
#undef NE_TAB2_O
#undef NE_TAB2
  NE_CUDA_CHECK_LAUNCH("ne_atmosphere_ocean_fluxes(tab2)");
  return NE_OK;
}

}  // namespace ne
