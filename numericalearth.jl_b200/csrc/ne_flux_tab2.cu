// ne_flux_tab2.cu — the a–o solve of the default plugin tree in Float64 as shipped in round 2 (sm_100a).
//
// Replaces _compute_atmosphere_ocean_interface_state! (InterfaceComputations/atmosphere_ocean_fluxes.jl:80-197) for
// SimilarityTheoryFluxes defaults + BulkTemperature + scalar surface/boundary-layer heights; every other tree keeps
// the kernels of ne_flux_kernels.cu.  See ne_flux_tab2.cuh for what changed in the iteration.
//
// Launch shape: one persistent 768-thread CTA per SM (24 warps, 80 registers); the CTA stages the 44 KB solver table in shared memory
// once, then each of its warps draws groups of 32 points from a global counter (so the load balances itself whatever the
// land mask and the trip counts look like) and the warps never synchronise again.
//
// Ordering of the points.  Two properties of a warp decide what it costs: (i) it iterates until its SLOWEST lane has
// converged — in memory order 18 % of the lane-trips of C4 are idle lanes (trip counts 7–24, mean 13.6); (ii) every trip
// gathers 11 x 16 bytes per lane from the lane's ψ-table record, and lanes on different records whose 16-byte chunks share
// banks serialise in the L1TEX data pipe.  Both the trip count and the table record of a point change little from one
// coupled step to the next, so the kernel leaves a 16-bit hint per point (trips << 8 | record of its last trip) in a
// library-owned scratch array, and before the next solve `trip_order_kernel` sorts every window of W = 1024 points by that
// hint (a permutation of window offsets, 2 bytes per point): lane l of group g takes the point perm[32 g + l].  The lanes
// of a warp then leave the loop together (idle lane-trips 3 %) and read the same or adjacent records (adjacent records
// never share banks: the record stride is an odd number of 16-byte chunks).  Which lane computes a point does not change
// its arithmetic: results are bit-identical with and without the ordering (tested); a stale or zero hint (first step) only
// costs the speed-up.  Loads and stores of a group are scattered over the window (8 KB per field) instead of 256
// contiguous bytes: L1/L2 absorb it.
#include <deque>
#include <mutex>

#include "ne_flux_tab2.cuh"
#include "ne_queue_host.cuh"

namespace ne {

constexpr int TAB2_WINDOW = 1024;
constexpr int TAB2_TRIP_BITS = 5, TAB2_REC_BITS = 8;                // hint = (min(trips, 31), record): 13 bits
static_assert(fm::PSI_NI <= (1 << TAB2_REC_BITS), "the record index must fit the hint");

__device__ __forceinline__ uint16_t tab2_hint(int trips, int record) {
  const int t = trips < 0 ? 0 : (trips > (1 << TAB2_TRIP_BITS) - 1 ? (1 << TAB2_TRIP_BITS) - 1 : trips);
  return (uint16_t)((t << TAB2_REC_BITS) | (record & ((1 << TAB2_REC_BITS) - 1)));
}

// Sort of every W-point window of the launch range by the previous step's hint: perm[window * W + k] = offset (within the
// window) of the point with the k-th smallest (trips, record).  One block of W/4 threads per window, bitonic network in
// shared memory on (hint << 10 | offset): stable, deterministic, no atomics; ~10 us for the 7100 windows of C4.
template <int W>
__global__ void __launch_bounds__(W / 4)
trip_order_kernel(const uint16_t* __restrict__ hint, const uint32_t n, uint16_t* __restrict__ perm) {
  constexpr int NT = W / 4;
  __shared__ uint32_t a[W];
  const int tid = threadIdx.x;
  const uint32_t t0 = blockIdx.x * (uint32_t)W;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = tid + NT * k;
    const uint32_t t = t0 + i;
    const uint32_t h = t < n ? (uint32_t)hint[t] : 0xffffu;     // beyond the launch range: last
    a[i] = (h << 10) | (uint32_t)i;
  }
  __syncthreads();
  for (int size = 2; size <= W; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int c = tid + NT * k;                                   // comparator index, W/2 per stage
        const int i = 2 * c - (c & (stride - 1));                     // lower element of the pair
        const int j = i + stride;
        const bool up = (i & size) == 0;
        const uint32_t x = a[i], y = a[j];
        if ((x > y) == up) { a[i] = y; a[j] = x; }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = tid + NT * k;
    perm[t0 + i] = (uint16_t)(a[i] & 1023u);
  }
}

// The same ordering by a STABLE counting sort on the trip count alone (5 bits), ~10 x fewer instructions than the bitonic
// network (ncu, cold: 160 us for the 7100 windows of C4 against the solve's 1190 us — an eighth of the solve phase).  Points of
// one trip class keep their memory order, and the table record of a point is a smooth function of position (L★ varies
// smoothly in space), so neighbouring lanes still read the same or adjacent records without sorting on the record.
// One block of 256 threads per window of W = 1024 points = 32 chunks of 32:
//   1. warp w ranks the elements of chunks w, w + 8, … within their chunk by key (match.any: the lanes that hold the same key)
//      and the first lane of every key class writes the class size to cnt[key][chunk] (unique writer, no atomics);
//   2. an exclusive scan over the 32 x 32 counters in key-major order (4 consecutive counters per thread);
//   3. element -> position off[key][chunk] + rank in chunk.
// Deterministic (no atomics), capturable, no global scratch.
template <int W>
__global__ void __launch_bounds__(W / 4)
trip_order_counting_kernel(const uint16_t* __restrict__ hint, const uint32_t n, uint16_t* __restrict__ perm) {
  static_assert(W == 1024, "32 chunks of 32 elements, 256 threads");
  constexpr int NT = W / 4, NCH = W / 32, NKEY = 1 << TAB2_TRIP_BITS;
  __shared__ uint16_t cnt[NKEY * NCH];          // [key][chunk], then the exclusive offsets
  __shared__ uint32_t warp_tot[NT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t t0 = blockIdx.x * (uint32_t)W;
  reinterpret_cast<uint2*>(cnt)[tid] = make_uint2(0u, 0u);      // 1024 x 2 bytes = 256 x 8 bytes
  __syncthreads();
  uint32_t key[4], rank[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = warp + (NT / 32) * k;                         // chunk
    const uint32_t t = t0 + (uint32_t)(c * 32 + lane);
    const uint32_t h = t < n ? (uint32_t)hint[t] : 0xffffu;     // beyond the launch range: last trip class, behind every valid point
    key[k] = (h >> TAB2_REC_BITS) & (NKEY - 1);
    const unsigned same = __match_any_sync(0xffffffffu, key[k]);
    rank[k] = __popc(same & ((1u << lane) - 1u));
    if (rank[k] == 0) cnt[key[k] * NCH + c] = (uint16_t)__popc(same);
  }
  __syncthreads();
  // exclusive scan of cnt[] in index order (key-major): thread `tid` owns counters 4 tid .. 4 tid + 3
  const uint2 mine = reinterpret_cast<const uint2*>(cnt)[tid];
  const uint32_t c0 = mine.x & 0xffffu, c1 = mine.x >> 16, c2 = mine.y & 0xffffu, c3 = mine.y >> 16;
  const uint32_t local = c0 + c1 + c2 + c3;
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  uint32_t base = incl - local;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) base += (w < warp) ? warp_tot[w] : 0u;
  __syncthreads();                                              // everyone has read its counters
  reinterpret_cast<uint2*>(cnt)[tid] = make_uint2(base | ((base + c0) << 16), (base + c0 + c1) | ((base + c0 + c1 + c2) << 16));
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = warp + (NT / 32) * k;
    const uint32_t pos = (uint32_t)cnt[key[k] * NCH + c] + rank[k];
    perm[t0 + pos] = (uint16_t)(c * 32 + lane);
  }
}

// Third form (shipped): counting sort on (trip count, table record / 4) — 11 bits — with a shared-memory histogram.  ncu of
// the solve under the two orderings above (C4, cold): sorted on (trips, record) 1167 us with 25 M shared-memory bank
// conflicts, sorted on trips alone 1222 us with 76 M: a ψ-table record is 11 chunks of 16 bytes, so the lanes of a quarter
// warp collide exactly when their records differ by a multiple of 8 — lanes within 8 consecutive records never do.  Classes
// of 4 consecutive records keep a group's lanes inside such a span at a fifth of the bitonic network's cost:
//   1. rank of an element within its key = what a shared-memory atomicAdd on cnt[key] returns (one atomic per key class of a
//      warp, the lanes of a class ranked by match.any);
//   2. exclusive scan over the 2048 counters (8 per thread);
//   3. element -> position off[key] + rank.
// The order inside a key class depends on the order the atomics land in: the permutation is not reproducible from run to
// run — the results are (which lane computes a point does not change its arithmetic; tested bit for bit).
template <int W>
__global__ void __launch_bounds__(W / 4)
trip_order_histogram_kernel(const uint16_t* __restrict__ hint, const uint32_t n, uint16_t* __restrict__ perm, const int descending,
                            uint32_t* __restrict__ group_counter) {
  static_assert(W == 1024, "256 threads, 4 elements and 8 counters each");
  constexpr int NT = W / 4, REC_SHIFT = 2, REC_CLASS_BITS = TAB2_REC_BITS - REC_SHIFT;
  constexpr int NKEY = 1 << (TAB2_TRIP_BITS + REC_CLASS_BITS);
  static_assert(NKEY == 8 * NT, "8 counters per thread");
  __shared__ __align__(16) uint32_t cnt[NKEY];
  __shared__ uint32_t warp_tot[NT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t t0 = blockIdx.x * (uint32_t)W;
  if (blockIdx.x == 0 && tid == 0) *group_counter = 0u;   // the solve's group queue: reset here instead of by a memset node of its own
  reinterpret_cast<uint4*>(cnt)[2 * tid] = make_uint4(0u, 0u, 0u, 0u);
  reinterpret_cast<uint4*>(cnt)[2 * tid + 1] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  uint32_t key[4], rank[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t t = t0 + (uint32_t)(tid + NT * k);
    const uint32_t h = t < n ? (uint32_t)hint[t] : 0xffffu;     // beyond the launch range: the last key
    uint32_t trips = (h >> TAB2_REC_BITS) & ((1u << TAB2_TRIP_BITS) - 1u);
    // descending: a window's longest groups are drawn first, so the groups drawn last in a launch — the ones the kernel's tail
    // waits for — are its shortest
    if (descending && t < n) trips = ((1u << TAB2_TRIP_BITS) - 1u) - trips;
    key[k] = (trips << REC_CLASS_BITS) | ((h & ((1u << TAB2_REC_BITS) - 1u)) >> REC_SHIFT);
    // one atomic per key class of the warp (lanes that hold the same key are ranked by match.any): 32 lanes on one key
    // would otherwise serialise on one shared-memory word
    const unsigned same = __match_any_sync(0xffffffffu, key[k]);
    const int leader = __ffs(same) - 1;
    uint32_t first = 0;
    if (lane == leader) first = atomicAdd(&cnt[key[k]], (uint32_t)__popc(same));
    rank[k] = __shfl_sync(0xffffffffu, first, leader) + (uint32_t)__popc(same & ((1u << lane) - 1u));
  }
  __syncthreads();
  uint4 a = reinterpret_cast<const uint4*>(cnt)[2 * tid], b = reinterpret_cast<const uint4*>(cnt)[2 * tid + 1];
  const uint32_t local = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  uint32_t base = incl - local;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) base += (w < warp) ? warp_tot[w] : 0u;
  uint4 oa, ob;
  oa.x = base; oa.y = oa.x + a.x; oa.z = oa.y + a.y; oa.w = oa.z + a.z;
  ob.x = oa.w + a.w; ob.y = ob.x + b.x; ob.z = ob.y + b.y; ob.w = ob.z + b.z;
  reinterpret_cast<uint4*>(cnt)[2 * tid] = oa;      // each thread overwrites only the counters it has read
  reinterpret_cast<uint4*>(cnt)[2 * tid + 1] = ob;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) perm[t0 + cnt[key[k]] + rank[k]] = (uint16_t)(tid + NT * k);
}

// 8 warps per CTA, 3 CTAs per SM (80 registers), persistent: a CTA stages the table once; its warps then draw groups of 32
// sorted points from a global counter until the launch range is exhausted and never synchronise.  (Round 2 first shipped
// one 1024-point window per CTA: the table was staged 7100 times per C4 launch and a CTA's fast warps idled until its
// slowest one had finished — 15 % of the warp slots, ncu `sm__warps_active` 32 % of a possible 37.5 %.)
// Shape of a CTA: NW warps, REGS registers per thread; everything in dynamic shared memory:
//   [ solver table | log table per lane class | SL_COUNT columns of NW*32 doubles | parameters of the rare paths ]
// PF (opt-in, NE_B200_TAB2_STAGING=1; measured: no gain): the inputs of a warp's NEXT group are staged by cp.async into the thread's own column of ST_COUNT
// doubles while the current group is being solved — ua, va, Ta, pa, qa, the two uo and two vo values of the cell-centre
// average, To, So, and one slot of two 32-bit words (the 4-byte words that hold the point's mask byte and the lane's entry of
// the next permutation row).  A group's prologue then reads shared memory instead of waiting one global-memory latency per
// group (ncu of the non-staged kernel: 1.7 of its 9 stall cycles per issued instruction were long-scoreboard, nearly all on
// these loads); nothing is carried in registers across the solve for it.
enum { ST_UA, ST_VA, ST_TA, ST_PA, ST_QA, ST_UO0, ST_UO1, ST_VO0, ST_VO1, ST_TO, ST_SO, ST_WORDS, ST_COUNT };
template <int NW, bool PF = false> constexpr size_t tab2_smem_bytes() {
  return sizeof(double) * (fm::TAB_SIZE + 2 * fm::LOG_N * fm::LOG_REP + (SL_COUNT + (PF ? ST_COUNT : 0)) * NW * 32) + sizeof(Tab2Rare);
}

__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ int64_t tab2_index(const Layout& L, uint32_t t) {
  const uint32_t jj = t / (uint32_t)L.ni;
  return L.at(L.i_lo + (int32_t)(t - jj * (uint32_t)L.ni), L.j_lo + (int32_t)jj);
}

// request the inputs of point `idx` into the thread's staging column (array slots only: constants are read from the descriptor)
template <int NT>
__device__ __forceinline__ void tab2_stage_issue(const NeAtmosOceanDesc& d, const Layout& L, int64_t idx, bool relative, double* st) {
  cp_async8(st + ST_UA * NT, (const double*)d.ua + idx);
  cp_async8(st + ST_VA * NT, (const double*)d.va + idx);
  cp_async8(st + ST_TA * NT, (const double*)d.Ta + idx);
  cp_async8(st + ST_PA * NT, (const double*)d.pa + idx);
  cp_async8(st + ST_QA * NT, (const double*)d.qa + idx);
  if (relative) {
    if (d.uo.ptr) { cp_async8(st + ST_UO0 * NT, (const double*)d.uo.ptr + idx); cp_async8(st + ST_UO1 * NT, (const double*)d.uo.ptr + idx + 1); }
    if (d.vo.ptr) { cp_async8(st + ST_VO0 * NT, (const double*)d.vo.ptr + idx); cp_async8(st + ST_VO1 * NT, (const double*)d.vo.ptr + idx + L.sx); }
  }
  if (d.To.ptr) cp_async8(st + ST_TO * NT, (const double*)d.To.ptr + idx);
  if (d.So.ptr) cp_async8(st + ST_SO * NT, (const double*)d.So.ptr + idx);
  if (d.inactive) {   // the aligned 4-byte word that holds the mask byte (never leaves the byte's own page)
    const uintptr_t a = (uintptr_t)(d.inactive + idx);
    cp_async4(reinterpret_cast<uint32_t*>(st + ST_WORDS * NT), reinterpret_cast<const void*>(a & ~(uintptr_t)3));
  }
}

// the staged point, exactly as tab2_load forms it
template <int NT>
__device__ __forceinline__ void tab2_stage_read(const NeAtmosOceanDesc& d, int64_t idx, bool celsius, bool relative, const double* st,
                                                Parked& k, double& uo, double& vo, double& So, bool& not_water) {
  k.du = st[ST_UA * NT]; k.dv = st[ST_VA * NT]; k.Ta = st[ST_TA * NT]; k.pa = st[ST_PA * NT]; k.qa = st[ST_QA * NT];
  uo = vo = 0;
  if (relative) {
    uo = d.uo.ptr ? (st[ST_UO0 * NT] + st[ST_UO1 * NT]) / 2 : d.uo.value;
    vo = d.vo.ptr ? (st[ST_VO0 * NT] + st[ST_VO1 * NT]) / 2 : d.vo.value;
  }
  double To = d.To.ptr ? st[ST_TO * NT] : d.To.value;
  if (celsius) To = To + 273.15;
  k.Ts = To;
  So = d.So.ptr ? st[ST_SO * NT] : d.So.value;
  not_water = false;
  if (d.inactive) {
    const uint32_t w = reinterpret_cast<const uint32_t*>(st + ST_WORDS * NT)[0];
    not_water = ((w >> (8u * (uint32_t)((uintptr_t)(d.inactive + idx) & 3))) & 0xffu) != 0;
  }
}

template <class CT, bool SORT, class O, int NW, int REGS, bool PF = false>
__global__ void __launch_bounds__(NW * 32) __maxnreg__(REGS)
ao_flux_tab2_kernel(const __grid_constant__ NeAtmosOceanDesc d, const __grid_constant__ Layout L,
                    const __grid_constant__ Thermo<CT> th, const __grid_constant__ FastParams P,
                    const __grid_constant__ TabParams T, const __grid_constant__ Micro Mi,
                    const __grid_constant__ Tab2First F, const double* __restrict__ gtab,
                    const uint16_t* __restrict__ perm, uint16_t* __restrict__ hint, unsigned long long* __restrict__ counts,
                    uint32_t* __restrict__ next_group) {
  constexpr int NT = NW * 32, W = TAB2_WINDOW;
  extern __shared__ __align__(16) double tab[];   // fm::TAB_SIZE doubles, then:
  double* const lrep = tab + fm::TAB_SIZE;         // the log table once per lane class (fm::log_pos_rep)
  double* const park = lrep + 2 * fm::LOG_N * fm::LOG_REP;
  Tab2Rare& rare = *reinterpret_cast<Tab2Rare*>(park + (SL_COUNT + (PF ? ST_COUNT : 0)) * NT);   // behind the staging columns
  if (threadIdx.x == 0) { rare.P = P; rare.T = T; }
  for (int k = threadIdx.x; k < fm::TAB_SIZE / 2; k += NT)
    reinterpret_cast<double2*>(tab)[k] = __ldg(reinterpret_cast<const double2*>(gtab) + k);
  for (int k = threadIdx.x; k < fm::LOG_N * fm::LOG_REP; k += NT)
    reinterpret_cast<double2*>(lrep)[k] = __ldg(reinterpret_cast<const double2*>(gtab + fm::TAB_LOG) + k / fm::LOG_REP);
  __syncthreads();
  O o;
  const uint32_t n = (uint32_t)L.ni * (uint32_t)L.nj;
  const uint32_t n_groups = ((n + W - 1) / W) * (W / 32);
  const bool celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  const bool relative = d.properties.velocity_formulation == NE_VEL_RELATIVE;
  const int lane = threadIdx.x & 31, tid = threadIdx.x;
  double* const col = park + tid;
  const Tab2Heights H = {d.boundary_layer_height.value, d.surface_layer_height.value - P.d_zero, T.log_hd};
  unsigned long long warp_trips = 0;
  // groups are drawn in memory order: a window's lines are fetched once.  (Every window's most expensive group first —
  // longest-processing-time-first, for the few-groups-per-warp launches of a multi-GPU band — was measured: no gain at
  // 1/8 of C4, 13 % slower at 1/4, 48 % slower on the full grid, where the windows no longer stay in L2.)
  auto grab = [&]() {
    uint32_t v = 0;
    if (lane == 0) v = atomicAdd(next_group, 1u);
    return __shfl_sync(0xffffffffu, v, 0);
  };
  // Software pipeline over a warp's groups.  Not staged (PF = false): the next group index is requested one group ahead, and so
  // is the next group's row of the permutation.  Staged (PF): group indices are requested two groups ahead, the permutation row
  // one and a half (as a cp.async word), the fields of the next group one ahead — every latency hides behind a whole solve.
  constexpr uint32_t END = 0xffffffffu;
  double* const st = park + SL_COUNT * NT + tid;                          // staging column (PF)
  uint32_t* const stw = reinterpret_cast<uint32_t*>(st + ST_WORDS * NT);  // [0] mask word, [1] permutation word
  auto point_of = [&](uint32_t gg, uint32_t p) {                          // position in the launch range of this lane's point of group gg
    const uint32_t slot = gg * 32u + (uint32_t)lane;
    return SORT ? (slot & ~(uint32_t)(W - 1)) + p : slot;
  };
  auto stage_perm = [&](uint32_t gg) {                                    // the aligned word that holds perm[gg * 32 + lane]
    if (SORT) cp_async4(stw + 1, perm + ((gg * 32u + (uint32_t)lane) & ~1u));
  };
  uint32_t g = grab();
  uint32_t g_next = g < n_groups ? grab() : g;
  uint32_t g_after = (PF && g_next < n_groups) ? grab() : g_next;
  uint32_t p_cur = (SORT && g < n_groups) ? perm[g * 32u + (uint32_t)lane] : 0u;
  uint32_t t_cur = g < n_groups ? point_of(g, p_cur) : END;
  if (PF) {
    if (t_cur < n) tab2_stage_issue<NT>(d, L, tab2_index(L, t_cur), relative, st);
    if (g_next < n_groups) stage_perm(g_next);
    cp_async_commit();
  }
#pragma unroll 1
  while (t_cur != END) {
    const uint32_t t = t_cur;
    const bool valid = t < n;
    int64_t idx = 0;
    bool not_water = true;
    Parked k;
    double So = 0;
    if (PF) {
      cp_async_wait_all();
      double uo, vo;
      if (valid) {
        idx = tab2_index(L, t);
        tab2_stage_read<NT>(d, idx, celsius, relative, st, k, uo, vo, So, not_water);
        if (!not_water) { k.du -= uo; k.dv -= vo; }
      }
      // the next group's point, its fields, the permutation row of the group after it, the index of the one after that
      t_cur = END;
      if (g_next < n_groups) {
        const uint32_t p = SORT ? ((stw[1] >> (16u * ((uint32_t)lane & 1u))) & 0xffffu) : 0u;
        t_cur = point_of(g_next, p);
        if (t_cur < n) tab2_stage_issue<NT>(d, L, tab2_index(L, t_cur), relative, st);
      }
      if (g_after < n_groups) stage_perm(g_after);
      cp_async_commit();
      g_next = g_after;
      if (g_after < n_groups) g_after = grab();
    } else {
      g = g_next;
      if (g < n_groups) g_next = grab();
      if (SORT && g < n_groups) p_cur = perm[g * 32u + (uint32_t)lane];   // consumed by the next pass
      t_cur = g < n_groups ? END - 1u : END;                              // resolved at the bottom, when p_cur has landed
      // (an L2 prefetch of the next group's ten input lines from here was measured: 1.360 vs 1.351 ms on C4; profiles/r02_notes.md)
      if (valid) {   // the mask and the fields are requested together: one memory latency per group, not two
        idx = tab2_index(L, t);
        const uint8_t mask = d.inactive ? d.inactive[idx] : (uint8_t)0;
        double uo, vo;
        tab2_load<true>(d, L, idx, celsius, relative, k, uo, vo, So);
        not_water = mask != 0;
        if (!not_water) { k.du -= uo; k.dv -= vo; }   // land: zero_interface_state, Δu = uₐ − 0 (interface_states.jl:800-803)
      }
    }
    const bool solve = valid && !(not_water && !P.fixed);   // needs_to_converge && not_water: no solve (atmosphere_ocean_fluxes.jl:144)
    const unsigned solving = __ballot_sync(0xffffffffu, solve);   // all 32 lanes are converged here (loop head)
    double ustar = 0, theta_star = 0, q_star = 0;
    int iters = 0;
    if (valid) {
      if (solve) {
        Tab2Point s;
        tab2_invariants<NT>(o, d, th, P, T, tab, k, So, s, col);
        col[SL_DU * NT] = k.du; col[SL_DV * NT] = k.dv; col[SL_TA * NT] = k.Ta; col[SL_PA * NT] = k.pa; col[SL_QA * NT] = k.qa; col[SL_TS * NT] = k.Ts;
        int record;
        iters = tab2_solve<NT>(o, P, T, rare, Mi, tab, lrep, lane & (fm::LOG_REP - 1), H, F, s, col, counts, record);
        if (hint) hint[t] = tab2_hint(iters, record);
        ustar = s.ustar;
        if (iters > 0) { theta_star = o.mul(s.chi_s, col[SL_DTH * NT]); q_star = o.mul(s.chi_s, col[SL_DQ * NT]); }
        else theta_star = q_star = 1e-4;
        k.du = col[SL_DU * NT]; k.dv = col[SL_DV * NT]; k.Ta = col[SL_TA * NT]; k.pa = col[SL_PA * NT]; k.qa = col[SL_QA * NT]; k.Ts = col[SL_TS * NT];
        if (!std::is_same<O, fm::OpsPlain>::value) {   // warp trips = the slowest lane's, once per group
          __syncwarp(solving);
          const int mx = __reduce_max_sync(solving, iters);
          if (lane == __ffs(solving) - 1) warp_trips += mx;
        }
      }
      else if (hint) hint[t] = 0;
      tab2_epilogue<O, CT>(o, d, th, idx, celsius, not_water, k, ustar, theta_star, q_star, iters);
    }
    if (!PF && t_cur != END) t_cur = point_of(g, p_cur);
  }
  if (!std::is_same<O, fm::OpsPlain>::value) {
    o.flush(counts);
    if (counts && warp_trips) atomicAdd(counts + 5, warp_trips);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------
// Scratch of the ordering, one per (device, output array, launch extent): the permutation and the hints, 2 + 2 bytes per
// point, allocated and zeroed on first use (outside CUDA-graph capture, like the solver tables) and kept for the life
// of the process.  Keyed by the friction-velocity output array and the launch range: two steps that write the same
// outputs over the same range share their hints.
struct OrderBuf { int device; const void* key; uint32_t n; int64_t i_lo, j_lo; uint16_t* perm; uint16_t* hint; };
static std::mutex g_order_mutex;
static std::deque<OrderBuf> g_order;

static const OrderBuf* order_buffer(const void* key, uint32_t n, int64_t i_lo, int64_t j_lo) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(g_order_mutex);
  for (const OrderBuf& b : g_order)
    if (b.device == dev && b.key == key && b.n == n && b.i_lo == i_lo && b.j_lo == j_lo) return &b;
  if (g_order.size() >= 64) {   // a host that keeps re-allocating its output arrays: start over
    for (OrderBuf& b : g_order) cudaFree(b.perm);
    g_order.clear();
  }
  OrderBuf b{dev, key, n, i_lo, j_lo, nullptr, nullptr};
  const size_t padded = (((size_t)n + 1023) / 1024) * 1024;
  if (cudaMalloc(&b.perm, 2 * padded * sizeof(uint16_t)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  b.hint = b.perm + padded;
  if (cudaMemset(b.perm, 0, 2 * padded * sizeof(uint16_t)) != cudaSuccess) { cudaGetLastError(); cudaFree(b.perm); return nullptr; }
  g_order.push_back(b);
  return &g_order.back();
}


// Group counters: a per-device pool used round robin; the launch zeroes its counter on its own stream right before the
// kernel (a memset node under CUDA-graph capture), so an aborted launch cannot poison a later one.
constexpr int TAB2_COUNTER_SLOTS = 1024;
static uint32_t* tab2_counter() {
  struct Pool { int device; uint32_t* dptr; unsigned next; };
  static std::mutex mutex;
  static std::deque<Pool> pools;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mutex);
  for (Pool& q : pools)
    if (q.device == dev) return q.dptr + (q.next++ % TAB2_COUNTER_SLOTS);
  Pool q = {dev, nullptr, 1};
  if (cudaMalloc(&q.dptr, sizeof(uint32_t) * TAB2_COUNTER_SLOTS) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  pools.push_back(q);
  return q.dptr;
}

static unsigned tab2_grid(uint32_t n_groups, int warps, int ctas_per_sm) {
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // one CTA per SM as soon as there is a group for each: a small launch is latency-bound (a lone warp runs a trip in ~1.5 us, six
  // warps on a scheduler already contend for issue slots), so its groups are spread over all SMs instead of packed 24 to a CTA
  // (C1, 1720 groups: 71 -> 45 us); warps that draw nothing exit at once
  (void)warps;
  return std::max(1u, std::min<uint32_t>(n_groups, (uint32_t)(sms * ctas_per_sm)));
}

bool tab2_eligible(const NeAtmosOceanDesc& d, const TabParams& TP) {
  if (env_flag("NE_B200_TAB_V1")) return false;
  if (TP.general_psi || d.surface_layer_height.ptr || d.boundary_layer_height.ptr) return false;
  const int64_t n = (int64_t)(d.grid.i_hi - d.grid.i_lo + 1) * (int64_t)(d.grid.j_hi - d.grid.j_lo + 1);
  return n > 0 && n < ((int64_t)1 << 31) - TAB2_WINDOW;
}

template <class CT, int NW, int REGS, int CTAS, bool PF = false>
static int launch_tab2_t(const NeAtmosOceanDesc& d, const Layout& L, const FastParams& P, const TabParams& TP, const Micro& Mi,
                         const Tab2First& F, const double* tab, cudaStream_t s, unsigned long long* counts, uint16_t* perm, uint16_t* hint, uint32_t n) {
  constexpr int W = TAB2_WINDOW;
  const uint32_t n_windows = (n + W - 1) / W;
  uint32_t* counter = tab2_counter();
  NE_REQUIRE(counter != nullptr, "atmosphere-ocean: could not allocate the group counter");
  const int order = env_int("NE_B200_TAB2_ORDER", 2);   // 0: bitonic (trips, record); 1: stable counting sort on trips; 2: histogram on (trips, record / 4)
  if (!perm || order == 0 || order == 1)                // the shipped ordering pass resets the counter itself
    if (cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), s); e != cudaSuccess)
      return cuda_error(e, "ne_atmosphere_ocean_fluxes(tab2: counter reset)");
  if (perm) {
    if (order == 0) trip_order_kernel<W><<<n_windows, W / 4, 0, s>>>(hint, n, perm);
    else if (order == 1) trip_order_counting_kernel<W><<<n_windows, W / 4, 0, s>>>(hint, n, perm);
    else trip_order_histogram_kernel<W><<<n_windows, W / 4, 0, s>>>(hint, n, perm, env_int("NE_B200_TAB2_DESCENDING", 0), counter);
    NE_CUDA_CHECK_LAUNCH("ne_atmosphere_ocean_fluxes(trip order)");
  }
  const unsigned grid = tab2_grid(n_windows * (W / 32), NW, CTAS);
  const Thermo<CT> th = Thermo<CT>::make(d.thermo);
  constexpr size_t smem = tab2_smem_bytes<NW, PF>();
#define NE_TAB2_GO(SORT, O, STAGED)                                                                                      \
  do {                                                                                                                   \
    if (cudaError_t e = allow_table_smem<ao_flux_tab2_kernel<CT, SORT, O, NW, REGS, STAGED>>(smem); e != cudaSuccess)    \
      return cuda_error(e, "ne_atmosphere_ocean_fluxes(tab2: shared memory opt-in)");                                    \
    ao_flux_tab2_kernel<CT, SORT, O, NW, REGS, STAGED><<<grid, NW * 32, smem, s>>>(d, L, th, P, TP, Mi, F, tab, perm, hint, counts, counter); \
  } while (0)
  if (counts) {
    if (perm) NE_TAB2_GO(true, fm::OpsCount, false); else NE_TAB2_GO(false, fm::OpsCount, false);
  } else {
    if (perm) NE_TAB2_GO(true, fm::OpsPlain, PF); else NE_TAB2_GO(false, fm::OpsPlain, PF);
  }
#undef NE_TAB2_GO
  NE_CUDA_CHECK_LAUNCH("ne_atmosphere_ocean_fluxes(tab2)");
  return NE_OK;
}

// counts != nullptr: the counting instantiation (6 device counters: fma, mul, add, other, thread trips, warp trips);
// host_tab: the host copy of the solver table (the micro records travel as a kernel parameter)
int launch_tab2(const NeAtmosOceanDesc& d, const Layout& L, const FastParams& P, const TabParams& TP, const double* tab,
                const double* host_tab, cudaStream_t s, unsigned long long* counts) {
  const uint32_t n = (uint32_t)L.ni * (uint32_t)L.nj;
  Micro Mi;
  for (int side = 0; side < 2; ++side)
    for (int k = 0; k < fm::MICRO_REC; ++k) Mi.rec[side][k] = host_tab[fm::TAB_MICRO + side * fm::MICRO_REC + k];
  uint16_t *perm = nullptr, *hint = nullptr;
  if (!env_flag("NE_B200_TAB2_NO_ORDER"))
    if (const OrderBuf* ob = order_buffer(d.friction_velocity, n, d.grid.i_lo, d.grid.j_lo)) { perm = ob->perm; hint = ob->hint; }
  const Tab2First F = make_tab2_first(P, TP, host_tab, d.surface_layer_height.value - P.d_zero, TP.log_hd);
  // CTA shape (NE_B200_TAB2_SHAPE; C4: 1.322 / 1.355 / 1.278 ms): 0 = 8 warps x 3 CTAs per SM at 80 registers, 1 = 32 warps x 1 CTA at 64
  // registers (a third more resident warps, but the spills push the L1TEX data pipe to 81 %), 2 = 24 warps x 1 CTA at 80 registers (shipped:
  // one table per SM instead of three), 3 = 28 warps x 1 CTA at 72 registers (25 warps at 80 do not launch: registers are granted per 4 warps)
  const int shape = counts ? 0 : env_int("NE_B200_TAB2_SHAPE", 2);
  const bool f64 = d.thermo.dtype == NE_F64;
#define NE_TAB2_SHAPE(NW, REGS, CTAS)                                                                            \
  return f64 ? launch_tab2_t<double, NW, REGS, CTAS>(d, L, P, TP, Mi, F, tab, s, counts, perm, hint, n)          \
             : launch_tab2_t<float, NW, REGS, CTAS>(d, L, P, TP, Mi, F, tab, s, counts, perm, hint, n)
  if (shape == 1) { NE_TAB2_SHAPE(32, 64, 1); }
  if (shape == 2) {
    // cp.async staging of the next group's inputs (PF): bit-identical, measured 1.295 vs 1.289 ms on C4, 0.212 vs 0.201 ms on C2
    // (profiles/r02_time_ao_j21_staging.log): the other warps of a scheduler already cover a group's load latency.  Opt-in.
    if (env_flag("NE_B200_TAB2_STAGING"))
      return f64 ? launch_tab2_t<double, 24, 80, 1, true>(d, L, P, TP, Mi, F, tab, s, counts, perm, hint, n)
                 : launch_tab2_t<float, 24, 80, 1, true>(d, L, P, TP, Mi, F, tab, s, counts, perm, hint, n);
    NE_TAB2_SHAPE(24, 80, 1);
  }
  if (shape == 3) { NE_TAB2_SHAPE(28, 72, 1); }   // 1.280 vs 1.288 ms on C4: ~5 local-memory accesses per trip eat the extra warps
  NE_TAB2_SHAPE(8, 80, 3);
#undef NE_TAB2_SHAPE
}

}  // namespace ne
