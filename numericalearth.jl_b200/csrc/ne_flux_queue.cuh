// ne_flux_queue.cuh — warp work-queue form of the table-driven a–o solve (default plugin tree), Float64 and
// Float32 models.
//
// Why: the fixed point of compute_interface_state.jl:31-58 needs a different number of trips at every point
// (Float64: 7–24, mean 13.6; Float32 models: 0.8 % of the points sit on a one-ulp limit cycle and run all 100
// trips, the rest ~14).  With one thread per point and the warp waiting for its slowest lane, 25 % (Float64) /
// 60 % (Float32) of the lane-trips are idle lanes (oracle iteration counts, profiles/r01_notes.md).  Here every
// warp is persistent and keeps a small ring of pending points in shared memory:
//
//   pull    32-point tiles (coalesced) → inactive points get their zero state written at once, active points are
//           appended to the ring (ballot + prefix rank), until a full round of 32 is available;
//   round   32 dense lanes: prologue (state loads, q_sat, invariants) → iterate with a convergence vote →
//           the round ends when at most THETA lanes are still iterating; those lanes push (point, trip count,
//           u★, θ★, q★) to a second ring of DEFERRED points, which is run as a round of its own whenever it
//           holds 32 (deferred lanes have few trips left: mixed with fresh points they would idle for most of
//           the round) — the iterate is carried bit for bit, so the result and the trip count of every point
//           are exactly those of the one-thread-per-point kernel (tested); converged lanes run the flux
//           epilogue and store.
//
// Tiles are handed out by a global atomic counter (self re-arming: the last CTA to leave zeroes it), so the load
// balances itself across warps whatever the land mask looks like.
// No block-level synchronisation between the table staging and the exit: warps of a CTA run independently.  Ring
// capacity: a pull adds at most 32 to at most 31 fresh points, a round at most 32 to at most 31 deferred ones:
// 64 slots per warp and ring suffice.
#pragma once

#include "ne_flux_tab.cuh"

namespace ne {

constexpr int QRING = 64;

template <class FT> struct QPointOf;
template <> struct QPointOf<double> { using type = FastPoint; };
template <> struct QPointOf<float> { using type = FastPointF; };

// flux epilogue + stores of one point (atmosphere_ocean_fluxes.jl:145-196); `not_water` applies
// zero_interface_state (interface_states.jl:800-803)
template <class FT, class CT>
__device__ __noinline__ void ao_write_outputs(const NeAtmosOceanDesc& d, const Layout& L, const Thermo<CT>& th,
                                              int64_t idx, FT ustar, FT theta_star, FT q_star, bool not_water, int iters) {
  const bool celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  const bool relative = d.properties.velocity_formulation == NE_VEL_RELATIVE;
  AtmosState<FT> a;
  a.u = __ldg((const FT*)d.ua + idx);
  a.v = __ldg((const FT*)d.va + idx);
  a.T = __ldg((const FT*)d.Ta + idx);
  a.p = __ldg((const FT*)d.pa + idx);
  a.q = __ldg((const FT*)d.qa + idx);
  a.z = 0; a.h_bl = 0;
  FT du = a.u, dv = a.v, Ts;
  if (not_water) {
    ustar = 0; theta_star = 0; q_star = 0; Ts = (FT)273.15;
  } else {
    if (relative) {
      du -= d.uo.ptr ? (slot_at<FT>(d.uo, idx) + slot_at<FT>(d.uo, idx + 1)) / 2 : (FT)d.uo.value;
      dv -= d.vo.ptr ? (slot_at<FT>(d.vo, idx) + slot_at<FT>(d.vo, idx + L.sx)) / 2 : (FT)d.vo.value;
    }
    Ts = slot_at<FT>(d.To, idx);
    if (celsius) Ts = Ts + (FT)273.15;
  }
  FluxEpilogue<FT, CT> e(th, a, ustar, theta_star, q_star, du, dv, false);
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

// iteration invariants of one point (BulkTemperature: everything but the iterate is fixed)
template <class CT, bool HS>
__device__ __forceinline__ void ao_prologue(const NeAtmosOceanDesc& d, const Layout& L, const Thermo<CT>& th,
                                            const FastParams& P, const FrontF32&, const TabParams& T, int64_t idx, FastPoint& s) {
  using FT = double;
  const bool celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  const bool relative = d.properties.velocity_formulation == NE_VEL_RELATIVE;
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
}

// Float32 model with Float32 thermodynamics (the PrescribedAtmosphere default: CT = eltype of the atmosphere data)
template <class CT, bool HS>
__device__ __forceinline__ void ao_prologue(const NeAtmosOceanDesc& d, const Layout& L, const Thermo<CT>& th,
                                            const FastParams&, const FrontF32& Q, const TabParams& T, int64_t idx, FastPointF& s) {
  using FT = float;
  static_assert(std::is_same<CT, float>::value, "Float32 fast path: Float32 thermodynamics only");
  const bool celsius = d.ocean.temperature_units == NE_DEGREES_CELSIUS;
  const bool relative = d.properties.velocity_formulation == NE_VEL_RELATIVE;
  const FT au = __ldg((const FT*)d.ua + idx), av = __ldg((const FT*)d.va + idx);
  const FT aT = __ldg((const FT*)d.Ta + idx), ap = __ldg((const FT*)d.pa + idx), aq = __ldg((const FT*)d.qa + idx);
  const FT az = HS ? (FT)d.surface_layer_height.value : slot_at<FT>(d.surface_layer_height, idx);
  FT du = au, dv = av;
  if (relative) {
    du -= d.uo.ptr ? (slot_at<FT>(d.uo, idx) + slot_at<FT>(d.uo, idx + 1)) / 2 : (FT)d.uo.value;
    dv -= d.vo.ptr ? (slot_at<FT>(d.vo, idx) + slot_at<FT>(d.vo, idx + L.sx)) / 2 : (FT)d.vo.value;
  }
  FT To = slot_at<FT>(d.To, idx);
  if (celsius) To = To + 273.15f;
  const FT qs = surface_specific_humidity<FT, CT>(d.properties, th, ap, To, slot_at<FT>(d.So, idx));
  const FT Tv = th.virtual_temperature(To, qs);
  s.gTv = Q.g / Tv;
  s.c1 = 1 + th.delta * qs;
  s.c2 = th.delta * Tv;
  s.dudv2 = __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
  s.h_bl = HS ? (FT)d.boundary_layer_height.value : slot_at<FT>(d.boundary_layer_height, idx);
  s.hd = az - Q.d_zero;
  s.log_hd = HS ? T.log_hd : log((double)s.hd);
  s.dtheta = (aT + Q.g * az / th.cp_m(aq)) - To;
  s.dq = aq - qs;
}

// ---- the problems the work-queue kernel runs ----------------------------------------------------------
// A Problem supplies: Params (one __grid_constant__ POD), FT, Point, NSTATE (iterate components carried by a
// deferred point), and the device functions used below.

// atmosphere–ocean, default tree (BulkTemperature): Float64 or Float32 model
template <class FT_, class CT, bool HS>
struct AoProblem {
  using FT = FT_;
  using Point = typename QPointOf<FT>::type;
  static constexpr int NSTATE = 3;
  struct Params {
    NeAtmosOceanDesc d;
    Layout L;
    Thermo<CT> th;
    FastParams P;
    TabParams T;
    FrontF32 Q;
  };
  __device__ static __forceinline__ const Layout& layout(const Params& p) { return p.L; }
  __device__ static __forceinline__ const FastParams& fast(const Params& p) { return p.P; }
  __device__ static __forceinline__ FT tolerance(const Params& p) {
    return std::is_same<FT, float>::value ? (FT)p.Q.tol : (FT)p.P.tol;
  }
  // true: the point iterates; false: its final state was written here
  __device__ static __forceinline__ bool admit(const Params& p, int32_t idx) {
    const bool not_water = p.d.inactive ? (p.d.inactive[idx] != 0) : false;
    const bool no_trips = p.P.fixed && p.P.maxiter <= 0;
    if ((not_water && !p.P.fixed) || no_trips) {   // untouched initial state (:131-137), zeroed when masked
      ao_write_outputs<FT, CT>(p.d, p.L, p.th, idx, (FT)1e-4, (FT)1e-4, (FT)1e-4, not_water, 0);
      return false;
    }
    return true;
  }
  __device__ static __forceinline__ void prologue(const Params& p, int32_t idx, Point& s, bool fresh) {
    ao_prologue<CT, HS>(p.d, p.L, p.th, p.P, p.Q, p.T, idx, s);
    if (fresh) s.ustar = s.theta_star = s.q_star = (FT)1e-4;
  }
  __device__ static __forceinline__ void get_state(const Point& s, FT* v) { v[0] = s.ustar; v[1] = s.theta_star; v[2] = s.q_star; }
  __device__ static __forceinline__ void set_state(Point& s, const FT* v) { s.ustar = v[0]; s.theta_star = v[1]; s.q_star = v[2]; }
  // one trip of iterate_interface_state; returns the drift |Δu★| + |Δθ★| + |Δq★|
  __device__ static __forceinline__ FT trip(const Params& p, const double* tab, Point& s, int) {
    const FT pu = s.ustar, pt = s.theta_star, pq = s.q_star;
    if constexpr (std::is_same<FT, float>::value) tab_iteration(p.P, p.Q, p.T, tab, s);
    else tab_iteration<false>(p.P, p.T, tab, s, nullptr);   // strict default tree only (dispatch: ne_flux_kernels.cu)
    return m_abs(s.ustar - pu) + m_abs(s.theta_star - pt) + m_abs(s.q_star - pq);
  }
  __device__ static __forceinline__ void finish(const Params& p, int32_t idx, const Point& s, int it) {
    const bool not_water = (p.P.fixed && p.d.inactive) ? (p.d.inactive[idx] != 0) : false;   // FixedIterations iterates masked points too (:144)
    ao_write_outputs<FT, CT>(p.d, p.L, p.th, idx, s.ustar, s.theta_star, s.q_star, not_water, it);
  }
};

template <class Problem, int NWARPS, int MINB>
__global__ void __launch_bounds__(NWARPS * 32, MINB)
flux_queue_kernel(const __grid_constant__ typename Problem::Params p, const double* __restrict__ gtab, const int theta,
                  uint32_t* __restrict__ counters) {
  using FT = typename Problem::FT;
  using Point = typename Problem::Point;
  constexpr int NS = Problem::NSTATE;
  extern __shared__ __align__(16) double tab[];       // fm::TAB_SIZE doubles (dynamic: see allow_table_smem)
  __shared__ int32_t f_idx[NWARPS][QRING];          // fresh points: only the point index
  __shared__ int32_t d_idx[NWARPS][QRING];          // deferred points: index, trips so far, iterate
  __shared__ int32_t d_it[NWARPS][QRING];
  __shared__ FT d_st[NWARPS][NS][QRING];
  for (int k = threadIdx.x; k < fm::TAB_SIZE / 2; k += NWARPS * 32)
    reinterpret_cast<double2*>(tab)[k] = __ldg(reinterpret_cast<const double2*>(gtab) + k);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t* const fidx = f_idx[warp];
  int32_t* const didx = d_idx[warp];
  int32_t* const dit = d_it[warp];
  const Layout& L = Problem::layout(p);
  const FastParams& P = Problem::fast(p);
  const uint32_t ni = (uint32_t)L.ni;
  const uint32_t n = ni * (uint32_t)L.nj;          // the host guarantees the parent array has < 2^31 elements
  const uint32_t ntiles = (n + 31u) >> 5;
  // Dynamic tile assignment: warps take the next 32-point tile from a global counter, so the moving front of
  // tiles in flight is contiguous in memory and a warp that drew masked (land, ice-free) tiles simply draws more.
  // The next tile index is requested one pull ahead; its latency hides behind the round in between.
  uint32_t* const tile_counter = counters;
  uint32_t* const cta_counter = counters + 1;
  auto grab = [&]() {
    uint32_t v = 0;
    if (lane == 0) v = atomicAdd(tile_counter, 1u);
    return __shfl_sync(0xffffffffu, v, 0);
  };
  uint32_t tile = grab();
  uint32_t next_tile = tile < ntiles ? grab() : tile;
  const unsigned below = (1u << lane) - 1u;
  int fhead = 0, fcount = 0, dhead = 0, dcount = 0;   // warp-uniform ring state
  const int maxiter = P.maxiter;
  const FT tol = P.fixed ? (FT)-1 : Problem::tolerance(p);   // drift ≥ 0 > −1: never "converged" under FixedIterations

  for (;;) {
    // ---- pull tiles until a full round (of deferred or of fresh points) is pending
    while (dcount < 32 && fcount < 32 && tile < ntiles) {
      const uint32_t t = tile * 32u + (uint32_t)lane;
      tile = next_tile;
      if (tile < ntiles) next_tile = grab();
      bool enq = false;
      int32_t idx = 0;
      if (t < n) {
        const uint32_t jj = t / ni;
        idx = (int32_t)L.at(L.i_lo + (int32_t)(t - jj * ni), L.j_lo + (int32_t)jj);
        enq = Problem::admit(p, idx);
      }
      const unsigned m = __ballot_sync(0xffffffffu, enq);
      if (enq) fidx[(fhead + fcount + __popc(m & below)) & (QRING - 1)] = idx;
      fcount += __popc(m);
    }
    __syncwarp();
    // ---- choose the round: 32 deferred points, else 32 fresh ones; once the tiles are exhausted, what is left
    // of both (deferred first)
    int nd, nf;
    if (dcount >= 32) { nd = 32; nf = 0; }
    else if (fcount >= 32) { nd = 0; nf = 32; }
    else { nd = dcount; nf = fcount < 32 - nd ? fcount : 32 - nd; }
    if (nd + nf == 0) break;
    const bool have = lane < nd + nf;
    int32_t idx = 0;
    int it = 0;
    Point s;
    if (have) {
      const bool fresh = lane >= nd;
      FT st[NS];
      if (!fresh) {
        const int slot = (dhead + lane) & (QRING - 1);
        idx = didx[slot];
        it = dit[slot];
#pragma unroll
        for (int k = 0; k < NS; ++k) st[k] = d_st[warp][k][slot];
      } else {
        idx = fidx[(fhead + lane - nd) & (QRING - 1)];
      }
      Problem::prologue(p, idx, s, fresh);
      if (!fresh) Problem::set_state(s, st);
    }
    dhead = (dhead + nd) & (QRING - 1); dcount -= nd;
    fhead = (fhead + nf) & (QRING - 1); fcount -= nf;
    __syncwarp();
    // deferring pays only while the deferred lanes can join a (nearly) full later round
    const int theta_eff = (tile < ntiles || dcount + fcount >= 32 - theta) ? theta : 0;
    bool done = !have;
    for (;;) {
      if (!done) {
        const FT drift = Problem::trip(p, tab, s, it);
        ++it;
        done = (drift < tol) || (it >= maxiter);
      }
      if (__popc(__ballot_sync(0xffffffffu, !done)) <= theta_eff) break;
    }
    const bool defer = !done;
    const unsigned dm = __ballot_sync(0xffffffffu, defer);
    if (defer) {
      const int pos = (dhead + dcount + __popc(dm & below)) & (QRING - 1);
      didx[pos] = idx;
      dit[pos] = it;
      FT st[NS];
      Problem::get_state(s, st);
#pragma unroll
      for (int k = 0; k < NS; ++k) d_st[warp][k][pos] = st[k];
    }
    dcount += __popc(dm);
    if (have && done) Problem::finish(p, idx, s, it);
    __syncwarp();
  }
  // the last CTA out re-arms the counters for the next launch that uses this slot
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(cta_counter, 1u) == gridDim.x - 1u) {
      *tile_counter = 0u;
      *cta_counter = 0u;
      __threadfence();
    }
  }
}

}  // namespace ne
