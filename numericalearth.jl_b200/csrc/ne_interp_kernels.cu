// ne_interp_kernels.cu — PrescribedAtmosphere / PrescribedRadiation interpolation onto the
// exchange grid (bilinear in space, linear in time) and the one-time fractional-index kernel.
//
// Replaces (citations relative to /root/reference/src/):
//   _interpolate_primary_atmospheric_state!  Atmospheres/interpolate_atmospheric_state.jl:91-137
//   interp_atmos_time_series                 Atmospheres/interpolate_atmospheric_state.jl:143-182
//   _interpolate_radiation_state!            Radiations/interpolate_radiation_state.jl:43-69
//   _compute_fractional_indices!             Atmospheres/prescribed_atmosphere_regridder.jl:51-71
//   (+ Oceananigans interpolator / _interpolate / FractionalIndices, third party, restated)
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: indices and weights must be bit-exact
// with the reference, which never contracts a*b+c.
//
// Design: HBM-bound streaming kernel.  One thread per exchange point, consecutive threads along x;
// the source grid is small and L2-resident (646x326x4 B per field and time level), the 4 corner
// gathers of a warp collapse to 1-2 sectors each and are served by L1/L2 through the read-only
// path, so DRAM traffic is the 2 fractional indices in and the n_fields values out.  Interpolators (i⁻, i⁺, ξ) are computed once per point and shared by all
// fields and both time levels.
#include <cstdlib>
#include <cstring>

#include "ne_interp_device.cuh"

namespace ne {

// intrinsic_vector (interpolate_atmospheric_state.jl:123-126; Oceananigans rotation operators, third party): the
// interpolated (u, v) turned into the frame of a rotated exchange grid, every operation separately rounded in the
// promoted type of (interpolated value, exchange element type)
template <class FT, class W>
__device__ __forceinline__ void store_rotated(const NeInterpDesc& d, int64_t idx, W u, W v) {
  using P = decltype(W() * FT());
  const P c = (P)__ldg((const FT*)d.rotation_cos + idx), sn = (P)__ldg((const FT*)d.rotation_sin + idx);
  const P ur = add_rn(mul_rn((P)u, c), mul_rn((P)v, sn));
  const P vr = add_rn(mul_rn(-(P)u, sn), mul_rn((P)v, c));
  ((FT*)d.out[d.rotate_u])[idx] = (FT)ur;
  ((FT*)d.out[d.rotate_v])[idx] = (FT)vr;
}
static bool rotation_requested(const NeInterpDesc& d) { return d.rotation_cos != nullptr || d.rotation_sin != nullptr; }

// One thread per exchange point, consecutive threads along x: a block writes one contiguous 2 KB run
// per field (DRAM-friendly streams; measured 2.3x faster than 32x4 tiles on B200, profiles/r01_notes.md).
// Fully unrolled over the (at most 9) fields so every pointer is a constant-bank operand;
// single-summand fields (everything except tuple-valued precipitation) skip the summand loop.
template <class FT, class AT, class TT>
__global__ void __launch_bounds__(256)
interp_state_kernel(const __grid_constant__ NeInterpDesc d, const __grid_constant__ Layout L,
                    const __grid_constant__ InterpSource S) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= (int64_t)L.ni * L.nj) return;
  const int32_t jj = (int32_t)(t / L.ni);
  const int64_t idx = L.at(L.i_lo + (int32_t)(t - (int64_t)jj * L.ni), L.j_lo + jj);
  const InterpPoint<AT> p = interp_point<AT>(d.frac_i, d.frac_j, idx, S);
  const TT nt = (TT)d.time.frac;
  const bool same = d.time.same != 0;
  using W = decltype(AT() * TT());
  const bool rotate = d.rotation_cos != nullptr;   // validated on the host: both arrays, two distinct output fields
  W vec_u = 0, vec_v = 0;
#pragma unroll
  for (int f = 0; f < 9; ++f) {
    if (f >= d.n_fields) break;
    FT* out = (FT*)d.out[f];
    if (!out) continue;
    W total;
    if (d.n_summands[f] == 1) {
      const AT* data = (const AT*)d.series[f][0].data;
      total = data ? interp_series<AT, TT>(data, p, S, nt, same) : (W)0;
    } else {
      total = interp_field<AT, TT>(d, f, p, S, nt, same);
    }
    if (rotate && f == d.rotate_u) { vec_u = total; continue; }
    if (rotate && f == d.rotate_v) { vec_v = total; continue; }
    out[idx] = (FT)total;
    if (d.potential && f == d.potential_from) ((FT*)d.potential)[idx] = div_rn((FT)total, (FT)d.ocean_reference_density);
  }
  if (rotate) store_rotated<FT, W>(d, idx, vec_u, vec_v);
}

// ---- shared-memory staged variant --------------------------------------------------------------------
// When the exchange grid is (much) finer than the source grid — 1/12°: 6.75 exchange points per JRA55
// cell, 1/48°: 27 — the 256 consecutive points of a block fall into a window of a few dozen source
// columns and two source rows.  The block finds that window (min/max of the interpolator indices),
// copies it once for every series and both time levels into shared memory with coalesced loads, and the
// per-point gathers become LDS with immediate plane offsets: 1 instruction per corner instead of 3
// (LEA + LEA.HI.X + LDG) and ~25 cycles of latency instead of an L1/L2 round trip.  A block whose window
// does not fit (periodic wrap inside the block, curvilinear exchange rows) takes the direct-gather path;
// the arithmetic is the same __*_rn sequence either way, so results are bit-identical.
// The staged kernel is specialised on the number of series NS (every field a single non-null series
// with a non-null output: 7 for the atmosphere, 2 for the radiation, 5 when precipitation is not
// requested) so the per-series code carries no run-time flags; anything else takes the direct kernel.
constexpr int STG_W = 64, STG_H = 4;

template <int NS> struct StagedPlan {
  const void* series[NS];
  void* out[NS];
  int32_t potential_series;             // series whose value also feeds the barotropic potential, −1: none
  int32_t chunks_x;                     // blocks per row
};

template <class FT, class AT, class TT, int NS>
__global__ void __launch_bounds__(256)
interp_staged_kernel(const __grid_constant__ NeInterpDesc d, const __grid_constant__ Layout L,
                     const __grid_constant__ InterpSource S, const __grid_constant__ StagedPlan<NS> P) {
  constexpr int PL = STG_H * STG_W;
  __shared__ AT win[NS * 2 * PL];          // [series][time level][STG_H][STG_W]
  __shared__ int32_t red[4][8];
  __shared__ int32_t box[4];
  const int tid = threadIdx.x;
  const int32_t lj = blockIdx.x / P.chunks_x;
  const int32_t li = (blockIdx.x - lj * P.chunks_x) * 256 + tid;
  const bool in = li < L.ni;
  const int64_t idx = L.at(L.i_lo + (in ? li : L.ni - 1), L.j_lo + lj);
  int32_t im, ip, jm, jp;
  AT xi, eta;
  {
    const FracPair<AT> fr = load_frac<AT>(d.frac_i, d.frac_j, idx);
    interpolator<AT>(d.frac_i != nullptr, fr.i, im, ip, xi);
    interpolator<AT>(d.frac_j != nullptr, fr.j, jm, jp, eta);
  }
  // window of the block: min/max over both x indices and both y indices
  {
    int32_t x0 = min(im, ip), x1 = max(im, ip), y0 = min(jm, jp), y1 = max(jm, jp);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o));
      x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
      y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
      y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = x0; red[1][tid >> 5] = x1; red[2][tid >> 5] = y0; red[3][tid >> 5] = y1; }
    __syncthreads();
    if (tid < 4) {
      int32_t v = red[tid][0];
#pragma unroll
      for (int w = 1; w < 8; ++w) v = (tid & 1) ? max(v, red[tid][w]) : min(v, red[tid][w]);
      box[tid] = v;
    }
    __syncthreads();
  }
  const int32_t x0 = box[0], y0 = box[2];
  const int32_t Wd = box[1] - x0 + 1, Hd = box[3] - y0 + 1;
  const bool staged = Wd <= STG_W && Hd <= STG_H;
  const bool same = d.time.same != 0;
  if (staged) {
    // the four 64-thread groups copy window rows 0..3 of every (series, level) plane: no index arithmetic
    const int g = tid >> 6, c = tid & 63;
    if (g < Hd && c < Wd) {
      const int64_t o = S.off + x0 + c + (int64_t)(y0 + g) * S.ssx;
#pragma unroll
      for (int sidx = 0; sidx < NS; ++sidx) {
        const AT* src = (const AT*)P.series[sidx] + o;
        win[(sidx * 2) * PL + g * STG_W + c] = __ldg(src + S.o1);
        if (!same) win[(sidx * 2 + 1) * PL + g * STG_W + c] = __ldg(src + S.o2);
      }
    }
    __syncthreads();
  }
  if (!in) return;
  const TT nt = (TT)d.time.frac;
  using W = decltype(AT() * TT());
  const AT cx = sub_rn((AT)1, xi), cy = sub_rn((AT)1, eta);
  const AT w1 = mul_rn(cx, cy), w3 = mul_rn(cx, eta), w5 = mul_rn(xi, cy), w7 = mul_rn(xi, eta);
  if (staged) {
    const AT* b_mm = win + (jm - y0) * STG_W + (im - x0);
    const AT* b_mp = win + (jp - y0) * STG_W + (im - x0);
    const AT* b_pm = win + (jm - y0) * STG_W + (ip - x0);
    const AT* b_pp = win + (jp - y0) * STG_W + (ip - x0);
    const W cnt = (W)sub_rn((TT)1, nt);
#pragma unroll
    for (int sidx = 0; sidx < NS; ++sidx) {
      constexpr int o2 = PL;
      const int o = sidx * 2 * PL;
      const AT p1 = add_rn(add_rn(add_rn(mul_rn(w1, b_mm[o]), mul_rn(w3, b_mp[o])), mul_rn(w5, b_pm[o])), mul_rn(w7, b_pp[o]));
      W val;
      if (same) val = (W)p1;
      else {
        const AT p2 = add_rn(add_rn(add_rn(mul_rn(w1, b_mm[o + o2]), mul_rn(w3, b_mp[o + o2])), mul_rn(w5, b_pm[o + o2])),
                             mul_rn(w7, b_pp[o + o2]));
        val = add_rn(mul_rn((W)p2, (W)nt), mul_rn((W)p1, cnt));
      }
      ((FT*)P.out[sidx])[idx] = (FT)val;
      if (sidx == P.potential_series) ((FT*)d.potential)[idx] = div_rn((FT)val, (FT)d.ocean_reference_density);
    }
  } else {
    InterpPoint<AT> p;
    p.w1 = w1; p.w3 = w3; p.w5 = w5; p.w7 = w7;
    p.o_mm = S.off + im + jm * S.ssx;
    p.o_mp = S.off + im + jp * S.ssx;
    p.o_pm = S.off + ip + jm * S.ssx;
    p.o_pp = S.off + ip + jp * S.ssx;
#pragma unroll
    for (int sidx = 0; sidx < NS; ++sidx) {
      const W val = interp_series<AT, TT>((const AT*)P.series[sidx], p, S, nt, same);
      ((FT*)P.out[sidx])[idx] = (FT)val;
      if (sidx == P.potential_series) ((FT*)d.potential)[idx] = div_rn((FT)val, (FT)d.ocean_reference_density);
    }
  }
}

// ---- the staged kernel over several exchange rows (end of round 2) ---------------------------------------------------
// interp_staged_kernel spends ~100 of its ~520 instructions per point on finding its window (two block reductions) and on
// copying it; neighbouring exchange rows fall into the same source rows (6.75 exchange rows per JRA55 row at 1/12 degree).  Here a
// block keeps its 256 columns and walks ROWS consecutive rows: the first row finds the window and stages the FULL 64 x 4 cells
// from its corner (clipped to the source array), every later row only checks — four comparisons per thread and one block
// vote — that its indices still fall inside what is staged, and restages when they do not.  One row's state at a time
// (unlike interp_tile_kernel): the register count of the one-row kernel.  Same __*_rn sequence per value: bit-identical.
template <class FT, class AT, class TT, int NS, int ROWS>
__global__ void __launch_bounds__(256)
interp_rows_kernel(const __grid_constant__ NeInterpDesc d, const __grid_constant__ Layout L,
                   const __grid_constant__ InterpSource S, const __grid_constant__ StagedPlan<NS> P,
                   const int32_t src_w, const int32_t src_h) {   // extents of the source parent array (halos included)
  constexpr int PL = STG_H * STG_W;
  __shared__ AT win[NS * 2 * PL];          // [series][time level][STG_H][STG_W]
  __shared__ int32_t red[4][8];
  __shared__ int32_t box[4];
  const int tid = threadIdx.x;
  const int32_t brow = blockIdx.x / P.chunks_x;
  const int32_t li = (blockIdx.x - brow * P.chunks_x) * 256 + tid;
  const bool in = li < L.ni;
  const bool same = d.time.same != 0;
  const TT nt = (TT)d.time.frac;
  using W = decltype(AT() * TT());
  const W cnt = (W)sub_rn((TT)1, nt);
  int32_t x0 = 0, y0 = 0, Wd = 0, Hd = 0;  // staged box: source columns x0 .. x0 + Wd - 1 (1-based indices + S.off), rows y0 .. y0 + Hd - 1
  bool staged = false;
  for (int r = 0; r < ROWS; ++r) {
    const int32_t lj = brow * ROWS + r;
    if (lj >= L.nj) break;                 // block-uniform
    const int64_t idx = L.at(L.i_lo + (in ? li : L.ni - 1), L.j_lo + lj);
    int32_t im, ip, jm, jp;
    AT xi, eta;
    {
      const FracPair<AT> fr = load_frac<AT>(d.frac_i, d.frac_j, idx);
      interpolator<AT>(d.frac_i != nullptr, fr.i, im, ip, xi);
      interpolator<AT>(d.frac_j != nullptr, fr.j, jm, jp, eta);
    }
    const int32_t lx = min(im, ip), hx = max(im, ip), ly = min(jm, jp), hy = max(jm, jp);
    bool inside = staged && lx >= x0 && hx < x0 + Wd && ly >= y0 && hy < y0 + Hd;
    if (r > 0) inside = __syncthreads_and(inside) != 0;      // also: every thread is done with the window of the row before
    if (!inside) {
      int32_t a0 = lx, a1 = hx, b0 = ly, b1 = hy;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a0 = min(a0, __shfl_xor_sync(0xffffffffu, a0, o));
        a1 = max(a1, __shfl_xor_sync(0xffffffffu, a1, o));
        b0 = min(b0, __shfl_xor_sync(0xffffffffu, b0, o));
        b1 = max(b1, __shfl_xor_sync(0xffffffffu, b1, o));
      }
      if ((tid & 31) == 0) { red[0][tid >> 5] = a0; red[1][tid >> 5] = a1; red[2][tid >> 5] = b0; red[3][tid >> 5] = b1; }
      __syncthreads();
      if (tid < 4) {
        int32_t v = red[tid][0];
#pragma unroll
        for (int w = 1; w < 8; ++w) v = (tid & 1) ? max(v, red[tid][w]) : min(v, red[tid][w]);
        box[tid] = v;
      }
      __syncthreads();
      x0 = box[0]; y0 = box[2];
      const int32_t need_w = box[1] - x0 + 1, need_h = box[3] - y0 + 1;
      staged = need_w <= STG_W && need_h <= STG_H;
      if (staged) {
        // everything from the corner that fits the buffer and the source array: later rows find their cells here
        const int32_t col0 = S.off % S.ssx + x0, row0 = S.off / S.ssx + y0;       // position of (x0, y0) in the parent array
        Wd = min((int32_t)STG_W, src_w - col0);
        Hd = min((int32_t)STG_H, src_h - row0);
        const int g = tid >> 6, c = tid & 63;
        if (g < Hd && c < Wd) {
          const int64_t o = S.off + x0 + c + (int64_t)(y0 + g) * S.ssx;
#pragma unroll
          for (int sidx = 0; sidx < NS; ++sidx) {
            const AT* src = (const AT*)P.series[sidx] + o;
            win[(sidx * 2) * PL + g * STG_W + c] = __ldg(src + S.o1);
            if (!same) win[(sidx * 2 + 1) * PL + g * STG_W + c] = __ldg(src + S.o2);
          }
        }
        __syncthreads();
      } else {
        Wd = 0; Hd = 0;
      }
    }
    if (!in) continue;
    const AT cx = sub_rn((AT)1, xi), cy = sub_rn((AT)1, eta);
    const AT w1 = mul_rn(cx, cy), w3 = mul_rn(cx, eta), w5 = mul_rn(xi, cy), w7 = mul_rn(xi, eta);
    if (staged) {
      const AT* b_mm = win + (jm - y0) * STG_W + (im - x0);
      const AT* b_mp = win + (jp - y0) * STG_W + (im - x0);
      const AT* b_pm = win + (jm - y0) * STG_W + (ip - x0);
      const AT* b_pp = win + (jp - y0) * STG_W + (ip - x0);
#pragma unroll
      for (int sidx = 0; sidx < NS; ++sidx) {
        constexpr int o2 = PL;
        const int o = sidx * 2 * PL;
        const AT p1 = add_rn(add_rn(add_rn(mul_rn(w1, b_mm[o]), mul_rn(w3, b_mp[o])), mul_rn(w5, b_pm[o])), mul_rn(w7, b_pp[o]));
        W val;
        if (same) val = (W)p1;
        else {
          const AT p2 = add_rn(add_rn(add_rn(mul_rn(w1, b_mm[o + o2]), mul_rn(w3, b_mp[o + o2])), mul_rn(w5, b_pm[o + o2])),
                               mul_rn(w7, b_pp[o + o2]));
          val = add_rn(mul_rn((W)p2, (W)nt), mul_rn((W)p1, cnt));
        }
        ((FT*)P.out[sidx])[idx] = (FT)val;
        if (sidx == P.potential_series) ((FT*)d.potential)[idx] = div_rn((FT)val, (FT)d.ocean_reference_density);
      }
    } else {
      InterpPoint<AT> p;
      p.w1 = w1; p.w3 = w3; p.w5 = w5; p.w7 = w7;
      p.o_mm = S.off + im + jm * S.ssx;
      p.o_mp = S.off + im + jp * S.ssx;
      p.o_pm = S.off + ip + jm * S.ssx;
      p.o_pp = S.off + ip + jp * S.ssx;
#pragma unroll
      for (int sidx = 0; sidx < NS; ++sidx) {
        const W val = interp_series<AT, TT>((const AT*)P.series[sidx], p, S, nt, same);
        ((FT*)P.out[sidx])[idx] = (FT)val;
        if (sidx == P.potential_series) ((FT*)d.potential)[idx] = div_rn((FT)val, (FT)d.ocean_reference_density);
      }
    }
  }
}

// Host: can this descriptor take the staged kernel with NS series?
template <int NS>
static bool make_staged_plan(const NeInterpDesc& d, const Layout& L, StagedPlan<NS>& P) {
  const char* off = std::getenv("NE_B200_INTERP_DIRECT");
  if (off && off[0] == '1') return false;
  if (rotation_requested(d)) return false;   // rotated (curvilinear) exchange grids take the direct-gather kernel
  // staging pays when a 256-point block spans few source columns: exchange grid at least 4x finer than the source
  if (d.grid.nx < 4 * d.src_nx) return false;
  std::memset(&P, 0, sizeof(P));
  P.potential_series = -1;
  int ns = 0;
  for (int f = 0; f < d.n_fields; ++f) {
    if (!d.out[f]) continue;
    if (d.n_summands[f] != 1 || !d.series[f][0].data || ns >= NS) return false;
    P.series[ns] = d.series[f][0].data;
    P.out[ns] = d.out[f];
    if (d.potential && f == d.potential_from) P.potential_series = ns;
    ++ns;
  }
  if (ns != NS) return false;
  if (d.potential && P.potential_series < 0) return false;
  P.chunks_x = (L.ni + 255) / 256;
  return true;
}

template <class FT, class AT, class TT, int NS>
static bool try_staged(const NeInterpDesc& d, const Layout& L, const InterpSource& S, cudaStream_t stream) {
  StagedPlan<NS> P;
  if (!make_staged_plan<NS>(d, L, P)) return false;
  // a block walks 4 exchange rows with one staged window (interp_rows_kernel; C4, 7 series: 0.149 -> 0.139 ms, 2 / 8 rows: 0.151 /
  // 0.142 ms; bit-identical); NE_B200_INTERP_ROWS=1: one row per block (interp_staged_kernel)
  const char* rows_env = std::getenv("NE_B200_INTERP_ROWS");
  const int rows = rows_env ? std::atoi(rows_env) : 4;
  if (rows == 2 || rows == 4 || rows == 8) {
    const int32_t src_w = (int32_t)(d.src_nx + 2 * d.src_hx), src_h = (int32_t)(d.src_ny + 2 * d.src_hy);
    const int64_t brows = (L.nj + rows - 1) / rows;
    const unsigned grid = (unsigned)((int64_t)P.chunks_x * brows);
    if (rows == 2) interp_rows_kernel<FT, AT, TT, NS, 2><<<grid, 256, 0, stream>>>(d, L, S, P, src_w, src_h);
    else if (rows == 4) interp_rows_kernel<FT, AT, TT, NS, 4><<<grid, 256, 0, stream>>>(d, L, S, P, src_w, src_h);
    else interp_rows_kernel<FT, AT, TT, NS, 8><<<grid, 256, 0, stream>>>(d, L, S, P, src_w, src_h);
    return true;
  }
  interp_staged_kernel<FT, AT, TT, NS><<<(unsigned)((int64_t)P.chunks_x * L.nj), 256, 0, stream>>>(d, L, S, P);
  return true;
}

// ---- tiled variant (round 2) ----------------------------------------------------------------------------
// ncu on the staged kernel above (7 series: 0.149 ms = 48 % of the copy peak; 525 instructions per warp, issue slots 74 %
// busy, DRAM 34 %): it is instruction-issue bound — 8 LDS.32 + 14 separately rounded Float32 operations + a Float64 time
// blend per value, and per 256 outputs one window search (two block reductions) and one staging pass.  This kernel keeps
// the staging idea and changes its granularity:
//   * a block owns a tile of 256 columns x TR rows of the exchange grid: the rows of a tile fall into the same two or three
//     source rows, so ONE window search and ONE staging pass serve TR x 256 outputs;
//   * the window is stored cell-major: all series and both time levels of a source cell are contiguous (16-byte chunks
//     of {series a level 1, a level 2, series b level 1, b level 2} for Float32 data), cells padded to an odd number of
//     chunks (conflict-free for the ~6 neighbouring cells a warp touches): a corner of two series and both levels is one
//     LDS.128 — 20 shared-memory loads per point instead of 72 for 9 series.
// Same __*_rn sequence per value as the other two kernels: bit-identical results (tested).  Measured on C4 (B200, ms per
// launch; staged / tiles of 1 / 2 / 4 rows): 7 series 0.149 / 0.177 / 0.155 / 0.160, merged 9 series (step time) 1.684 /
// 1.734 / 1.702 / 1.729, 2 series 0.101 / 0.099 / 0.080 / 0.087.  The per-row state of a tile (interpolators of TR points)
// costs registers — 48 / 64 against 40, i.e. 5 / 4 resident blocks per SM against 8 — and that outweighs the instructions
// saved wherever there are many series; the tiled kernel therefore only takes launches of at most two series (a radiation
// component on its own source grid), two rows per tile.
// (cp.async.bulk.tensor for the staging copy: the row pitch of the source planes is (Nx + 2 Hx) x 4 B = 2584 B for JRA55, not
// a multiple of 16 B, so cuTensorMapEncodeTiled rejects the array as the host model lays it out; 1-D cp.async.bulk with
// 16-byte over-fetch would replace ~150 of the ~4800 warp-instructions a block issues: not where the time is.)
constexpr int TILE_ROWS = 2;

template <class AT> struct TileVec;
template <> struct TileVec<float> { using type = float4; static constexpr int SERIES = 2; };
template <> struct TileVec<double> { using type = double2; static constexpr int SERIES = 1; };
template <class AT> __device__ __forceinline__ AT tile_get(const typename TileVec<AT>::type& v, int k);
template <> __device__ __forceinline__ float tile_get<float>(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
template <> __device__ __forceinline__ double tile_get<double>(const double2& v, int k) { return k == 0 ? v.x : v.y; }

template <class AT, int NS> struct TileGeom {
  static constexpr int SPC = TileVec<AT>::SERIES;                 // series per 16-byte chunk
  static constexpr int CHUNKS = (NS + SPC - 1) / SPC;
  static constexpr int STRIDE = CHUNKS | 1;                       // chunks per cell, odd
  static constexpr int CELLS = STG_W * STG_H;
  static constexpr size_t BYTES = (size_t)CELLS * STRIDE * 16;
};

template <class FT, class AT, class TT, int NS, int TR>
__global__ void __launch_bounds__(256, 5)
interp_tile_kernel(const __grid_constant__ NeInterpDesc d, const __grid_constant__ Layout L,
                   const __grid_constant__ InterpSource S, const __grid_constant__ StagedPlan<NS> P) {
  using G = TileGeom<AT, NS>;
  using V = typename TileVec<AT>::type;
  __shared__ __align__(16) V win[G::CELLS * G::STRIDE];
  __shared__ int32_t red[4][8];
  __shared__ int32_t box[4];
  const int tid = threadIdx.x;
  const int32_t tj = blockIdx.x / P.chunks_x;
  const int32_t li = (blockIdx.x - tj * P.chunks_x) * 256 + tid;
  const int32_t lj0 = tj * TR;
  const bool in_x = li < L.ni;
  int32_t im[TR], ip[TR], jm[TR], jp[TR];
  AT xi[TR], eta[TR];
  int64_t idx[TR];
  int32_t x0 = INT32_MAX, x1 = INT32_MIN, y0 = INT32_MAX, y1 = INT32_MIN;
#pragma unroll
  for (int r = 0; r < TR; ++r) {
    const int32_t lj = min(lj0 + r, L.nj - 1);
    idx[r] = L.at(L.i_lo + (in_x ? li : L.ni - 1), L.j_lo + lj);
    const FracPair<AT> fr = load_frac<AT>(d.frac_i, d.frac_j, idx[r]);
    interpolator<AT>(d.frac_i != nullptr, fr.i, im[r], ip[r], xi[r]);
    interpolator<AT>(d.frac_j != nullptr, fr.j, jm[r], jp[r], eta[r]);
    x0 = min(x0, min(im[r], ip[r])); x1 = max(x1, max(im[r], ip[r]));
    y0 = min(y0, min(jm[r], jp[r])); y1 = max(y1, max(jm[r], jp[r]));
  }
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o));
      x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
      y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
      y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = x0; red[1][tid >> 5] = x1; red[2][tid >> 5] = y0; red[3][tid >> 5] = y1; }
    __syncthreads();
    if (tid < 4) {
      int32_t v = red[tid][0];
#pragma unroll
      for (int w = 1; w < 8; ++w) v = (tid & 1) ? max(v, red[tid][w]) : min(v, red[tid][w]);
      box[tid] = v;
    }
    __syncthreads();
  }
  x0 = box[0]; y0 = box[2];
  const int32_t Wd = box[1] - x0 + 1, Hd = box[3] - y0 + 1;
  const bool staged = Wd <= STG_W && Hd <= STG_H;
  const bool same = d.time.same != 0;
  if (staged) {
    // thread (g, c) owns window cell (row g, column c): coalesced loads per (series, level) plane, then the cell's chunks
    const int g = tid >> 6, c = tid & 63;
    if (g < Hd && c < Wd) {
      const int64_t o = S.off + x0 + c + (int64_t)(y0 + g) * S.ssx;
      V* cell = win + (g * STG_W + c) * G::STRIDE;
#pragma unroll
      for (int k = 0; k < G::CHUNKS; ++k) {   // a chunk at a time: four loads in flight, no register array of 2 NS values
        V v;
        const AT* sa = (const AT*)P.series[k * G::SPC] + o;
        v.x = __ldg(sa + S.o1);
        v.y = same ? v.x : __ldg(sa + S.o2);
        if constexpr (G::SPC == 2) {
          if (2 * k + 1 < NS) {
            const AT* sb = (const AT*)P.series[2 * k + 1 < NS ? 2 * k + 1 : 0] + o;
            v.z = __ldg(sb + S.o1);
            v.w = same ? v.z : __ldg(sb + S.o2);
          } else {
            v.z = v.w = (AT)0;
          }
        }
        cell[k] = v;
      }
    }
    __syncthreads();
  }
  if (!in_x) return;
  const TT nt = (TT)d.time.frac;
  using W = decltype(AT() * TT());
  const W cnt = (W)sub_rn((TT)1, nt);
#pragma unroll
  for (int r = 0; r < TR; ++r) {
    if (lj0 + r >= L.nj) break;
    const AT cx = sub_rn((AT)1, xi[r]), cy = sub_rn((AT)1, eta[r]);
    const AT w1 = mul_rn(cx, cy), w3 = mul_rn(cx, eta[r]), w5 = mul_rn(xi[r], cy), w7 = mul_rn(xi[r], eta[r]);
    if (staged) {
      const V* c_mm = win + ((jm[r] - y0) * STG_W + (im[r] - x0)) * G::STRIDE;
      const V* c_mp = win + ((jp[r] - y0) * STG_W + (im[r] - x0)) * G::STRIDE;
      const V* c_pm = win + ((jm[r] - y0) * STG_W + (ip[r] - x0)) * G::STRIDE;
      const V* c_pp = win + ((jp[r] - y0) * STG_W + (ip[r] - x0)) * G::STRIDE;
#pragma unroll
      for (int k = 0; k < G::CHUNKS; ++k) {
        const V mm = c_mm[k], mp = c_mp[k], pm = c_pm[k], pp = c_pp[k];
#pragma unroll
        for (int q = 0; q < G::SPC; ++q) {
          const int sidx = k * G::SPC + q;
          if (sidx >= NS) break;
          const AT p1 = add_rn(add_rn(add_rn(mul_rn(w1, tile_get<AT>(mm, 2 * q)), mul_rn(w3, tile_get<AT>(mp, 2 * q))),
                                      mul_rn(w5, tile_get<AT>(pm, 2 * q))), mul_rn(w7, tile_get<AT>(pp, 2 * q)));
          W val;
          if (same) val = (W)p1;
          else {
            const AT p2 = add_rn(add_rn(add_rn(mul_rn(w1, tile_get<AT>(mm, 2 * q + 1)), mul_rn(w3, tile_get<AT>(mp, 2 * q + 1))),
                                        mul_rn(w5, tile_get<AT>(pm, 2 * q + 1))), mul_rn(w7, tile_get<AT>(pp, 2 * q + 1)));
            val = add_rn(mul_rn((W)p2, (W)nt), mul_rn((W)p1, cnt));
          }
          ((FT*)P.out[sidx])[idx[r]] = (FT)val;
          if (sidx == P.potential_series) ((FT*)d.potential)[idx[r]] = div_rn((FT)val, (FT)d.ocean_reference_density);
        }
      }
    } else {
      InterpPoint<AT> p;
      p.w1 = w1; p.w3 = w3; p.w5 = w5; p.w7 = w7;
      p.o_mm = S.off + im[r] + jm[r] * S.ssx;
      p.o_mp = S.off + im[r] + jp[r] * S.ssx;
      p.o_pm = S.off + ip[r] + jm[r] * S.ssx;
      p.o_pp = S.off + ip[r] + jp[r] * S.ssx;
#pragma unroll
      for (int sidx = 0; sidx < NS; ++sidx) {
        const W val = interp_series<AT, TT>((const AT*)P.series[sidx], p, S, nt, same);
        ((FT*)P.out[sidx])[idx[r]] = (FT)val;
        if (sidx == P.potential_series) ((FT*)d.potential)[idx[r]] = div_rn((FT)val, (FT)d.ocean_reference_density);
      }
    }
  }
}

template <class FT, class AT, class TT, int NS>
static bool try_tiled(const NeInterpDesc& d, const Layout& L, const InterpSource& S, cudaStream_t stream) {
  const char* v1 = std::getenv("NE_B200_INTERP_STAGED_V1");
  if (v1 && v1[0] == '1') return false;
  StagedPlan<NS> P;
  if (!make_staged_plan<NS>(d, L, P)) return false;
  if (NS > 2) return false;                                       // see the measurements above
  // the tile's rows must share source rows: exchange grid at least TILE_ROWS times finer in y than the source
  if (d.grid.ny < TILE_ROWS * d.src_ny) return false;
  const int64_t tiles_y = (L.nj + TILE_ROWS - 1) / TILE_ROWS;
  interp_tile_kernel<FT, AT, TT, NS, TILE_ROWS><<<(unsigned)((int64_t)P.chunks_x * tiles_y), 256, 0, stream>>>(d, L, S, P);
  return true;
}

// ---- fractional indices ---------------------------------------------------------------------------
template <class T> __device__ __forceinline__ T m_fmod(T a, T b);
template <> __device__ __forceinline__ double m_fmod<double>(double a, double b) { return fmod(a, b); }
template <> __device__ __forceinline__ float m_fmod<float>(float a, float b) { return fmodf(a, b); }

template <class AT>
__device__ __forceinline__ AT fractional_index_search(AT x, const AT* xs, int64_t N) {  // 1-based
  int64_t low = 0, high = N - 1;
  while (low + 1 < high) {
    int64_t mid = (low + high) >> 1;
    AT v = __ldg(xs + mid);
    if (v == x) return (AT)(mid + 1);
    else if (v < x) low = mid;
    else high = mid;
  }
  int64_t i1, i2;
  if (__ldg(xs + high) == x) { i1 = i2 = high + 1; }
  else if (__ldg(xs + low) == x) { i1 = i2 = low + 1; }
  else { i1 = low + 1; i2 = high + 1; }
  if (i1 == i2) return (AT)i1;
  AT x1 = __ldg(xs + i1 - 1), x2 = __ldg(xs + i2 - 1);
  return (AT)(i2 - i1) / (x2 - x1) * (x - x1) + (AT)i1;
}

template <class FT, class AT>
__global__ void __launch_bounds__(128)
frac_indices_kernel(const __grid_constant__ NeFracIndexDesc d, const __grid_constant__ Layout L) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)L.ni * L.nj) return;
  const int32_t jj = (int32_t)(t / L.ni);
  const int32_t i = L.i_lo + (int32_t)(t - (int64_t)jj * L.ni), j = L.j_lo + jj;
  const int64_t idx = L.at(i, j);
  const AT* lam_n = (const AT*)d.src_lam_nodes;
  const AT* phi_n = (const AT*)d.src_phi_nodes;
  const AT lam0 = __ldg(lam_n), dlam = __ldg(lam_n + 1) - lam0;
  const AT phi0 = __ldg(phi_n), dphi = __ldg(phi_n + 1) - phi0;
  const int64_t hx = (L.off % L.sx) + 1;
  FT lam = d.nodes_2d ? __ldg((const FT*)d.lam + idx) : __ldg((const FT*)d.lam + (i + hx - 1));
  FT phi = d.nodes_2d ? __ldg((const FT*)d.phi + idx) : __ldg((const FT*)d.phi + (j + L.hy - 1));
  using W = decltype(FT() + AT());
  const W base = (W)(lam0 - dlam / 2);
  // convert_to_λ₀_λ₀_plus360(x, λ₀) = ((x - λ₀) % 360 + 360) % 360 + λ₀
  W lc = m_fmod<W>(m_fmod<W>((W)lam - base, (W)360) + (W)360, (W)360) + base;
  AT fi, fj;
  if (d.src_x_regular) fi = (AT)((lc - lam0) / dlam);
  else fi = fractional_index_search<AT>((AT)lc, lam_n, d.src_nx) - 1;
  if (d.src_y_regular) fj = (AT)(((W)phi - phi0) / dphi);
  else fj = fractional_index_search<AT>((AT)phi, phi_n, d.src_ny) - 1;
  if (d.frac_i) ((AT*)d.frac_i)[idx] = fi;
  if (d.frac_j) ((AT*)d.frac_j)[idx] = fj;
}

template <class FT, class AT, class TT>
static int launch_interp(const NeInterpDesc& d, cudaStream_t stream) {
  Layout L = make_layout(d.grid);
  InterpSource S = make_interp_source(d);
  {
    int active = 0;
    for (int f = 0; f < d.n_fields; ++f) active += d.out[f] != nullptr;
    bool done = false;
    if (active == 9) done = try_staged<FT, AT, TT, 9>(d, L, S, stream);   // atmosphere + radiation merged (ne_fused.cu)
    else if (active == 7) done = try_staged<FT, AT, TT, 7>(d, L, S, stream);
    else if (active == 5) done = try_staged<FT, AT, TT, 5>(d, L, S, stream);
    else if (active == 2) done = try_tiled<FT, AT, TT, 2>(d, L, S, stream) || try_staged<FT, AT, TT, 2>(d, L, S, stream);
    if (done) {
      NE_CUDA_CHECK_LAUNCH("ne_interp_state(staged)");
      return NE_OK;
    }
  }
  const int64_t n = (int64_t)L.ni * L.nj;
  interp_state_kernel<FT, AT, TT><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d, L, S);
  NE_CUDA_CHECK_LAUNCH("ne_interp_state");
  return NE_OK;
}

template <class FT>
static int interp_entry(const NeInterpDesc* d, void* stream) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  NE_REQUIRE(grid_ok(d->grid, 0, 0), "interp: launch range leaves the parent array");
  NE_REQUIRE(d->n_fields >= 0 && d->n_fields <= 9, "interp: n_fields out of range");
  NE_REQUIRE(d->src_nx > 0 && d->src_ny > 0 && d->src_nt > 0, "interp: bad source extents");
  NE_REQUIRE((d->src_nx + 2 * d->src_hx) * (d->src_ny + 2 * d->src_hy) < (int64_t)1 << 31, "interp: source plane exceeds 2^31 elements");
  NE_REQUIRE(d->time.m1 >= 1 && d->time.m1 <= d->src_nt && d->time.m2 >= 1 && d->time.m2 <= d->src_nt,
             "interp: time memory slots out of range");
  for (int f = 0; f < d->n_fields; ++f)
    NE_REQUIRE(d->n_summands[f] >= 0 && d->n_summands[f] <= NE_MAX_SUMMANDS, "interp: too many summands");
  if (d->potential) NE_REQUIRE(d->potential_from >= 0 && d->potential_from < d->n_fields, "interp: potential_from out of range");
  if (rotation_requested(*d)) {
    NE_REQUIRE(d->rotation_cos && d->rotation_sin, "interp: rotation needs both the cos and the sin array");
    NE_REQUIRE(d->rotate_u >= 0 && d->rotate_u < d->n_fields && d->rotate_v >= 0 && d->rotate_v < d->n_fields &&
               d->rotate_u != d->rotate_v && d->out[d->rotate_u] && d->out[d->rotate_v] &&
               !(d->potential && (d->potential_from == d->rotate_u || d->potential_from == d->rotate_v)),
               "interp: rotate_u / rotate_v must name two distinct stored fields");
  }
  cudaStream_t s = (cudaStream_t)stream;
  const bool a64 = d->src_dtype == NE_F64, t64 = d->time.frac_dtype == NE_F64;
  if (a64) return t64 ? launch_interp<FT, double, double>(*d, s) : launch_interp<FT, double, float>(*d, s);
  return t64 ? launch_interp<FT, float, double>(*d, s) : launch_interp<FT, float, float>(*d, s);
}

template <class FT>
static int frac_entry(const NeFracIndexDesc* d, void* stream) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  NE_REQUIRE(grid_ok(d->grid, 0, 0), "frac_indices: launch range leaves the parent array");
  NE_REQUIRE(d->lam && d->phi && d->src_lam_nodes && d->src_phi_nodes, "frac_indices: null node array");
  NE_REQUIRE(d->src_nx >= 2 && d->src_ny >= 2, "frac_indices: source grid too small");
  Layout L = make_layout(d->grid);
  const int64_t n = (int64_t)L.ni * L.nj;
  const int64_t blocks = (n + 127) / 128;
  if (d->src_dtype == NE_F64) frac_indices_kernel<FT, double><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(*d, L);
  else frac_indices_kernel<FT, float><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(*d, L);
  NE_CUDA_CHECK_LAUNCH("ne_frac_indices");
  return NE_OK;
}

}  // namespace ne

extern "C" {
int ne_interp_state_f64(const NeInterpDesc* d, void* stream) { NE_NVTX(); return ne::interp_entry<double>(d, stream); }
int ne_interp_state_f32(const NeInterpDesc* d, void* stream) { NE_NVTX(); return ne::interp_entry<float>(d, stream); }
int ne_frac_indices_f64(const NeFracIndexDesc* d, void* stream) { NE_NVTX(); return ne::frac_entry<double>(d, stream); }
int ne_frac_indices_f32(const NeFracIndexDesc* d, void* stream) { NE_NVTX(); return ne::frac_entry<float>(d, stream); }
}
