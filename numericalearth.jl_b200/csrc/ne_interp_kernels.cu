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
// Design: HBM-bound streaming kernel.  One thread per exchange point, 32x4 tiles so that a block
// touches a compact window of the (small, L2-resident: 646x326x4 B per field and time level)
// source grid; the 4 corner gathers of a warp collapse to 1-2 sectors each and are served by
// L1/L2 through the read-only path, so DRAM traffic is the 2 fractional indices in and the
// n_fields values out.  Interpolators (i⁻, i⁺, ξ) are computed once per point and shared by all
// fields and both time levels.
#include "ne_common.cuh"

namespace ne {

template <class AT> struct Interpolator { int64_t im, ip; AT xi; };

__device__ __forceinline__ double m_trunc(double x) { return trunc(x); }
__device__ __forceinline__ float m_trunc(float x) { return truncf(x); }

// Base.mod(x, one(x)) for floats: rem, then shift negatives into [0, 1)
template <class AT> __device__ __forceinline__ AT julia_mod1(AT x) {
  AT r = x - m_trunc(x);  // exact: equals fmod(x, 1)
  if (r == 0) return (AT)0;
  if (!(r > 0)) return r + (AT)1;
  return r;
}

// interpolator(fractional_idx): (unsafe_trunc(Int, f) + 1, i⁻ + Int(sign(f)), mod(f, 1))
template <class AT> __device__ __forceinline__ Interpolator<AT> interpolator(AT f) {
  Interpolator<AT> it;
  it.im = (int64_t)f + 1;
  it.ip = it.im + ((f > 0) ? 1 : ((f < 0) ? -1 : 0));
  it.xi = julia_mod1(f);
  return it;
}

struct InterpSource {
  int64_t ssx;     // source row stride
  int64_t off;     // (hx-1) + (hy-1)*ssx : 1-based source index -> parent offset
  int64_t o1, o2;  // time-slot offsets
};

template <class FT, class AT, class TT>
__global__ void __launch_bounds__(128)
interp_state_kernel(const __grid_constant__ NeInterpDesc d, const __grid_constant__ Layout L,
                    const __grid_constant__ InterpSource S) {
  // 32 x 4 tile per block
  const int32_t tiles_x = (L.ni + 31) / 32;
  const int32_t bx = blockIdx.x % tiles_x, by = blockIdx.x / tiles_x;
  const int32_t li = bx * 32 + (threadIdx.x & 31), lj = by * 4 + (threadIdx.x >> 5);
  if (li >= L.ni || lj >= L.nj) return;
  const int32_t i = L.i_lo + li, j = L.j_lo + lj;
  const int64_t idx = L.at(i, j);

  Interpolator<AT> ix = {1, 1, (AT)0}, iy = {1, 1, (AT)0};  // interpolator(nothing) = (1, 1, 0)
  if (d.frac_i) ix = interpolator<AT>(__ldg((const AT*)d.frac_i + idx));
  if (d.frac_j) iy = interpolator<AT>(__ldg((const AT*)d.frac_j + idx));
  const AT xi = ix.xi, eta = iy.xi;
  // ϕ₁, ϕ₃, ϕ₅, ϕ₇ with ζ = 0; the k⁺ terms are exact zeros and do not change the sum
  const AT w1 = (1 - xi) * (1 - eta), w3 = (1 - xi) * eta, w5 = xi * (1 - eta), w7 = xi * eta;
  const int64_t a_mm = S.off + ix.im + iy.im * S.ssx;
  const int64_t a_mp = S.off + ix.im + iy.ip * S.ssx;
  const int64_t a_pm = S.off + ix.ip + iy.im * S.ssx;
  const int64_t a_pp = S.off + ix.ip + iy.ip * S.ssx;
  const TT nt = (TT)d.time.frac;
  const bool same = d.time.same != 0;
  using W = decltype(AT() * TT());

#pragma unroll 1
  for (int f = 0; f < d.n_fields; ++f) {
    FT* out = (FT*)d.out[f];
    if (!out) continue;
    W total = 0;
    for (int s = 0; s < d.n_summands[f]; ++s) {
      const AT* data = (const AT*)d.series[f][s].data;
      W val = 0;  // `nothing` contributes the literal 0 (:143)
      if (data) {
        const AT* d1 = data + S.o1;
        AT p1 = w1 * __ldg(d1 + a_mm) + w3 * __ldg(d1 + a_mp) + w5 * __ldg(d1 + a_pm) + w7 * __ldg(d1 + a_pp);
        if (same) {
          val = (W)p1;
        } else {
          const AT* d2 = data + S.o2;
          AT p2 = w1 * __ldg(d2 + a_mm) + w3 * __ldg(d2 + a_mp) + w5 * __ldg(d2 + a_pm) + w7 * __ldg(d2 + a_pp);
          val = p2 * nt + p1 * (1 - nt);
        }
      }
      total = (s == 0) ? val : total + val;
    }
    out[idx] = (FT)total;
    if (d.potential && f == d.potential_from) ((FT*)d.potential)[idx] = (FT)total / (FT)d.ocean_reference_density;
  }
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
  InterpSource S;
  S.ssx = d.src_nx + 2 * d.src_hx;
  const int64_t plane = S.ssx * (d.src_ny + 2 * d.src_hy);
  S.off = (d.src_hx - 1) + (d.src_hy - 1) * S.ssx;
  S.o1 = (int64_t)(d.time.m1 - 1) * plane;
  S.o2 = (int64_t)(d.time.m2 - 1) * plane;
  const int64_t tiles = (int64_t)((L.ni + 31) / 32) * ((L.nj + 3) / 4);
  interp_state_kernel<FT, AT, TT><<<(unsigned)tiles, 128, 0, stream>>>(d, L, S);
  NE_CUDA_CHECK_LAUNCH("ne_interp_state");
  return NE_OK;
}

template <class FT>
static int interp_entry(const NeInterpDesc* d, void* stream) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  NE_REQUIRE(grid_ok(d->grid, 0, 0), "interp: launch range leaves the parent array");
  NE_REQUIRE(d->n_fields >= 0 && d->n_fields <= 9, "interp: n_fields out of range");
  NE_REQUIRE(d->src_nx > 0 && d->src_ny > 0 && d->src_nt > 0, "interp: bad source extents");
  NE_REQUIRE(d->time.m1 >= 1 && d->time.m1 <= d->src_nt && d->time.m2 >= 1 && d->time.m2 <= d->src_nt,
             "interp: time memory slots out of range");
  for (int f = 0; f < d->n_fields; ++f)
    NE_REQUIRE(d->n_summands[f] >= 0 && d->n_summands[f] <= NE_MAX_SUMMANDS, "interp: too many summands");
  if (d->potential) NE_REQUIRE(d->potential_from >= 0 && d->potential_from < d->n_fields, "interp: potential_from out of range");
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
int ne_interp_state_f64(const NeInterpDesc* d, void* stream) { return ne::interp_entry<double>(d, stream); }
int ne_interp_state_f32(const NeInterpDesc* d, void* stream) { return ne::interp_entry<float>(d, stream); }
int ne_frac_indices_f64(const NeFracIndexDesc* d, void* stream) { return ne::frac_entry<double>(d, stream); }
int ne_frac_indices_f32(const NeFracIndexDesc* d, void* stream) { return ne::frac_entry<float>(d, stream); }
}
