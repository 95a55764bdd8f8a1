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
#include "ne_interp_device.cuh"

namespace ne {

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
    out[idx] = (FT)total;
    if (d.potential && f == d.potential_from) ((FT*)d.potential)[idx] = div_rn((FT)total, (FT)d.ocean_reference_density);
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
  InterpSource S = make_interp_source(d);
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
