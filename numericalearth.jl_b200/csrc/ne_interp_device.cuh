// ne_interp_device.cuh — device helpers of the PrescribedAtmosphere / PrescribedRadiation
// interpolation (bilinear in space, linear in time), shared by the stand-alone interpolation kernel
// (ne_interp_kernels.cu) and the fused interpolation + solve kernel (ne_flux_kernels.cu).
//
// Restates (citations relative to /root/reference/src/):
//   _interpolate_primary_atmospheric_state!  Atmospheres/interpolate_atmospheric_state.jl:91-137
//   interp_atmos_time_series                 Atmospheres/interpolate_atmospheric_state.jl:143-182
//   (+ Oceananigans interpolator / _interpolate / FractionalIndices, third party)
//
// Bit-exactness: every product and sum goes through __*_rn intrinsics, which nvcc never contracts
// into an FMA — the reference evaluates ϕ₁d₁ + ϕ₃d₃ + ϕ₅d₅ + ϕ₇d₇ and ψ₂ñ + ψ₁(1 − ñ) with separately
// rounded operations — so these helpers are safe in translation units compiled with -fmad=true.
#pragma once

#include "ne_common.cuh"

namespace ne {

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ double trunc_(double x) { return trunc(x); }
__device__ __forceinline__ float trunc_(float x) { return truncf(x); }

// source-array geometry shared by every series of one descriptor
struct InterpSource {
  int32_t ssx;     // source row stride
  int32_t off;     // (hx-1) + (hy-1)*ssx : 1-based source index -> parent offset
  int64_t o1, o2;  // time-slot offsets (elements)
};

inline InterpSource make_interp_source(const NeInterpDesc& d) {
  InterpSource S;
  S.ssx = (int32_t)(d.src_nx + 2 * d.src_hx);
  const int64_t plane = (int64_t)S.ssx * (d.src_ny + 2 * d.src_hy);
  S.off = (int32_t)((d.src_hx - 1) + (d.src_hy - 1) * S.ssx);
  S.o1 = (int64_t)(d.time.m1 - 1) * plane;
  S.o2 = (int64_t)(d.time.m2 - 1) * plane;
  return S;
}

// per-point interpolator: the four corner offsets and weights, computed once and shared by all
// fields and both time levels
template <class AT> struct InterpPoint {
  int32_t o_mm, o_mp, o_pm, o_pp;
  AT w1, w3, w5, w7;
};

// interpolator(fractional_idx) = (unsafe_trunc(Int, f) + 1, i⁻ + Int(sign(f)), mod(f, 1));
// interpolator(nothing) = (1, 1, 0).  `f` was loaded by the caller (so it can be prefetched).
template <class AT>
__device__ __forceinline__ void interpolator(bool present, AT f, int32_t& im, int32_t& ip, AT& xi) {
  if (!present) { im = 1; ip = 1; xi = (AT)0; return; }
  im = (int32_t)f + 1;
  ip = im + ((f > 0) ? 1 : ((f < 0) ? -1 : 0));
  AT r = sub_rn(f, trunc_(f));          // exact: rem(f, 1)
  if (r == 0) r = (AT)0;                // Base.mod(x, one(x)): +0 for exact integers
  else if (!(r > 0)) r = add_rn(r, (AT)1);
  xi = r;
}

template <class AT> struct FracPair { AT i, j; };

template <class AT>
__device__ __forceinline__ FracPair<AT> load_frac(const void* frac_i, const void* frac_j, int64_t idx) {
  FracPair<AT> f;
  f.i = frac_i ? __ldg((const AT*)frac_i + idx) : (AT)0;
  f.j = frac_j ? __ldg((const AT*)frac_j + idx) : (AT)0;
  return f;
}

template <class AT>
__device__ __forceinline__ InterpPoint<AT> interp_point(bool has_i, bool has_j, FracPair<AT> fr, const InterpSource& S) {
  int32_t im, ip, jm, jp;
  AT xi, eta;
  interpolator<AT>(has_i, fr.i, im, ip, xi);
  interpolator<AT>(has_j, fr.j, jm, jp, eta);
  InterpPoint<AT> p;
  // ϕ₁, ϕ₃, ϕ₅, ϕ₇ with ζ = 0; the k⁺ terms are exact zeros and do not change the sum
  const AT cx = sub_rn((AT)1, xi), cy = sub_rn((AT)1, eta);
  p.w1 = mul_rn(cx, cy);
  p.w3 = mul_rn(cx, eta);
  p.w5 = mul_rn(xi, cy);
  p.w7 = mul_rn(xi, eta);
  p.o_mm = S.off + im + jm * S.ssx;
  p.o_mp = S.off + im + jp * S.ssx;
  p.o_pm = S.off + ip + jm * S.ssx;
  p.o_pp = S.off + ip + jp * S.ssx;
  return p;
}
template <class AT>
__device__ __forceinline__ InterpPoint<AT> interp_point(const void* frac_i, const void* frac_j, int64_t idx,
                                                        const InterpSource& S) {
  return interp_point<AT>(frac_i != nullptr, frac_j != nullptr, load_frac<AT>(frac_i, frac_j, idx), S);
}

// the 8 gathered values of one series (4 corners x 2 time levels), loaded together so that all the
// gathers of a point can be in flight at once
template <class AT> struct Corners8 { AT a[8]; };

template <class AT>
__device__ __forceinline__ Corners8<AT> gather8(const AT* __restrict__ data, const InterpPoint<AT>& p, const InterpSource& S, bool same) {
  Corners8<AT> c;
  const AT* d1 = data + S.o1;
  c.a[0] = __ldg(d1 + p.o_mm); c.a[1] = __ldg(d1 + p.o_mp); c.a[2] = __ldg(d1 + p.o_pm); c.a[3] = __ldg(d1 + p.o_pp);
  if (!same) {
    const AT* d2 = data + S.o2;
    c.a[4] = __ldg(d2 + p.o_mm); c.a[5] = __ldg(d2 + p.o_mp); c.a[6] = __ldg(d2 + p.o_pm); c.a[7] = __ldg(d2 + p.o_pp);
  } else {
    c.a[4] = c.a[5] = c.a[6] = c.a[7] = (AT)0;
  }
  return c;
}

template <class AT, class TT>
__device__ __forceinline__ auto blend8(const Corners8<AT>& c, const InterpPoint<AT>& p, TT nt, bool same) -> decltype(AT() * TT()) {
  using W = decltype(AT() * TT());
  const AT p1 = add_rn(add_rn(add_rn(mul_rn(p.w1, c.a[0]), mul_rn(p.w3, c.a[1])), mul_rn(p.w5, c.a[2])), mul_rn(p.w7, c.a[3]));
  if (same) return (W)p1;
  const AT p2 = add_rn(add_rn(add_rn(mul_rn(p.w1, c.a[4]), mul_rn(p.w3, c.a[5])), mul_rn(p.w5, c.a[6])), mul_rn(p.w7, c.a[7]));
  return add_rn(mul_rn((W)p2, (W)nt), mul_rn((W)p1, (W)sub_rn((TT)1, nt)));
}

template <class AT>
__device__ __forceinline__ AT bilinear(const AT* __restrict__ d, const InterpPoint<AT>& p) {
  return add_rn(add_rn(add_rn(mul_rn(p.w1, __ldg(d + p.o_mm)), mul_rn(p.w3, __ldg(d + p.o_mp))),
                       mul_rn(p.w5, __ldg(d + p.o_pm))),
                mul_rn(p.w7, __ldg(d + p.o_pp)));
}

// one series at the two bracketing time levels: ifelse(n₁ == n₂, ψ₁, ψ₂ñ + ψ₁(1 − ñ))
template <class AT, class TT>
__device__ __forceinline__ auto interp_series(const AT* __restrict__ data, const InterpPoint<AT>& p, const InterpSource& S,
                                              TT nt, bool same) -> decltype(AT() * TT()) {
  using W = decltype(AT() * TT());
  const AT p1 = bilinear<AT>(data + S.o1, p);
  if (same) return (W)p1;
  const AT p2 = bilinear<AT>(data + S.o2, p);
  return add_rn(mul_rn((W)p2, (W)nt), mul_rn((W)p1, (W)sub_rn((TT)1, nt)));
}

// field f of descriptor d: sum over its summands (`nothing` contributes the literal 0, :143)
template <class AT, class TT>
__device__ __forceinline__ auto interp_field(const NeInterpDesc& d, int f, const InterpPoint<AT>& p, const InterpSource& S,
                                             TT nt, bool same) -> decltype(AT() * TT()) {
  using W = decltype(AT() * TT());
  W total = 0;
  const int ns = d.n_summands[f];
  for (int s = 0; s < ns; ++s) {
    const AT* data = (const AT*)d.series[f][s].data;
    const W val = data ? interp_series<AT, TT>(data, p, S, nt, same) : (W)0;
    total = (s == 0) ? val : add_rn(total, val);
  }
  return total;
}

}  // namespace ne
