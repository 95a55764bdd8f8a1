// ne_fastmath.cuh — branch-free Float64 elementary functions and the piecewise-polynomial ψ tables
// of the specialised a–o solve (ne_flux_fast.cuh).
//
// Why: ncu on the libdevice-based solve (profiles/r01_notes.md, "v2") shows only 37 % of the issued
// instructions are FP64 arithmetic; 32 % are UMOV/IMAD/MOV that materialise libdevice's 64-bit
// polynomial constants and 8 % are the special-case branches of log/exp/cbrt/atan.  The functions
// below take their coefficients from the kernel-parameter constant bank (direct c[][] operands),
// assume the argument ranges the solve guarantees (positive, normal), and are accurate to ≲ 2 ulp —
// NOT fast-math: no approximate-only MUFU result is ever returned, every seed is refined to full
// double precision by Newton steps.  The same code compiles on the host (seeds emulated in Float32)
// so tests/test_fastmath.py checks every function against long double without a GPU.
//
// ψ tables: the Edson et al. (2013) stability functions (similarity_theory_turbulent_fluxes.jl:501-532,
// 586-618) are smooth on each side of ζ = 0; on quarter-octave intervals 2^-6 ≤ |ζ| < 2^7 (plus one
// interval |ζ| < 2^-6 per side) they are replaced by degree-13 polynomials
// interpolated at Chebyshev nodes from the closed forms evaluated in long double; the fit is
// verified on the host at build time (max abs error ≤ 2e-15 required, else the closed-form kernel
// is used).  ψ_m and ψ_s coefficients are interleaved so one 16-byte load feeds both Horner chains.
#pragma once

#include <stdint.h>

#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define NE_HD __host__ __device__ __forceinline__
#else
#define NE_HD inline
#endif

namespace ne {
namespace fm {

constexpr int LOG_N = 64;       // log table entries (invc, logc)
constexpr int LOG_DEG = 6;      // log1p(r) = r + r²·P(r), P has LOG_DEG coefficients (|r| ≤ 2^-7)
constexpr int EXP_DEG = 11;     // e^r on |r| ≤ ln2/2
// Resolution of the ψ tables.  Round 1 shipped quarter octaves with degree 13 (30 doubles per record: 15 LDS.128 per
// lookup).  ncu (profiles/r02_ncu_tab2_v1_summary.txt) shows the L1TEX data pipe at 80-88 % of its peak — shared-memory
// wavefronts of exactly these lookups — while FP64 and issue sit near 55 %: eighth octaves from 2^-8 with degree 9 need
// 11 LDS.128 and 8 fewer DFMA per lookup for the same verified accuracy (3e-16; tools/fastmath_check.cu), at 44 KB
// instead of 27 KB of shared memory per CTA.  (Quarter octaves need degree 11 from 2^-7, degree 8 misses 2e-15.)
#ifndef NE_PSI_SUB_BITS
#define NE_PSI_SUB_BITS 3
#endif
#ifndef NE_PSI_DEG
#define NE_PSI_DEG 9
#endif
constexpr int PSI_DEG = NE_PSI_DEG;
constexpr int PSI_SUB_BITS = NE_PSI_SUB_BITS;   // 2^PSI_SUB_BITS intervals per octave
constexpr int PSI_SUB = 1 << PSI_SUB_BITS;
#ifndef NE_PSI_OCT_LO
#define NE_PSI_OCT_LO -8
#endif
constexpr int PSI_OCT_LO = NE_PSI_OCT_LO;  // table covers 2^PSI_OCT_LO ≤ |ζ| < 2^7 on each side (+ one record for [0, 2^PSI_OCT_LO))
constexpr int PSI_OCT_HI = 7;
constexpr int PSI_NQ = PSI_SUB * (PSI_OCT_HI - PSI_OCT_LO);  // intervals per side above 2^PSI_OCT_LO
constexpr int PSI_NS = PSI_NQ + 1;                     // records per side: [0, 2^-6) then the quarter octaves
constexpr int PSI_NI = 2 * PSI_NS;                     // unstable side (ζ < 0) first, then the stable side
constexpr int PSI_REC = 2 + 2 * (PSI_DEG + 1);         // (a, b) of w = a|ζ| + b, then (c_m, c_s) pairs
constexpr int TINY_DEG = 9;                            // |ζ| < 2^TINY_EXP (the ψ(ℓ/L★) terms): low-degree records
constexpr int TINY_EXP = -7;
constexpr int TINY_REC = 2 + 2 * (TINY_DEG + 1);       // record 0: unstable side, record 1: stable side
constexpr int MICRO_DEG = 5;                           // |ζ| < 2^MICRO_EXP: where ψ(ℓ/L★) sits once u★ has left its initial guess
constexpr int MICRO_EXP = -12;
constexpr int MICRO_REC = 2 + 2 * (MICRO_DEG + 1);
constexpr int TAB_LOG = 0;
constexpr int TAB_PSI = 2 * LOG_N;
constexpr int TAB_TINY = TAB_PSI + PSI_NI * PSI_REC;
constexpr int TAB_MICRO = TAB_TINY + 2 * TINY_REC;
constexpr int EXP2_N = 32;                              // 2^(j/32), j = 0..31: exp_lo
constexpr int TAB_EXP2 = TAB_MICRO + 2 * MICRO_REC;
constexpr int TAB_SIZE = TAB_EXP2 + EXP2_N;            // doubles
static_assert(TAB_SIZE % 2 == 0, "the table is staged with 16-byte copies");

struct MathConsts {
  double logp[LOG_DEG];      // P(r) = Σ logp[k] r^k
  double expp[EXP_DEG + 1];  // e^r ≈ Σ expp[k] r^k
  // 64-bit literals kept in the constant bank (an immediate costs two extra moves per use)
  double ln2, log2e, ln2_hi, ln2_lo, third;
  double log2e_n, ln2_n_hi, ln2_n_lo;   // EXP2_N log2(e), ln2/EXP2_N split (exp_lo)
};
inline void fill_literals(MathConsts& C) {
  C.ln2 = 0.693147180559945309417;
  C.log2e = 1.44269504088896340736;
  C.ln2_hi = 6.93147180369123816490e-01;
  C.ln2_lo = 1.90821492927058770002e-10;
  C.third = 0.33333333333333333;
  C.log2e_n = C.log2e * EXP2_N;
  C.ln2_n_hi = C.ln2_hi / EXP2_N;   // exact: powers of two
  C.ln2_n_lo = C.ln2_lo / EXP2_N;
}

// select-based min/max (inputs are never NaN here; IEEE fmin/fmax cost 6-7 instructions in FP64)
NE_HD double dmin(double a, double b) { return a < b ? a : b; }
NE_HD double dmax(double a, double b) { return a > b ? a : b; }

// ---- bit access -----------------------------------------------------------------------------------
NE_HD int32_t hi32(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  int64_t b; std::memcpy(&b, &x, 8); return (int32_t)(b >> 32);
#endif
}
NE_HD int32_t lo32(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  int64_t b; std::memcpy(&b, &x, 8); return (int32_t)(b & 0xffffffffLL);
#endif
}
NE_HD double mk64(int32_t hi, int32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  int64_t b = ((int64_t)hi << 32) | (int64_t)(uint32_t)lo; double x; std::memcpy(&x, &b, 8); return x;
#endif
}
NE_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}

// separately rounded product (never contracted into a neighbouring add, whatever the inlining context): the
// table iteration gives bit-identical iterates in every kernel that embeds it
NE_HD double mul_(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double p = a * b; return p;
#endif
}

// ---- operation policies ---------------------------------------------------------------------------
// Every FP64 operation of the hot path goes through a policy object: OpsPlain issues the separately rounded
// instruction and nothing else; OpsCount additionally counts it per thread, so an instrumented instantiation of
// the SAME kernel source reports the FP64 instructions a launch executes (bench.py's roofline numerator) without
// a profiler.  add/sub/mul are never contracted (explicit __d*_rn), fma is an explicit fused operation.
struct OpsPlain {
  NE_HD double fma(double a, double b, double c) { return fma_(a, b, c); }
  NE_HD double mul(double a, double b) { return mul_(a, b); }
  NE_HD double add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
  }
  NE_HD double sub(double a, double b) { return add(a, -b); }
  NE_HD void other(int) {}
  NE_HD void trip() {}
  NE_HD void flush(unsigned long long*) {}
};
struct OpsCount {
  unsigned long long n_fma = 0, n_mul = 0, n_add = 0;
  OpsPlain base;
  NE_HD double fma(double a, double b, double c) { ++n_fma; return base.fma(a, b, c); }
  NE_HD double mul(double a, double b) { ++n_mul; return base.mul(a, b); }
  NE_HD double add(double a, double b) { ++n_add; return base.add(a, b); }
  NE_HD double sub(double a, double b) { ++n_add; return base.sub(a, b); }
  unsigned long long n_other = 0, n_trips = 0;
  NE_HD void other(int n) { n_other += (unsigned)n; }   // library code the policy cannot see (IEEE division, libdevice): a declared estimate, reported apart
  NE_HD void trip() { ++n_trips; }
  // counts[0..4] += {fma, mul, add, other, thread trips}
  NE_HD void flush(unsigned long long* counts) {
#if defined(__CUDA_ARCH__)
    if (counts) {
      atomicAdd(counts + 0, n_fma); atomicAdd(counts + 1, n_mul); atomicAdd(counts + 2, n_add);
      atomicAdd(counts + 3, n_other); atomicAdd(counts + 4, n_trips);
    }
#else
    if (counts) { counts[0] += n_fma; counts[1] += n_mul; counts[2] += n_add; counts[3] += n_other; counts[4] += n_trips; }
#endif
    n_fma = n_mul = n_add = n_other = n_trips = 0;
  }
};

// ---- seeds (≈ 20 bits on the device; the host emulation rounds to Float32) ------------------------
NE_HD double rcp_seed(double x) {
#if defined(__CUDA_ARCH__)
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r;
#else
  return (double)(1.0f / (float)x) * (1.0 + 3e-7);
#endif
}
NE_HD double rsqrt_seed(double x) {
#if defined(__CUDA_ARCH__)
  double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r;
#else
  return (double)(1.0f / std::sqrt((float)x)) * (1.0 - 3e-7);
#endif
}

// 1/x, x finite, normal, non-zero (either sign): seed + two Newton steps
template <class O> NE_HD double rcp(O& o, double x) {
  double r = rcp_seed(x);
  double e = o.fma(-x, r, 1.0);
  r = o.fma(r, e, r);
  e = o.fma(-x, r, 1.0);
  return o.fma(r, e, r);
}
// a/b with one residual correction (≲ 1 ulp)
template <class O> NE_HD double div(O& o, double a, double b) {
  const double r = rcp(o, b);
  const double q = o.mul(a, r);
  return o.fma(o.fma(-b, q, a), r, q);
}
// sqrt(x), x > 0 normal: coupled Goldschmidt iteration + residual correction
template <class O> NE_HD double sqrt_pos(O& o, double x) {
  const double r = rsqrt_seed(x);
  double g = o.mul(x, r), h = o.mul(0.5, r);
  double e = o.fma(-h, g, 0.5);
  g = o.fma(g, e, g); h = o.fma(h, e, h);
  e = o.fma(-h, g, 0.5);
  g = o.fma(g, e, g); h = o.fma(h, e, h);
  return o.fma(o.fma(-g, g, x), h, g);
}
// cbrt(x), x > 0 normal.  x = 2^(3q)·m with m ∈ [1, 8): seed m^(-1/3) in Float32 (MUFU lg2/ex2 on the
// device), two Newton steps on r ↦ r + r(1 − m r³)/3, result m·r²·2^q.
template <class O> NE_HD double cbrt_pos(O& o, const MathConsts& C, double x) {
  const int32_t hi = hi32(x);
  const int32_t e = (hi >> 20) - 1023;                 // unbiased exponent
  const int32_t q = (e + 3072) * 21846 >> 16;          // floor((e + 3072)/3), exact for |e| ≤ 1100
  const int32_t q3 = q - 1024;                         // floor(e/3)
  const double m = mk64(hi - ((3 * q3) << 20), lo32(x));  // [1, 8)
#if defined(__CUDA_ARCH__)
  float lg, ex;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float)m));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-0.333333333f * lg));
  double r = (double)ex;
#else
  double r = (double)(float)std::exp2(-std::log2((float)m) / 3.0f) * (1.0 + 4e-7);
#endif
  const double third = C.third;
  double r2 = o.mul(r, r);
  double e1 = o.fma(o.mul(-m, r), r2, 1.0);
  r = o.fma(o.mul(r, third), e1, r);
  r2 = o.mul(r, r);
  e1 = o.fma(o.mul(-m, r), r2, 1.0);
  r = o.fma(o.mul(r, third), e1, r);
  double y = o.mul(o.mul(m, r), r);                    // m^(1/3) ∈ [1, 2)
  // one residual step on y: y ← y − (y³ − m)/(3y²) = y + (m − y³)·r²/3  (r ≈ 1/y)
  const double y2 = o.mul(y, y);
  y = o.fma(o.fma(-y2, y, m), o.mul(o.mul(r, r), third), y);
  return mk64(hi32(y) + (q3 << 20), lo32(y));
}

// ---- one third-order step instead of two Newton steps (tab2 kernel) ---------------------------------------------
// The seeds carry ≥ 20 bits (relative error e ≤ 2^-20).  A second-order (Newton) step squares e and has to be applied twice;
// one step built from the series of the function around the seed, truncated after e², leaves c·e³ ≤ 2^-60 — below half an
// ulp — with half the operations and half the dependent chain:
//   1/x      = r (1 + e + e²) (1 + O(e³)),                 e = 1 − x r
//   √x       = t (1 + e/2 + 3e²/8) (1 + O(e³)),            t = x r,  e = 1 − t r        (r ≈ 1/√x)
//   m^(-1/3) = r (1 + e/3 + 2e²/9) (1 + O(e³)),            e = 1 − m r³
// Results within 1.5 ulp (checked against long double in tools/fastmath_check.cu and on the device in
// tools/fastmath_gpu_check.cu); e itself is exact to its last bit because it comes out of one fused operation.
template <class O> NE_HD double rcp3(O& o, double x) {
  const double r = rcp_seed(x);
  const double e = o.fma(-x, r, 1.0);
  return o.fma(r, o.fma(e, e, e), r);
}
template <class O> NE_HD double sqrt3(O& o, double x) {
  const double r = rsqrt_seed(x);
  const double t = o.mul(x, r);
  const double e = o.fma(-t, r, 1.0);
  return o.fma(o.mul(t, e), o.fma(e, 0.375, 0.5), t);
}
template <class O> NE_HD double cbrt3(O& o, const MathConsts& C, double x) {
  const int32_t hi = hi32(x);
  const int32_t e = (hi >> 20) - 1023;                 // unbiased exponent
  const int32_t q = (e + 3072) * 21846 >> 16;          // floor((e + 3072)/3), exact for |e| ≤ 1100
  const int32_t q3 = q - 1024;                         // floor(e/3)
  const double m = mk64(hi - ((3 * q3) << 20), lo32(x));  // [1, 8)
#if defined(__CUDA_ARCH__)
  float lg, ex;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float)m));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-0.333333333f * lg));
  const double r = (double)ex;
#else
  const double r = (double)(float)std::exp2(-std::log2((float)m) / 3.0f) * (1.0 + 4e-7);
#endif
  const double r2 = o.mul(r, r);
  const double mr = o.mul(m, r);
  const double e1 = o.fma(-mr, r2, 1.0);                              // 1 − m r³
  const double r1 = o.fma(o.mul(r, e1), o.fma(e1, 2.0 / 9.0, C.third), r);   // m^(-1/3)
  const double y = o.mul(o.mul(m, r1), r1);                           // m^(1/3) ∈ [1, 2)
  return mk64(hi32(y) + (q3 << 20), lo32(y));
}

// log(x), x > 0 normal.  x = 2^k z, z ∈ [√½, √2); z = c_i(1 + r), |r| ≤ 2^-7; table holds 1/c_i and log c_i.
template <class O> NE_HD double log_pos(O& o, const double* __restrict__ tab, const MathConsts& C, double x) {
  const int32_t hi = hi32(x);
  const int32_t tmp = hi - 0x3fe6a09e;
  const int32_t k = tmp >> 20;
  const int32_t i = (tmp >> 14) & (LOG_N - 1);
  const double z = mk64(hi - (k << 20), lo32(x));
  const double invc = tab[TAB_LOG + 2 * i], logc = tab[TAB_LOG + 2 * i + 1];
  const double r = o.fma(z, invc, -1.0);
  double p = C.logp[LOG_DEG - 1];
#pragma unroll
  for (int n = LOG_DEG - 2; n >= 0; --n) p = o.fma(p, r, C.logp[n]);
  const double r2 = o.mul(r, r);
  const double base = o.fma((double)k, C.ln2, logc);
  return o.add(base, o.fma(r2, p, r));
}

// The same with a table replicated per lane class (tab2 kernel).  A lookup with 32 unrelated indices into the 64 x 16-byte
// table costs 8 shared-memory wavefronts instead of the 4 its 512 bytes need: entries i and i + 8 share banks (ncu:
// 7.9 wavefronts per LDS.128, ideal 3.7).  In `rep`, entry i of lane class c = lane mod 8 sits at 16-byte chunk 8 i + c, i.e.
// in bank group c: the four lanes of a class can collide only with each other — at most 4 wavefronts, the ideal.
constexpr int LOG_REP = 8;
template <class O> NE_HD double log_pos_rep(O& o, const double* __restrict__ rep, int lane_class, const MathConsts& C, double x) {
  const int32_t hi = hi32(x);
  const int32_t tmp = hi - 0x3fe6a09e;
  const int32_t k = tmp >> 20;
  const int32_t i = (tmp >> 14) & (LOG_N - 1);
  const double z = mk64(hi - (k << 20), lo32(x));
  const double* e = rep + 2 * (LOG_REP * i + lane_class);
  const double invc = e[0], logc = e[1];
  const double r = o.fma(z, invc, -1.0);
  double p = C.logp[LOG_DEG - 1];
#pragma unroll
  for (int n = LOG_DEG - 2; n >= 0; --n) p = o.fma(p, r, C.logp[n]);
  const double r2 = o.mul(r, r);
  const double base = o.fma((double)k, C.ln2, logc);
  return o.add(base, o.fma(r2, p, r));
}

// exp(x), |x| ≤ 700.  x = k ln2 + r; 2^k applied by exponent arithmetic (result stays normal).
template <class O> NE_HD double exp_mid(O& o, const MathConsts& C, double x) {
  const double magic = 6755399441055744.0;  // 1.5·2^52
  const double t = o.fma(x, C.log2e, magic);
  const int32_t k = lo32(t);
  const double kf = o.sub(t, magic);
  double r = o.fma(-kf, C.ln2_hi, x);
  r = o.fma(-kf, C.ln2_lo, r);
  double p = C.expp[EXP_DEG];
#pragma unroll
  for (int n = EXP_DEG - 1; n >= 0; --n) p = o.fma(p, r, C.expp[n]);
  return mk64(hi32(p) + (k << 20), lo32(p));
}

// exp(x) to ~2e-12 relative, |x| ≤ 700: x = (32 k + j) ln2/32 + r, |r| ≤ ln2/64; 2^(j/32) from the table, e^r to degree 4
// (r^5/120 ≤ 1.3e-12).  For quantities whose own weight in the result is ≤ 1e-3 (the roughness length ℓs that only
// feeds ψ(ℓs/L★), |ψ| < 1e-3 against Π ~ 10): 9 FP64 operations in a chain of 6 instead of 15 in a chain of 13.
template <class O> NE_HD double exp_lo(O& o, const double* __restrict__ tab, const MathConsts& C, double x) {
  const double magic = 6755399441055744.0;  // 1.5·2^52
  const double t = o.fma(x, C.log2e_n, magic);
  const int32_t n = lo32(t);
  const double nf = o.sub(t, magic);
  double r = o.fma(-nf, C.ln2_n_hi, x);
  r = o.fma(-nf, C.ln2_n_lo, r);
  const double tj = tab[TAB_EXP2 + (n & (EXP2_N - 1))];
  double p = o.fma(r, 1.0 / 24.0, 1.0 / 6.0);
  p = o.fma(p, r, 0.5);
  p = o.fma(p, r, 1.0);
  p = o.fma(p, r, 1.0);
  const double y = o.mul(tj, p);
  return mk64(hi32(y) + ((n >> 5) << 20), lo32(y));
}

// the policy-free names used by the kernels of round 1 (same operations, same bits)
NE_HD double rcp(double x) { OpsPlain o; return rcp(o, x); }
NE_HD double div(double a, double b) { OpsPlain o; return div(o, a, b); }
NE_HD double sqrt_pos(double x) { OpsPlain o; return sqrt_pos(o, x); }
NE_HD double cbrt_pos(const MathConsts& C, double x) { OpsPlain o; return cbrt_pos(o, C, x); }
NE_HD double log_pos(const double* __restrict__ tab, const MathConsts& C, double x) { OpsPlain o; return log_pos(o, tab, C, x); }
NE_HD double exp_mid(const MathConsts& C, double x) { OpsPlain o; return exp_mid(o, C, x); }

// ---- ψ table lookup --------------------------------------------------------------------------------
// record index of ζ (branch-free); `outside` is set when |ζ| ≥ 2^7 or ζ is NaN (closed forms then)
NE_HD int psi_interval(double zeta, bool& outside) {
  const int32_t q = ((hi32(zeta) & 0x7fffffff) >> (20 - PSI_SUB_BITS)) - ((1023 + PSI_OCT_LO) << PSI_SUB_BITS);
  outside = q >= PSI_NQ;
  int32_t i = q + 1;
  i = i < 0 ? 0 : i;
  i = i > PSI_NQ ? PSI_NQ : i;
  return zeta < 0 ? i : i + PSI_NS;
}
NE_HD bool psi_is_micro(double zeta) {
  return ((hi32(zeta) & 0x7fffffff) >> 20) < 1023 + MICRO_EXP;
}
NE_HD bool psi_is_tiny(double zeta) {
  return ((hi32(zeta) & 0x7fffffff) >> 20) < 1023 + TINY_EXP;
}

// Horner split into even and odd parts (two half-length dependency chains per polynomial)
template <int DEG, class O>
NE_HD double poly_eo(O& o, const double* __restrict__ c, int stride, double w, double w2) {
  constexpr int TE = DEG & ~1, TO = (DEG & 1) ? DEG : DEG - 1;   // top even / odd degree
  double e = c[TE * stride], od = c[TO * stride];
#pragma unroll
  for (int k = TE - 2; k >= 0; k -= 2) e = o.fma(e, w2, c[k * stride]);
#pragma unroll
  for (int k = TO - 2; k >= 1; k -= 2) od = o.fma(od, w2, c[k * stride]);
  return o.fma(od, w, e);
}
template <int DEG>
NE_HD double poly_eo(const double* __restrict__ c, int stride, double w, double w2) {
  OpsPlain o; return poly_eo<DEG>(o, c, stride, w, w2);
}

// both ψ_m and ψ_s at the same |ζ| from interval record `rec`
template <class O> NE_HD void psi_pair(O& o, const double* __restrict__ rec, double az, double& pm, double& ps) {
  const double w = o.fma(az, rec[0], rec[1]);
  const double w2 = o.mul(w, w);
  pm = poly_eo<PSI_DEG>(o, rec + 2, 2, w, w2);
  ps = poly_eo<PSI_DEG>(o, rec + 3, 2, w, w2);
}
NE_HD void psi_pair(const double* __restrict__ rec, double az, double& pm, double& ps) { OpsPlain o; psi_pair(o, rec, az, pm, ps); }
// The record's (a, b) never has to be loaded: on [2^e (1 + k/S), 2^e (1 + (k+1)/S)) (S = PSI_SUB sub-intervals per octave)
// w = a|ζ| + b = 2 f − 1 with f the fraction of the mantissa below its top PSI_SUB_BITS bits — shifting those bits out of
// |ζ|'s mantissa gives u = 1 + f exactly and w = 2u − 3 in one fused operation: the same real number rounded once, i.e.
// the same bits as fma(|ζ|, a, b) with the stored power-of-two a and integer b (checked in tools/fastmath_check.cu).
// Below 2^PSI_OCT_LO (record 0 of a side) w = 2^(1 − PSI_OCT_LO) |ζ| − 1.  One LDS.128 (4.3 wavefronts) less per lookup.
template <class O> NE_HD double psi_w_from_bits(O& o, double az) {
  const int32_t hi = hi32(az), lo = lo32(az);
  const uint32_t mh = (((uint32_t)hi << PSI_SUB_BITS) & 0x000fffffu) | ((uint32_t)lo >> (32 - PSI_SUB_BITS));
  const double u = mk64((int32_t)(0x3ff00000u | mh), (int32_t)((uint32_t)lo << PSI_SUB_BITS));
  const double w_hi = o.fma(u, 2.0, -3.0);
  const double w_lo = o.fma(az, (double)(1ll << (1 - PSI_OCT_LO)), -1.0);
  return (hi >> 20) < 1023 + PSI_OCT_LO ? w_lo : w_hi;
}
template <class O> NE_HD void psi_pair_bits(O& o, const double* __restrict__ rec, double az, double& pm, double& ps) {
  const double w = psi_w_from_bits(o, az);
  const double w2 = o.mul(w, w);
  pm = poly_eo<PSI_DEG>(o, rec + 2, 2, w, w2);
  ps = poly_eo<PSI_DEG>(o, rec + 3, 2, w, w2);
}
// one of the two (which = 0: ψ_m, 1: ψ_s): same operations as psi_pair, so the same bits
template <class O> NE_HD double psi_single(O& o, const double* __restrict__ rec, double az, int which) {
  const double w = o.fma(az, rec[0], rec[1]);
  return poly_eo<PSI_DEG>(o, rec + 2 + which, 2, w, o.mul(w, w));
}
NE_HD double psi_single(const double* __restrict__ rec, double az, int which) { OpsPlain o; return psi_single(o, rec, az, which); }
// ψ_m(|ζ_u|) and ψ_s(|ζ_s|) for |ζ| < 2^TINY_EXP, both on the side (record) `rec`
template <class O> NE_HD void psi_tiny_pair(O& o, const double* __restrict__ rec, double azu, double azs, double& pm, double& ps) {
  const double wu = o.fma(azu, rec[0], rec[1]), ws = o.fma(azs, rec[0], rec[1]);
  pm = poly_eo<TINY_DEG>(o, rec + 2, 2, wu, o.mul(wu, wu));
  ps = poly_eo<TINY_DEG>(o, rec + 3, 2, ws, o.mul(ws, ws));
}
NE_HD void psi_tiny_pair(const double* __restrict__ rec, double azu, double azs, double& pm, double& ps) {
  OpsPlain o; psi_tiny_pair(o, rec, azu, azs, pm, ps);
}

// the same for |ζ| < 2^MICRO_EXP with the degree-MICRO_DEG records
template <class O> NE_HD void psi_micro_pair(O& o, const double* __restrict__ rec, double azu, double azs, double& pm, double& ps) {
  const double wu = o.fma(azu, rec[0], rec[1]), ws = o.fma(azs, rec[0], rec[1]);
  pm = poly_eo<MICRO_DEG>(o, rec + 2, 2, wu, o.mul(wu, wu));
  ps = poly_eo<MICRO_DEG>(o, rec + 3, 2, ws, o.mul(ws, ws));
}
NE_HD void psi_micro_pair(const double* __restrict__ rec, double azu, double azs, double& pm, double& ps) {
  OpsPlain o; psi_micro_pair(o, rec, azu, azs, pm, ps);
}

}  // namespace fm
}  // namespace ne
