// ne_common.cuh — shared host/device helpers for libne_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <type_traits>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: a stub until a tool (nsys, ncu --nvtx) injects itself

#include "../../include/ne_b200.h"

namespace ne {

// One NVTX range per C-ABI entry point, named after it: a timeline shows the phases of update_state! (interpolation, the three
// turbulent-flux solves, sea-ice-ocean, assembly, radiation, ring loads) as the host enqueues them; `ncu --nvtx --nvtx-include`
// selects the kernels of one phase.  Costs a pointer test per call when no tool is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define NE_NVTX() ::ne::NvtxRange ne_nvtx_range_(__func__)

// ---- error plumbing (thread-local message, negative codes; include/ne_b200.h) ----------------
void set_error(const char* fmt, ...);
int cuda_error(cudaError_t e, const char* where);

#define NE_REQUIRE(cond, ...)              \
  do {                                     \
    if (!(cond)) {                         \
      ne::set_error(__VA_ARGS__);          \
      return NE_E_INVALID;                 \
    }                                      \
  } while (0)

#define NE_NO_VARIANT(...)                 \
  do {                                     \
    ne::set_error(__VA_ARGS__);            \
    return NE_E_NO_VARIANT;                \
  } while (0)

#define NE_CUDA_CHECK_LAUNCH(where)                       \
  do {                                                    \
    cudaError_t e__ = cudaGetLastError();                 \
    if (e__ != cudaSuccess) return ne::cuda_error(e__, where); \
  } while (0)

// ---- exchange-grid layout ------------------------------------------------------------------
// parent (nx+2hx) x (ny+2hy), column-major; reference index (i,j) -> (i+hx-1) + (j+hy-1)*sx.
struct Layout {
  int64_t sx;       // row stride in elements
  int64_t off;      // (hx-1) + (hy-1)*sx
  int32_t ni, nj;   // launch extent
  int32_t i_lo, j_lo;
  int32_t hy;
  int32_t pad_;
  __host__ __device__ inline int64_t at(int32_t i, int32_t j) const { return off + i + (int64_t)j * sx; }
};

inline Layout make_layout(const NeExchangeGrid& g) {
  Layout L;
  L.sx = g.nx + 2 * g.hx;
  L.off = (g.hx - 1) + (g.hy - 1) * L.sx;
  L.ni = (int32_t)(g.i_hi - g.i_lo + 1);
  L.nj = (int32_t)(g.j_hi - g.j_lo + 1);
  L.i_lo = (int32_t)g.i_lo;
  L.j_lo = (int32_t)g.j_lo;
  L.hy = (int32_t)g.hy;
  L.pad_ = 0;
  return L;
}

// launch range must stay inside the parent, with `sx`/`sy` extra cells for the 2-point stencils
inline bool grid_ok(const NeExchangeGrid& g, int stencil_lo, int stencil_hi) {
  if (g.nx <= 0 || g.ny <= 0 || g.hx < 0 || g.hy < 0) return false;
  if (g.i_hi < g.i_lo || g.j_hi < g.j_lo) return false;
  if (g.i_lo - stencil_lo < 1 - g.hx || g.i_hi + stencil_hi > g.nx + g.hx) return false;
  if (g.j_lo - stencil_lo < 1 - g.hy || g.j_hi + stencil_hi > g.ny + g.hy) return false;
  return true;
}

template <class FT>
__device__ __forceinline__ FT slot_at(const NeSlot& s, int64_t idx) {
  return s.ptr ? __ldg(static_cast<const FT*>(s.ptr) + idx) : static_cast<FT>(s.value);
}

// ---- typed math: float and double overloads, never fast-math ------------------------------------
// Float32 transcendentals are evaluated in Float64 and rounded once: the correctly rounded Float32 function (up to
// double rounding, probability ~2^-29 per call).  That is what Julia's Float32 `^` does (Base.Math.pow_body widens to
// Float64) and within half an ulp of its other Float32 functions; CUDA's powf / expf (2-4 ulp) would put a Float32
// q_sat (interface_states.jl:56-59) 1e-7 away from the oracle's.
__device__ __forceinline__ float  m_exp(float x)   { return (float)exp((double)x); }
__device__ __forceinline__ double m_exp(double x)  { return exp(x); }
__device__ __forceinline__ float  m_log(float x)   { return (float)log((double)x); }
__device__ __forceinline__ double m_log(double x)  { return log(x); }
__device__ __forceinline__ float  m_atan(float x)  { return (float)atan((double)x); }
__device__ __forceinline__ double m_atan(double x) { return atan(x); }
__device__ __forceinline__ float  m_cbrt(float x)  { return (float)cbrt((double)x); }
__device__ __forceinline__ double m_cbrt(double x) { return cbrt(x); }
__device__ __forceinline__ float  m_sqrt(float x)  { return sqrtf(x); }
__device__ __forceinline__ double m_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float  m_pow(float x, float y)   { return (float)pow((double)x, (double)y); }
__device__ __forceinline__ double m_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ double m_pow(float x, double y)  { return pow((double)x, y); }
__device__ __forceinline__ double m_pow(double x, float y)  { return pow(x, (double)y); }
__device__ __forceinline__ float  m_abs(float x)   { return fabsf(x); }
__device__ __forceinline__ double m_abs(double x)  { return fabs(x); }
__device__ __forceinline__ float  m_cos(float x)   { return cosf(x); }
__device__ __forceinline__ double m_cos(double x)  { return cos(x); }

// min/max with the reference's promotion (float∘double -> double); inputs are never NaN here
template <class A, class B>
__device__ __forceinline__ auto mn(A a, B b) -> decltype(a + b) {
  using W = decltype(a + b);
  return ((W)b < (W)a) ? (W)b : (W)a;
}
template <class A, class B>
__device__ __forceinline__ auto mx(A a, B b) -> decltype(a + b) {
  using W = decltype(a + b);
  return ((W)a < (W)b) ? (W)b : (W)a;
}
template <class T> __device__ __forceinline__ T sq(T x) { return x * x; }
template <class T> __device__ __forceinline__ T cube(T x) { return x * x * x; }
template <class T> __device__ __forceinline__ T clampv(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }
// x^4 / x^6: Float32 powers above 3 are evaluated in Float64 and rounded (Base.^(::Float32, ::Integer))
__device__ __forceinline__ double pow4(double x) { double x2 = x * x; return x2 * x2; }
__device__ __forceinline__ float pow4(float x) { return (float)pow4((double)x); }
__device__ __forceinline__ double pow6(double x) { double x2 = x * x; return x2 * x2 * x2; }
__device__ __forceinline__ float pow6(float x) { return (float)pow6((double)x); }

template <class T> struct Inf;
template <> struct Inf<double> { __device__ static double v() { return __longlong_as_double(0x7ff0000000000000LL); } };
template <> struct Inf<float> { __device__ static float v() { return __int_as_float(0x7f800000); } };

}  // namespace ne
