// ne_series_ring.cu — FieldTimeSeries window on the device.
//
// The reference's update_state!(::PrescribedAtmosphere) (src/Atmospheres/prescribed_atmosphere.jl:154-162) calls
// Oceananigans' update_field_time_series! for every series; when the interpolating time indices leave the in-memory
// window the whole window is re-read and re-set synchronously (set!(fts), JRA55_field_time_series.jl:60-76, :78-124:
// set_region_data! + fill_halo_regions!).  Here each series' device array is a ring of slices and ONE slice at a
// time is replaced on the ring's own copy stream while the interface kernels run on the compute stream:
//   pinned host raw slice --cudaMemcpyAsync--> device staging --slot_fill_kernel--> ring slot (with halos).
// Ordering is by events only; the host never waits.
#include <vector>

#include "ne_common.cuh"

#define NE_CUDA_TRY(expr, where)                             \
  do {                                                       \
    cudaError_t e__ = (expr);                                \
    if (e__ != cudaSuccess) return ne::cuda_error(e__, where); \
  } while (0)

namespace ne {

struct SeriesRing {
  NeSeriesRingDesc d;
  int device;
  cudaStream_t copy;
  void* staging;                      // n_series raw slices
  std::vector<cudaEvent_t> ready;     // slot s has been written
  std::vector<char> ever_loaded;
  cudaEvent_t released;               // readers enqueued before the last release
  bool have_release;
};

struct SlotFillArgs {
  const void* raw[NE_RING_MAX_SERIES];
  void* dst[NE_RING_MAX_SERIES];
  int conv_kind[NE_RING_MAX_SERIES];
  double conv_a[NE_RING_MAX_SERIES], conv_b[NE_RING_MAX_SERIES];
  int has_missing[NE_RING_MAX_SERIES];
  double missing_value[NE_RING_MAX_SERIES];
  int mangling[NE_RING_MAX_SERIES];
  long long nx, ny, hx, hy;
  long long raw_nx, raw_ny, di, dj;
  int periodic_x;
  int region_kind, column_interpolation;
  long long col_im, col_ip, col_jm, col_jp;
  double col_wx, col_wy;
};

// mirrored (zero-flux) halo index: halo cell -k (k = 1..h) takes interior cell k - 1; n - 1 + k takes n - k
__device__ __forceinline__ long long mirror_index(long long i, long long n) {
  if (i < 0) i = -i - 1;
  if (i >= n) i = 2 * n - 1 - i;
  return min(max(i, 0LL), n - 1);
}

// mangle + nan_convert_missing (set_region_data.jl:48-53, :162-163): file indices clamped to the file extent, ShiftSouth reads
// j - 1, AverageNorthSouth averages j and j + 1, the missing value becomes NaN
template <typename T>
__device__ __forceinline__ T ring_read(const T* __restrict__ raw, const SlotFillArgs& a, int mg, long long fi, long long fj, bool hm, T mv) {
  const long long ii = min(max(fi, 0LL), a.raw_nx - 1);
  if (mg == NE_MANGLE_SHIFT_SOUTH) fj -= 1;
  const long long j0 = min(max(fj, 0LL), a.raw_ny - 1);
  T v = raw[j0 * a.raw_nx + ii];
  if (mg == NE_MANGLE_AVERAGE_NORTH_SOUTH) {
    const long long j1 = min(max(fj + 1, 0LL), a.raw_ny - 1);
    v = (v + raw[j1 * a.raw_nx + ii]) / (T)2;
  }
  if (hm && v == mv) v = (T)NAN;
  return v;
}

// blend(::Linear) / blend(::Nearest) of a Column region (set_region_data.jl:168-193): land corners (NaN) are dropped and the
// weights renormalised over the wet ones; NaN only when all four are land.  Every operation in the series element type, in
// the reference's order (this translation unit is compiled without FMA contraction of these expressions: explicit
// intrinsics below).
template <typename T> __device__ __forceinline__ T t_mul(T x, T y);
template <> __device__ __forceinline__ float t_mul<float>(float x, float y) { return __fmul_rn(x, y); }
template <> __device__ __forceinline__ double t_mul<double>(double x, double y) { return __dmul_rn(x, y); }
template <typename T> __device__ __forceinline__ T t_add(T x, T y);
template <> __device__ __forceinline__ float t_add<float>(float x, float y) { return __fadd_rn(x, y); }
template <> __device__ __forceinline__ double t_add<double>(double x, double y) { return __dadd_rn(x, y); }

template <typename T>
__device__ __forceinline__ T column_blend(const T* __restrict__ raw, const SlotFillArgs& a, int mg, bool hm, T mv) {
  const T wx = (T)a.col_wx, wy = (T)a.col_wy;
  if (a.column_interpolation == NE_COLUMN_NEAREST) {
    const long long i = wx >= (T)0.5 ? a.col_ip : a.col_im, j = wy >= (T)0.5 ? a.col_jp : a.col_jm;
    const T near = ring_read<T>(raw, a, mg, i, j, hm, mv);
    if (!isnan(near)) return near;        // the closest corner is land: fall back to the NaN-aware linear blend
  }
  const T d00 = ring_read<T>(raw, a, mg, a.col_im, a.col_jm, hm, mv), d10 = ring_read<T>(raw, a, mg, a.col_ip, a.col_jm, hm, mv);
  const T d01 = ring_read<T>(raw, a, mg, a.col_im, a.col_jp, hm, mv), d11 = ring_read<T>(raw, a, mg, a.col_ip, a.col_jp, hm, mv);
  const T cx = t_add<T>((T)1, -wx), cy = t_add<T>((T)1, -wy);
  const T w00 = isnan(d00) ? (T)0 : t_mul<T>(cx, cy), w10 = isnan(d10) ? (T)0 : t_mul<T>(wx, cy);
  const T w01 = isnan(d01) ? (T)0 : t_mul<T>(cx, wy), w11 = isnan(d11) ? (T)0 : t_mul<T>(wx, wy);
  const T sw = t_add<T>(t_add<T>(t_add<T>(w00, w10), w01), w11);
  const T num = t_add<T>(t_add<T>(t_add<T>(t_mul<T>(w00, isnan(d00) ? (T)0 : d00), t_mul<T>(w10, isnan(d10) ? (T)0 : d10)),
                                  t_mul<T>(w01, isnan(d01) ? (T)0 : d01)), t_mul<T>(w11, isnan(d11) ? (T)0 : d11));
  if (sw == (T)0) return (T)NAN;
  return num / sw;
}

// One thread per ring-slice element (interior and halos), blockIdx.y = series.  Each element is ONE read of the raw
// slice (L2-resident: the staging copy has just landed) and one coalesced write: HBM/L2-bound, 2 x sizeof(T) bytes
// per element.  The arithmetic is the reference's, one rounding per operation in the series element type.
// A halo element is first mapped to the interior cell whose value fill_halo_regions! copies, then read like it.
template <typename T>
__global__ void __launch_bounds__(256) slot_fill_kernel(const SlotFillArgs a) {
  const int s = blockIdx.y;
  const long long W = a.nx + 2 * a.hx, H = a.ny + 2 * a.hy;
  const T* __restrict__ raw = (const T*)a.raw[s];
  T* __restrict__ dst = (T*)a.dst[s];
  const int kind = a.conv_kind[s];
  const T ca = (T)a.conv_a[s], cb = (T)a.conv_b[s], mv = (T)a.missing_value[s];
  const bool hm = a.has_missing[s] != 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < W * H; e += (long long)gridDim.x * blockDim.x) {
    const long long jp = e / W, ip = e - jp * W;
    long long i = ip - a.hx, j = jp - a.hy;
    if (a.periodic_x) {
      i %= a.nx;
      if (i < 0) i += a.nx;
    } else {
      i = mirror_index(i, a.nx);
    }
    j = mirror_index(j, a.ny);
    T v;
    if (a.region_kind == NE_REGION_COLUMN) v = column_blend<T>(raw, a, a.mangling[s], hm, mv);
    else v = ring_read<T>(raw, a, a.mangling[s], i + a.di, j + a.dj, hm, mv);   // read_data(…, ::BoundingBoxOffset, …)
    switch (kind) {
      case NE_CONV_NEGATE: v = -v; break;
      case NE_CONV_ADD: v = v + ca; break;
      case NE_CONV_SUB: v = v - ca; break;
      case NE_CONV_MUL: v = v * ca; break;
      case NE_CONV_DIV: v = v / ca; break;
      case NE_CONV_MUL_DIV: v = (v * ca) / cb; break;
      default: break;
    }
    dst[e] = v;
  }
}

static size_t elem_bytes(const NeSeriesRingDesc& d) { return d.dtype == NE_F64 ? 8 : 4; }

}  // namespace ne

extern "C" {

int ne_series_ring_create(void** handle, const NeSeriesRingDesc* d) {
  NE_REQUIRE(handle != nullptr && d != nullptr, "series ring: null argument");
  NE_REQUIRE(d->n_series >= 1 && d->n_series <= NE_RING_MAX_SERIES, "series ring: n_series out of range");
  NE_REQUIRE(d->n_slots >= 2 && d->n_slots <= 4096, "series ring: needs 2 to 4096 slots (the two interpolating slices must be resident together)");
  NE_REQUIRE(d->dtype == NE_F32 || d->dtype == NE_F64, "series ring: dtype must be NE_F32 or NE_F64");
  NE_REQUIRE(d->nx >= 1 && d->ny >= 1 && d->hx >= 0 && d->hy >= 0, "series ring: bad source grid size");
  NE_REQUIRE(d->hx <= d->nx && d->hy <= d->ny, "series ring: halo wider than the interior");
  for (int k = 0; k < d->n_series; ++k) {
    NE_REQUIRE(d->ring[k] != nullptr, "series ring: null ring pointer");
    NE_REQUIRE(d->conv_kind[k] >= NE_CONV_NONE && d->conv_kind[k] <= NE_CONV_MUL_DIV, "series ring: unknown unit conversion");
    NE_REQUIRE(d->mangling[k] >= NE_MANGLE_NONE && d->mangling[k] <= NE_MANGLE_AVERAGE_NORTH_SOUTH, "series ring: unknown mangling");
  }
  NE_REQUIRE(d->raw_nx >= 0 && d->raw_ny >= 0 && d->di >= 0 && d->dj >= 0, "series ring: negative raw extent or region offset");
  NE_REQUIRE(d->region_kind == NE_REGION_BOX || d->region_kind == NE_REGION_COLUMN, "series ring: unknown region kind");
  if (d->region_kind == NE_REGION_COLUMN) {
    const int64_t rnx = d->raw_nx ? d->raw_nx : d->nx, rny = d->raw_ny ? d->raw_ny : d->ny;
    NE_REQUIRE(d->nx == 1 && d->ny == 1, "series ring: a Column region fills a 1 x 1 series");
    NE_REQUIRE(d->column_interpolation == NE_COLUMN_LINEAR || d->column_interpolation == NE_COLUMN_NEAREST,
               "series ring: unknown Column interpolation");
    NE_REQUIRE(d->col_i_minus >= 0 && d->col_i_minus < rnx && d->col_i_plus >= 0 && d->col_i_plus < rnx &&
               d->col_j_minus >= 0 && d->col_j_minus < rny && d->col_j_plus >= 0 && d->col_j_plus < rny,
               "series ring: Column bracketing indices outside the raw slice");
    NE_REQUIRE(d->col_wx >= 0 && d->col_wx <= 1 && d->col_wy >= 0 && d->col_wy <= 1, "series ring: Column weights outside [0, 1]");
  }
  ne::SeriesRing* r = new ne::SeriesRing();
  r->d = *d;
  r->copy = nullptr;
  r->staging = nullptr;
  r->released = nullptr;
  r->have_release = false;
  r->ever_loaded.assign(d->n_slots, 0);
  if (r->d.raw_nx == 0) r->d.raw_nx = d->nx;
  if (r->d.raw_ny == 0) r->d.raw_ny = d->ny;
  const size_t raw_bytes = (size_t)r->d.raw_nx * r->d.raw_ny * ne::elem_bytes(*d);
  cudaError_t e = cudaGetDevice(&r->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->copy, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&r->staging, raw_bytes * d->n_series);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->released, cudaEventDisableTiming);
  for (int s = 0; s < d->n_slots && e == cudaSuccess; ++s) {
    cudaEvent_t ev;
    e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) r->ready.push_back(ev);
  }
  if (e != cudaSuccess) {
    const int rc = ne::cuda_error(e, "ne_series_ring_create");
    ne_series_ring_destroy(r);
    cudaGetLastError();
    return rc;
  }
  *handle = r;
  return NE_OK;
}

int ne_series_ring_destroy(void* handle) {
  ne::SeriesRing* r = (ne::SeriesRing*)handle;
  if (!r) return NE_OK;
  if (r->copy) cudaStreamSynchronize(r->copy);   // a copy in flight reads host memory the caller is about to free
  for (cudaEvent_t ev : r->ready) cudaEventDestroy(ev);
  if (r->released) cudaEventDestroy(r->released);
  if (r->staging) cudaFree(r->staging);
  if (r->copy) cudaStreamDestroy(r->copy);
  delete r;
  return NE_OK;
}

int ne_series_ring_load(void* handle, int32_t slot, const void* const* host_raw) { NE_NVTX();
  ne::SeriesRing* r = (ne::SeriesRing*)handle;
  NE_REQUIRE(r != nullptr && host_raw != nullptr, "series ring load: null argument");
  const NeSeriesRingDesc& d = r->d;
  NE_REQUIRE(slot >= 0 && slot < d.n_slots, "series ring load: slot out of range");
  const size_t eb = ne::elem_bytes(d), raw_bytes = (size_t)d.raw_nx * d.raw_ny * eb;
  const size_t slice_bytes = (size_t)(d.nx + 2 * d.hx) * (d.ny + 2 * d.hy) * eb;
  // the slot may still be read by kernels enqueued before the last release
  if (r->have_release) NE_CUDA_TRY(cudaStreamWaitEvent(r->copy, r->released, 0), "series ring load (wait readers)");
  ne::SlotFillArgs a;
  for (int k = 0; k < d.n_series; ++k) {
    NE_REQUIRE(host_raw[k] != nullptr, "series ring load: null host slice");
    char* st = (char*)r->staging + k * raw_bytes;
    NE_CUDA_TRY(cudaMemcpyAsync(st, host_raw[k], raw_bytes, cudaMemcpyHostToDevice, r->copy), "series ring load (H2D)");
    a.raw[k] = st;
    a.dst[k] = (char*)d.ring[k] + (size_t)slot * slice_bytes;
    a.conv_kind[k] = d.conv_kind[k];
    a.conv_a[k] = d.conv_a[k];
    a.conv_b[k] = d.conv_b[k];
    a.has_missing[k] = d.has_missing[k];
    a.missing_value[k] = d.missing_value[k];
    a.mangling[k] = d.mangling[k];
  }
  a.raw_nx = d.raw_nx; a.raw_ny = d.raw_ny; a.di = d.di; a.dj = d.dj;
  a.nx = d.nx; a.ny = d.ny; a.hx = d.hx; a.hy = d.hy; a.periodic_x = d.periodic_x;
  a.region_kind = d.region_kind; a.column_interpolation = d.column_interpolation;
  a.col_im = d.col_i_minus; a.col_ip = d.col_i_plus; a.col_jm = d.col_j_minus; a.col_jp = d.col_j_plus;
  a.col_wx = d.col_wx; a.col_wy = d.col_wy;
  const long long elems = (long long)(d.nx + 2 * d.hx) * (d.ny + 2 * d.hy);
  const unsigned bx = (unsigned)std::min<long long>((elems + 255) / 256, 148LL * 8);
  dim3 grid(bx, (unsigned)d.n_series);
  if (d.dtype == NE_F64) ne::slot_fill_kernel<double><<<grid, 256, 0, r->copy>>>(a);
  else ne::slot_fill_kernel<float><<<grid, 256, 0, r->copy>>>(a);
  NE_CUDA_TRY(cudaGetLastError(), "series ring load (slot_fill_kernel)");
  NE_CUDA_TRY(cudaEventRecord(r->ready[slot], r->copy), "series ring load (record)");
  r->ever_loaded[slot] = 1;
  return NE_OK;
}

int ne_series_ring_acquire(void* handle, int32_t slot, void* compute_stream) { NE_NVTX();
  ne::SeriesRing* r = (ne::SeriesRing*)handle;
  NE_REQUIRE(r != nullptr, "series ring acquire: null handle");
  NE_REQUIRE(slot >= 0 && slot < r->d.n_slots, "series ring acquire: slot out of range");
  NE_REQUIRE(r->ever_loaded[slot], "series ring acquire: slot %d has never been loaded", (int)slot);
  NE_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)compute_stream, r->ready[slot], 0), "series ring acquire");
  return NE_OK;
}

int ne_series_ring_release(void* handle, void* compute_stream) { NE_NVTX();
  ne::SeriesRing* r = (ne::SeriesRing*)handle;
  NE_REQUIRE(r != nullptr, "series ring release: null handle");
  NE_CUDA_TRY(cudaEventRecord(r->released, (cudaStream_t)compute_stream), "series ring release");
  r->have_release = true;
  return NE_OK;
}

}
