// ne_api.cu — library plumbing: version, thread-local error, copies, FP64 peak microbenchmark.
#include <dlfcn.h>
#include <mutex>
#include <cstdlib>
#include <cstdarg>

#include "ne_common.cuh"

namespace ne {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_error(cudaError_t e, const char* where) {
  set_error("%s: CUDA error %d (%s)", where, (int)e, cudaGetErrorString(e));
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? NE_E_NO_DEVICE : NE_E_CUDA;
}

// Dependent DFMA chains: 8 independent accumulators per thread, enough warps to fill the
// FP64 pipe of every SM.  2 flop per DFMA.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace ne

extern "C" {

int ne_version(void) { return NE_ABI_VERSION; }

const char* ne_last_error(void) { return ne::g_error; }

int64_t ne_struct_size(const char* name) {
#define NE_SZ(T) if (std::strcmp(name, #T) == 0) return (int64_t)sizeof(T);
  if (!name) return -1;
  NE_SZ(NeSlot) NE_SZ(NeExchangeGrid) NE_SZ(NeTimeSeries) NE_SZ(NeTimeInterp) NE_SZ(NeInterpDesc)
  NE_SZ(NeFracIndexDesc) NE_SZ(NeThermoParams) NE_SZ(NeStabilityFn) NE_SZ(NeStabilityProfile)
  NE_SZ(NeRoughnessLength) NE_SZ(NeSubgridVelocity) NE_SZ(NeStopCriteria) NE_SZ(NePolynomialDrag)
  NE_SZ(NeTransferCoefficient) NE_SZ(NeLargeYeager) NE_SZ(NeFluxFormulation) NE_SZ(NeInterfaceProperties)
  NE_SZ(NeMediumProperties) NE_SZ(NeSurfaceRadiation) NE_SZ(NeAtmosOceanDesc) NE_SZ(NeAtmosSeaIceDesc) NE_SZ(NeLandHumidity) NE_SZ(NeAtmosLandDesc)
  NE_SZ(NeSeaIceOceanDesc) NE_SZ(NeSeaIceOceanStressDesc) NE_SZ(NeAssembleOceanDesc) NE_SZ(NeAssembleSeaIceDesc)
  NE_SZ(NeApplyRadiationDesc) NE_SZ(NeHostField) NE_SZ(NeHostStepDesc) NE_SZ(NeSeriesRingDesc) NE_SZ(NeFusedStepDesc) NE_SZ(NeDiagDesc) NE_SZ(NeElevationCorrectionDesc) NE_SZ(NeSeaIceAlbedo) NE_SZ(NeTabulatedAlbedo)
#undef NE_SZ
  return -1;
}

int ne_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    ne::cuda_error(e, "ne_device_count");
    cudaGetLastError();
    return 0;
  }
  return n;
}

int ne_memcpy_h2d(void* dst, const void* src, uint64_t bytes, void* stream) { NE_NVTX();
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  return e == cudaSuccess ? NE_OK : ne::cuda_error(e, "ne_memcpy_h2d");
}

int ne_memcpy_d2h(void* dst, const void* src, uint64_t bytes, void* stream) { NE_NVTX();
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  return e == cudaSuccess ? NE_OK : ne::cuda_error(e, "ne_memcpy_d2h");
}

// ---- the diagnostics all-reduce: NCCL bound at run time -----------------------------------------------------------
namespace {
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
struct NcclBinding { void* handle = nullptr; nccl_allreduce_fn all_reduce = nullptr; nccl_errstr_fn err = nullptr; bool tried = false; };
NcclBinding& nccl_binding() {
  static NcclBinding b;
  static std::mutex m;
  std::lock_guard<std::mutex> lock(m);
  if (!b.tried) {
    b.tried = true;
    const char* name = std::getenv("NE_B200_NCCL_LIB");
    const char* names[] = {name, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !n[0]) continue;
      // a host that already carries NCCL (NCCL.jl, torch) has it mapped: RTLD_NOLOAD finds that copy first
      b.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
      if (!b.handle) b.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (b.handle) break;
    }
    if (b.handle) {
      b.all_reduce = (nccl_allreduce_fn)dlsym(b.handle, "ncclAllReduce");
      b.err = (nccl_errstr_fn)dlsym(b.handle, "ncclGetErrorString");
    }
  }
  return b;
}
}  // namespace

int ne_diag_allreduce_f64(void* nccl_comm, double* sums, int32_t n, void* stream) { NE_NVTX();
  NE_REQUIRE(nccl_comm != nullptr && sums != nullptr && n > 0, "diag_allreduce: null communicator / buffer or n <= 0");
  NcclBinding& b = nccl_binding();
  if (!b.all_reduce) {
    ne::set_error("diag_allreduce: NCCL is not available in this process (dlopen libnccl.so.2 failed; set NE_B200_NCCL_LIB)");
    return NE_E_NO_VARIANT;
  }
  // ncclFloat64 = 8, ncclSum = 0 (nccl.h; stable since NCCL 2.0)
  const int rc = b.all_reduce(sums, sums, (size_t)n, 8, 0, nccl_comm, (cudaStream_t)stream);
  if (rc != 0) {
    ne::set_error("diag_allreduce: ncclAllReduce failed: %s", b.err ? b.err(rc) : "unknown NCCL error");
    return NE_E_CUDA;
  }
  return NE_OK;
}

int ne_stream_synchronize(void* stream) {
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  return e == cudaSuccess ? NE_OK : ne::cuda_error(e, "ne_stream_synchronize");
}

int ne_measure_fp64_peak(double* tflops, double* sm_clock_mhz_estimate) {
  int dev = 0, sms = 0, khz = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return ne::cuda_error(e, "ne_measure_fp64_peak");
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int threads = 256, blocks = sms * 8, iters = 4096;
  double* out = nullptr;
  e = cudaMalloc(&out, sizeof(double) * (size_t)threads * blocks);
  if (e != cudaSuccess) return ne::cuda_error(e, "ne_measure_fp64_peak: cudaMalloc");
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(t0);
    ne::dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(t1);
    e = cudaEventSynchronize(t1);
    if (e != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    double flops = 2.0 * 8 * 16 * (double)iters * threads * blocks;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(out);
  if (e != cudaSuccess) return ne::cuda_error(e, "ne_measure_fp64_peak");
  if (tflops) *tflops = best;
  if (sm_clock_mhz_estimate) *sm_clock_mhz_estimate = khz / 1000.0;
  return NE_OK;
}

}  // extern "C"
