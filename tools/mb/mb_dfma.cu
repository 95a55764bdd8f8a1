// Microbenchmark: FP64 pipe of one SM sub-partition (development tool, not product code).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mb/mb_dfma tools/mb/mb_dfma.cu
// Dependent-DFMA latency, and the DFMA rate as a function of warps per scheduler x independent chains per warp x
// integer instructions interleaved per DFMA (does the other work of the solve ride for free under the FP64 pipe?).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int NINT>
__global__ void k(double* out, int iters, double a, double b, int seed) {
  double x[ILP];
  int z[4] = {seed, seed + 1, seed + 2, seed + 3};
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        x[i] = __fma_rn(x[i], a, b);
#pragma unroll
        for (int q = 0; q < NINT; ++q) z[(i + q) & 3] = z[(i + q) & 3] * 3 + (z[(i + q + 1) & 3] ^ 5);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (z[0] ^ z[1] ^ z[2] ^ z[3]);
}

template <int ILP, int NINT>
void run(int warps_per_smsp, double* out, int clock_khz) {
  const int iters = 4000;
  dim3 grid(148), block(warps_per_smsp * 4 * 32);
  k<ILP, NINT><<<grid, block>>>(out, 10, 1.0000001, 1e-9, 1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ILP, NINT><<<grid, block>>>(out, iters, 1.0000001, 1e-9, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cycles = ms * 1e-3 * clock_khz * 1e3;
  const double dfma_per_smsp = (double)iters * 8 * ILP * warps_per_smsp;
  printf("{\"warps_per_smsp\": %d, \"ilp\": %d, \"int_per_dfma\": %d, \"cycles_per_warp_dfma_per_smsp\": %.3f, \"tflops\": %.2f}\n",
         warps_per_smsp, ILP, NINT, cycles / dfma_per_smsp, dfma_per_smsp * 148 * 4 * 32 * 2 / (ms * 1e-3) / 1e12);
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("{\"clock_khz\": %d}\n", khz);
  for (int w : {1, 2, 4, 6, 8}) {
    run<1, 0>(w, out, khz); run<2, 0>(w, out, khz); run<4, 0>(w, out, khz);
  }
  for (int w : {4, 6}) {
    run<1, 1>(w, out, khz); run<2, 1>(w, out, khz); run<1, 2>(w, out, khz); run<2, 2>(w, out, khz);
  }
  return 0;
}
