// Microbenchmark: L1TEX data-pipe cost (wavefronts) of shared-memory loads by width and address pattern on sm_100a
// (development tool, not product code).  One CTA of 256 threads per SM keeps issuing dependent-free LDS of one kind;
// cycles per warp-level LDS = SM-level issue interval of the data pipe.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mb/mb_lds tools/mb/mb_lds.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int SMEM_DOUBLES = 4096;   // 32 KB
constexpr int ITERS = 4096;

// pattern: 0 uniform address; 1 two addresses (lane parity); 2 eight addresses (lane / 4); 3 all distinct, stride 16 B * 15 (240 B records);
//          4 all distinct contiguous 16 B; 5 random records (hash) stride 240 B; 6 random 16 B-granular within 1 KB (the log table)
__device__ __forceinline__ int pattern_offset(int pattern, int lane, int k) {
  switch (pattern) {
    case 0: return 0;
    case 1: return (lane & 1) * 14;
    case 2: return (lane >> 2) * 30;
    case 3: return lane * 30;
    case 4: return lane * 2;
    case 5: return ((lane * 2654435761u + k * 40503u) >> 7) % 110 * 30;
    default: return (((lane * 2654435761u + k * 40503u) >> 9) & 63) * 2;
  }
}

template <int WIDTH>   // bytes per lane: 4, 8, 16
__global__ void __launch_bounds__(256) k_lds(int pattern, double* out, long long* cycles) {
  __shared__ __align__(16) double tab[SMEM_DOUBLES];
  for (int i = threadIdx.x; i < SMEM_DOUBLES; i += 256) tab[i] = i * 1e-3;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double acc0 = 0, acc1 = 0;
  int off[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) off[k] = pattern_offset(pattern, lane, k);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int o = off[k] + 2 * (it & 7);
      if (WIDTH == 16) { const double2 v = *reinterpret_cast<const double2*>(tab + o); acc0 += v.x; acc1 += v.y; }
      else if (WIDTH == 8) { acc0 += tab[o]; }
      else { acc0 += (double)reinterpret_cast<const float*>(tab)[2 * o]; }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 256 + threadIdx.x] = acc0 + acc1;
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 256 * sms);
  cudaMallocManaged(&cyc, sizeof(long long) * sms);
  const char* names[] = {"uniform", "2 addresses", "8 addresses (4 lanes each)", "32 records stride 240 B", "32 contiguous 16 B",
                         "random records stride 240 B", "random 16 B entries in 1 KB"};
  printf("{\n");
  for (int width : {4, 8, 16}) {
    for (int p = 0; p < 7; ++p) {
      for (int rep = 0; rep < 2; ++rep) {
        if (width == 16) k_lds<16><<<sms, 256>>>(p, out, cyc);
        else if (width == 8) k_lds<8><<<sms, 256>>>(p, out, cyc);
        else k_lds<4><<<sms, 256>>>(p, out, cyc);
        cudaDeviceSynchronize();
      }
      // 8 warps x ITERS x 8 loads per CTA (one CTA per SM)
      const double per_lds = (double)cyc[0] / (8.0 * ITERS * 8.0);
      printf("  \"LDS.%d %s\": %.2f,\n", width * 8, names[p], per_lds);
    }
  }
  printf("  \"unit\": \"SM cycles per warp-level load instruction (data-pipe wavefronts)\"\n}\n");
  return 0;
}
