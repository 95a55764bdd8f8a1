// Microbenchmark: what bounds the interpolation kernel? (development tool, not product code)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mb/mb_interp tools/mb/mb_interp.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int NX = 4322, NY = 1682, SX = 4334;      // launch extent and row stride (C4)
constexpr int SSX = 646, SSY = 326, NT = 2;         // source plane with halos
constexpr int NF = 7;

struct Args {
  const float* fi; const float* fj; const float* src[NF]; double* out[NF]; double nt;
};

__device__ __forceinline__ void point(const Args& a, int64_t idx, float& w1, float& w3, float& w5, float& w7, int& omm, int& omp, int& opm, int& opp) {
  float f = __ldg(a.fi + idx), g = __ldg(a.fj + idx);
  int im = (int)f, jm = (int)g;
  float xi = f - truncf(f), eta = g - truncf(g);
  w1 = (1 - xi) * (1 - eta); w3 = (1 - xi) * eta; w5 = xi * (1 - eta); w7 = xi * eta;
  omm = 3 + im + (3 + jm) * SSX; omp = omm + SSX; opm = omm + 1; opp = omp + 1;
}

// A: pure stream: 2 float loads -> NF double stores
template <int NOUT>
__global__ void k_stream(Args a) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)NX * NY) return;
  int j = t / NX, i = t - (int64_t)j * NX;
  int64_t idx = 6 + i + (int64_t)(6 + j) * SX;
  float f = __ldg(a.fi + idx), g = __ldg(a.fj + idx);
#pragma unroll
  for (int k = 0; k < NOUT; ++k) a.out[k][idx] = (double)f * (k + 1) + g;
}

// B: per-field sequential gathers (the shipped structure), 1-D blocks
template <int UNROLL>
__global__ void k_gather_seq(Args a) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)NX * NY) return;
  int j = t / NX, i = t - (int64_t)j * NX;
  int64_t idx = 6 + i + (int64_t)(6 + j) * SX;
  float w1, w3, w5, w7; int omm, omp, opm, opp;
  point(a, idx, w1, w3, w5, w7, omm, omp, opm, opp);
  float nt = (float)a.nt;
#pragma unroll UNROLL
  for (int k = 0; k < NF; ++k) {
    const float* d1 = a.src[k]; const float* d2 = d1 + SSX * SSY;
    float p1 = w1 * __ldg(d1 + omm) + w3 * __ldg(d1 + omp) + w5 * __ldg(d1 + opm) + w7 * __ldg(d1 + opp);
    float p2 = w1 * __ldg(d2 + omm) + w3 * __ldg(d2 + omp) + w5 * __ldg(d2 + opm) + w7 * __ldg(d2 + opp);
    a.out[k][idx] = (double)(p2 * nt + p1 * (1 - nt));
  }
}

// C: gathers but only ONE store (is it the gathers or the stores?)
__global__ void k_gather_onestore(Args a) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)NX * NY) return;
  int j = t / NX, i = t - (int64_t)j * NX;
  int64_t idx = 6 + i + (int64_t)(6 + j) * SX;
  float w1, w3, w5, w7; int omm, omp, opm, opp;
  point(a, idx, w1, w3, w5, w7, omm, omp, opm, opp);
  float nt = (float)a.nt, acc = 0;
#pragma unroll
  for (int k = 0; k < NF; ++k) {
    const float* d1 = a.src[k]; const float* d2 = d1 + SSX * SSY;
    float p1 = w1 * __ldg(d1 + omm) + w3 * __ldg(d1 + omp) + w5 * __ldg(d1 + opm) + w7 * __ldg(d1 + opp);
    float p2 = w1 * __ldg(d2 + omm) + w3 * __ldg(d2 + omp) + w5 * __ldg(d2 + opm) + w7 * __ldg(d2 + opp);
    acc += p2 * nt + p1 * (1 - nt);
  }
  a.out[0][idx] = acc;
}

// D: shared-memory staged source window: block = 256 consecutive x points of one row
__global__ void k_gather_smem(Args a) {
  __shared__ float win[NF * 2 * 2 * 48];   // per field: 2 times x 2 rows x up to 48 columns
  __shared__ int s_lo, s_jm;
  const int tiles_x = (NX + 255) / 256;
  int j = blockIdx.x / tiles_x, i = (blockIdx.x % tiles_x) * 256 + threadIdx.x;
  bool in = i < NX;
  int64_t idx = 6 + (in ? i : NX - 1) + (int64_t)(6 + j) * SX;
  float f = __ldg(a.fi + idx), g = __ldg(a.fj + idx);
  int im = (int)f, jm = (int)g;
  if (threadIdx.x == 0) { s_lo = im; s_jm = jm; }
  __syncthreads();
  const int lo = s_lo, jm0 = s_jm;
  // stage NF fields x 2 times x 2 rows x 48 columns
  for (int e = threadIdx.x; e < NF * 2 * 2 * 48; e += 256) {
    int c = e % 48, r = (e / 48) % 2, tt = (e / 96) % 2, k = e / 192;
    win[e] = __ldg(a.src[k] + tt * SSX * SSY + 3 + lo + c + (3 + jm0 + r) * SSX);
  }
  __syncthreads();
  if (!in) return;
  float xi = f - truncf(f), eta = g - truncf(g);
  float w1 = (1 - xi) * (1 - eta), w3 = (1 - xi) * eta, w5 = xi * (1 - eta), w7 = xi * eta;
  int c = im - lo;
  float nt = (float)a.nt;
#pragma unroll
  for (int k = 0; k < NF; ++k) {
    const float* w = win + k * 192;
    float p1 = w1 * w[c] + w3 * w[48 + c] + w5 * w[c + 1] + w7 * w[48 + c + 1];
    float p2 = w1 * w[96 + c] + w3 * w[144 + c] + w5 * w[96 + c + 1] + w7 * w[144 + c + 1];
    a.out[k][idx] = (double)(p2 * nt + p1 * (1 - nt));
  }
}

// E: 4 consecutive x points per thread, per-field sequential gathers, vector (2 x double2) stores when aligned
__global__ void k_gather_x4(Args a) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int qx = (NX + 3) / 4;
  if (t >= (int64_t)qx * NY) return;
  int j = t / qx, i0 = (t - (int64_t)j * qx) * 4;
  int64_t base = 6 + i0 + (int64_t)(6 + j) * SX;
  float w1[4], w3[4], w5[4], w7[4]; int omm[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int omp, opm, opp;
    if (i0 + q < NX) point(a, base + q, w1[q], w3[q], w5[q], w7[q], omm[q], omp, opm, opp);
    else { w1[q] = w3[q] = w5[q] = w7[q] = 0; omm[q] = 0; }
  }
  float nt = (float)a.nt;
#pragma unroll 1
  for (int k = 0; k < NF; ++k) {
    const float* d1 = a.src[k]; const float* d2 = d1 + SSX * SSY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int o = omm[q];
      float p1 = w1[q] * __ldg(d1 + o) + w3[q] * __ldg(d1 + o + SSX) + w5[q] * __ldg(d1 + o + 1) + w7[q] * __ldg(d1 + o + SSX + 1);
      float p2 = w1[q] * __ldg(d2 + o) + w3[q] * __ldg(d2 + o + SSX) + w5[q] * __ldg(d2 + o + 1) + w7[q] * __ldg(d2 + o + SSX + 1);
      if (i0 + q < NX) a.out[k][base + q] = (double)(p2 * nt + p1 * (1 - nt));
    }
  }
}

int main() {
  const size_t n_ex = (size_t)SX * (NY + 12);
  Args a;
  float *fi, *fj;
  cudaMalloc(&fi, n_ex * 4); cudaMalloc(&fj, n_ex * 4);
  std::vector<float> hfi(n_ex), hfj(n_ex);
  for (int j = 0; j < NY + 12; ++j)
    for (int i = 0; i < SX; ++i) {
      hfi[(size_t)j * SX + i] = fmodf((i - 6) * (640.0f / 4320.0f) + 0.3f + 640.f, 640.f);
      hfj[(size_t)j * SX + i] = 35.f + (j - 6) * (250.0f / 1682.0f);
    }
  cudaMemcpy(fi, hfi.data(), n_ex * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(fj, hfj.data(), n_ex * 4, cudaMemcpyHostToDevice);
  a.fi = fi; a.fj = fj; a.nt = 0.37;
  for (int k = 0; k < NF; ++k) {
    float* s; cudaMalloc(&s, (size_t)SSX * SSY * NT * 4); cudaMemset(s, 0, (size_t)SSX * SSY * NT * 4); a.src[k] = s;
    cudaMalloc(&a.out[k], n_ex * 8);
  }
  const int64_t n = (int64_t)NX * NY;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto launch, double bytes) {
    for (int w = 0; w < 3; ++w) launch();
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    cudaError_t err = cudaGetLastError();
    printf("%-40s %8.3f ms  %7.1f GB/s  %s\n", name, ms, bytes / ms / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
  };
  const double B7 = n * (8.0 + 56.0), B1 = n * (8.0 + 8.0);
  for (int bs : {128, 256, 512}) {
    unsigned nb = (unsigned)((n + bs - 1) / bs);
    char nm[64];
    snprintf(nm, 64, "A stream 7 stores, block %d", bs); run(nm, [&] { k_stream<7><<<nb, bs>>>(a); }, B7);
    snprintf(nm, 64, "A stream 1 store, block %d", bs); run(nm, [&] { k_stream<1><<<nb, bs>>>(a); }, B1);
    snprintf(nm, 64, "B gather seq unroll1, block %d", bs); run(nm, [&] { k_gather_seq<1><<<nb, bs>>>(a); }, B7);
    snprintf(nm, 64, "B gather seq unroll7, block %d", bs); run(nm, [&] { k_gather_seq<7><<<nb, bs>>>(a); }, B7);
    snprintf(nm, 64, "C gather one store, block %d", bs); run(nm, [&] { k_gather_onestore<<<nb, bs>>>(a); }, B1);
    unsigned nb4 = (unsigned)(((int64_t)((NX + 3) / 4) * NY + bs - 1) / bs);
    snprintf(nm, 64, "E gather x4/thread, block %d", bs); run(nm, [&] { k_gather_x4<<<nb4, bs>>>(a); }, B7);
  }
  run("D smem staged, block 256", [&] { k_gather_smem<<<(unsigned)(((NX + 255) / 256) * NY), 256>>>(a); }, B7);
  return 0;
}
