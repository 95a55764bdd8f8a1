// fastmath_gpu_check.cu — DEVICE-side accuracy check of ne_fastmath.cuh: the functions as the kernels run them (MUFU seeds of
// the real hardware, not the host emulation of tools/fastmath_check.cu) against long double.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/fastmath_gpu_check tools/fastmath_gpu_check.cu
// Prints one JSON object: max relative error in units of 2^-53 (exp_lo: plain relative error).
#include <cstdio>
#include <random>
#include <vector>

#include "../numericalearth.jl_b200/csrc/ne_flux_tab.cuh"

using namespace ne;

enum { F_RCP, F_RCP3, F_SQRT, F_SQRT3, F_CBRT, F_CBRT3, F_LOG, F_LOG_REP, F_EXP, F_EXP_LO, F_COUNT };

__global__ void eval_kernel(int fn, const double* x, double* y, int n, const double* tab, const double* lrep, fm::MathConsts mc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fm::OpsPlain o;
  const double v = x[i];
  double r = 0;
  switch (fn) {
    case F_RCP: r = fm::rcp(o, v); break;
    case F_RCP3: r = fm::rcp3(o, v); break;
    case F_SQRT: r = fm::sqrt_pos(o, v); break;
    case F_SQRT3: r = fm::sqrt3(o, v); break;
    case F_CBRT: r = fm::cbrt_pos(o, mc, v); break;
    case F_CBRT3: r = fm::cbrt3(o, mc, v); break;
    case F_LOG: r = fm::log_pos(o, tab, mc, v); break;
    case F_LOG_REP: r = fm::log_pos_rep(o, lrep, threadIdx.x & (fm::LOG_REP - 1), mc, v); break;
    case F_EXP: r = fm::exp_mid(o, mc, v); break;
    case F_EXP_LO: r = fm::exp_lo(o, tab, mc, v); break;
  }
  y[i] = r;
}

int main() {
  NeFluxFormulation f;
  std::memset(&f, 0, sizeof(f));
  double m[11] = {50, 0.35, 0.7, 0.75, 5 / 0.35, 15, 2, 3.141592653589793 / 2, 10.15, 3, 3.141592653589793 / std::sqrt(3.0)};
  double s[12] = {50, 0.35, 2.0 / 3, 1.5, 14.28, 8.525, 15, 2, 0, 34.15, 3, 3.141592653589793 / std::sqrt(3.0)};
  f.psi_momentum.a.kind = NE_PSI_EDSON_MOMENTUM;
  f.psi_temperature.a.kind = f.psi_water_vapor.a.kind = NE_PSI_EDSON_SCALAR;
  for (int k = 0; k < 11; ++k) f.psi_momentum.a.p[k] = m[k];
  for (int k = 0; k < 12; ++k) { f.psi_temperature.a.p[k] = s[k]; f.psi_water_vapor.a.p[k] = s[k]; }
  f.subgrid_velocities.gustiness_parameter = 1.2;
  f.subgrid_velocities.minimum_gustiness = 0.01;
  static double tab[fm::TAB_SIZE];
  TabParams T;
  build_solver_tables(f, tab, T);
  std::vector<double> lrep(2 * fm::LOG_N * fm::LOG_REP);
  for (int k = 0; k < fm::LOG_N * fm::LOG_REP; ++k) { lrep[2 * k] = tab[fm::TAB_LOG + 2 * (k / fm::LOG_REP)]; lrep[2 * k + 1] = tab[fm::TAB_LOG + 2 * (k / fm::LOG_REP) + 1]; }
  const int n = 1 << 20;
  double *dx, *dy, *dtab, *dlrep;
  cudaMalloc(&dx, n * 8); cudaMalloc(&dy, n * 8); cudaMalloc(&dtab, sizeof(tab)); cudaMalloc(&dlrep, lrep.size() * 8);
  cudaMemcpy(dtab, tab, sizeof(tab), cudaMemcpyHostToDevice);
  cudaMemcpy(dlrep, lrep.data(), lrep.size() * 8, cudaMemcpyHostToDevice);
  std::mt19937_64 rng(20261017);
  std::uniform_real_distribution<double> u(0, 1);
  const char* names[F_COUNT] = {"rcp_ulp", "rcp3_ulp", "sqrt_ulp", "sqrt3_ulp", "cbrt_ulp", "cbrt3_ulp", "log_ulp", "log_rep_ulp", "exp_ulp", "exp_lo_rel"};
  std::vector<double> x(n), y(n);
  printf("{");
  for (int fn = 0; fn < F_COUNT; ++fn) {
    for (int i = 0; i < n; ++i) {
      const double t = u(rng);
      if (fn == F_EXP || fn == F_EXP_LO) x[i] = -700 + 740 * t;
      else if (fn == F_LOG || fn == F_LOG_REP) x[i] = (i & 1) ? std::exp(-27 + 26.3 * t) : std::exp(0.7 + 27 * t);   // away from 1 (relative error)
      else x[i] = std::exp(std::log(1e-12) + t * (std::log(1e12) - std::log(1e-12)));
    }
    cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice);
    eval_kernel<<<(n + 255) / 256, 256>>>(fn, dx, dy, n, dtab, dlrep, T.mc);
    if (cudaMemcpy(y.data(), dy, n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("\"error\": \"cuda\"}\n"); return 1; }
    double worst = 0;
    for (int i = 0; i < n; ++i) {
      const long double v = x[i];
      long double ref = 0;
      switch (fn) {
        case F_RCP: case F_RCP3: ref = 1 / v; break;
        case F_SQRT: case F_SQRT3: ref = sqrtl(v); break;
        case F_CBRT: case F_CBRT3: ref = cbrtl(v); break;
        case F_LOG: case F_LOG_REP: ref = logl(v); break;
        default: ref = expl(v);
      }
      double e = (double)(fabsl((long double)y[i] - ref) / fabsl(ref));
      if (fn != F_EXP_LO) e *= 9007199254740992.0;
      if (!(e <= worst)) worst = e;
    }
    printf("%s\"%s\": %.4g", fn ? ", " : "", names[fn], worst);
  }
  printf("}\n");
  return 0;
}
