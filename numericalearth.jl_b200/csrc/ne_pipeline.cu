// ne_pipeline.cu — host-buffer interface step, pipelined natively: the runtime around the kernels for a host
// ocean (or an ocean behind a host bounce buffer) whose surface state arrives in pinned HOST memory every
// coupled step (the call sequence of time_step_earth_system_model.jl:38-83 behind one C-ABI entry).
//
// The exchange grid is cut into latitude chunks.  Chunk c's parent rows of every ocean surface field go
// host->device on the pipeline's COPY stream; the caller's COMPUTE stream waits for chunk c's event and runs
// the a–o solve and the post-solve kernel (net-flux assembly + radiation) on the launch rows whose inputs
// (including the (j+1) row of the v stencil, atmosphere_ocean_fluxes.jl:64-65) have landed — the descriptors'
// launch range restricts each call to its rows.  The atmosphere interpolation does not depend on the ocean and
// runs up front, behind the first copy.  PCIe time and kernel time overlap instead of adding up; results are
// those of the unchunked step bit for bit (pointwise kernels; the (i-1, j-1) stencils of the assembly only
// reach back into rows an earlier band computed).  Everything is enqueued; nothing synchronises.
#include <cstdlib>
#include <vector>

#include "ne_common.cuh"

namespace ne {
int post_solve_f64(const NeFusedStepDesc* d, void* stream);   // ne_surface_kernels.cu
int post_solve_f32(const NeFusedStepDesc* d, void* stream);
int interp_phase(const NeFusedStepDesc* d, void* stream, bool f64);   // ne_fused.cu

struct HostPipeline {
  int device;
  cudaStream_t copy;
  std::vector<cudaEvent_t> landed;   // chunk c is on the device
  cudaEvent_t done;                  // the kernels of the previous step have read the device arrays
  bool first;
  // results back to the host (NeHostStepDesc.out_fields): band c's rows leave on their own stream as soon as the
  // post-solve kernel of band c has written them, behind the kernels of band c + 1 and the H2D copies (PCIe is full duplex)
  cudaStream_t back;
  std::vector<cudaEvent_t> computed;
  cudaEvent_t drained;
};

#define NE_CUDA_TRY(expr, where)                             \
  do {                                                       \
    cudaError_t e__ = (expr);                                \
    if (e__ != cudaSuccess) return ne::cuda_error(e__, where); \
  } while (0)

static int pipelined_step(HostPipeline* hp, const NeHostStepDesc* d, void* stream, bool f64) {
  NE_REQUIRE(hp != nullptr && d != nullptr, "host pipeline: null handle or descriptor");
  NE_REQUIRE(d->n_fields >= 0 && d->n_fields <= NE_HOST_MAX_FIELDS, "host pipeline: n_fields out of range");
  NE_REQUIRE(d->n_chunks >= 1 && d->n_chunks <= (int)hp->landed.size(), "host pipeline: n_chunks exceeds the handle's capacity");
  NE_REQUIRE(d->row_bytes > 0, "host pipeline: row_bytes must be positive");
  NE_REQUIRE(d->n_out_fields >= 0 && d->n_out_fields <= NE_HOST_MAX_FIELDS, "host pipeline: n_out_fields out of range");
  for (int f = 0; f < d->n_out_fields; ++f)
    NE_REQUIRE(d->out_fields[f].host && d->out_fields[f].device, "host pipeline: null output field pointer");
  const NeExchangeGrid& g = d->step.ao.grid;
  NE_REQUIRE(g.ny >= 1 && g.hy >= 1, "host pipeline: needs a one-row halo");
  cudaStream_t compute = (cudaStream_t)stream;
  const int64_t rows_total = g.ny + 2 * g.hy;
  const int nc = (int)std::min<int64_t>(d->n_chunks, g.ny);
  // the copies of this step must not overtake the kernels of the previous one that still read the device arrays
  if (!hp->first) NE_CUDA_TRY(cudaStreamWaitEvent(hp->copy, hp->done, 0), "host pipeline (wait done)");
  hp->first = false;
  // development knobs: time the two sides of the pipeline alone
  const char* kc = std::getenv("NE_B200_PIPE_NO_COPY");
  const char* kk = std::getenv("NE_B200_PIPE_NO_COMPUTE");
  const bool no_copy = kc && kc[0] == '1', no_compute = kk && kk[0] == '1';
  std::vector<int64_t> edge(nc + 1);
  // chunk edges: equal shares, except that the last three chunks taper (8 %, 5 %, 3 % of the rows): what remains
  // to compute after the last byte has landed is the last band only
  {
    std::vector<double> share(nc, 1.0 / nc);
    if (nc >= 6) {
      const double tail[3] = {0.08, 0.05, 0.03};
      for (int c = 0; c < nc - 3; ++c) share[c] = (1.0 - 0.16) / (nc - 3);
      for (int k = 0; k < 3; ++k) share[nc - 3 + k] = tail[k];
    }
    double acc = 0;
    edge[0] = 0;
    for (int c = 0; c < nc; ++c) {
      acc += share[c];
      edge[c + 1] = c == nc - 1 ? rows_total : std::min<int64_t>(rows_total, (int64_t)(acc * rows_total + 0.5));
    }
  }
  for (int c = 0; c < nc; ++c) {
    const int64_t off = edge[c] * d->row_bytes, nbytes = (edge[c + 1] - edge[c]) * d->row_bytes;
    for (int f = 0; f < d->n_fields && !no_copy; ++f) {
      NE_REQUIRE(d->fields[f].host && d->fields[f].device, "host pipeline: null field pointer");
      NE_CUDA_TRY(cudaMemcpyAsync((char*)d->fields[f].device + off, (const char*)d->fields[f].host + off, (size_t)nbytes,
                                  cudaMemcpyHostToDevice, hp->copy), "host pipeline (H2D)");
    }
    NE_CUDA_TRY(cudaEventRecord(hp->landed[c], hp->copy), "host pipeline (record)");
  }
  int rc;
  if (no_compute) {
    for (int c = 0; c < nc; ++c) NE_CUDA_TRY(cudaStreamWaitEvent(compute, hp->landed[c], 0), "host pipeline (wait chunk)");
    NE_CUDA_TRY(cudaEventRecord(hp->done, compute), "host pipeline (record done)");
    return NE_OK;
  }
  rc = interp_phase(&d->step, stream, f64);
  if (rc) return rc;

  NeFusedStepDesc band = d->step;   // launch ranges rewritten per band
  band.diag.n_fields = 0;            // the diagnostics sums run once, over the whole grid, after the last band
  const int64_t jl = d->step.ao.grid.j_lo, jh = d->step.ao.grid.j_hi;
  int64_t j_next = jl;
  for (int c = 0; c < nc; ++c) {
    NE_CUDA_TRY(cudaStreamWaitEvent(compute, hp->landed[c], 0), "host pipeline (wait chunk)");
    // last launch row whose (j+1) parent row (index j + hy) lies inside the rows copied so far
    int64_t b = (c == nc - 1) ? jh : std::min<int64_t>((edge[c + 1] - 1) - g.hy, jh);
    if (b < j_next) continue;
    band.ao.grid.j_lo = j_next; band.ao.grid.j_hi = b;
    rc = f64 ? ne_atmosphere_ocean_fluxes_f64(&band.ao, stream) : ne_atmosphere_ocean_fluxes_f32(&band.ao, stream);
    if (rc) return rc;
    const int64_t a0 = std::max<int64_t>(j_next, d->step.assemble.grid.j_lo), a1 = std::min<int64_t>(b, d->step.assemble.grid.j_hi);
    if (a1 >= a0) {
      band.assemble.grid.j_lo = a0; band.assemble.grid.j_hi = a1;
      band.apply_radiation.grid.j_lo = a0; band.apply_radiation.grid.j_hi = a1;
      rc = f64 ? post_solve_f64(&band, stream) : post_solve_f32(&band, stream);
      if (rc < 0) return rc;
      if (rc > 0) {   // descriptors do not line up for the one-kernel form: component kernels
        rc = f64 ? ne_assemble_net_ocean_fluxes_f64(&band.assemble, stream) : ne_assemble_net_ocean_fluxes_f32(&band.assemble, stream);
        if (rc) return rc;
        if (band.apply_radiation.radiation.enabled) {
          rc = f64 ? ne_apply_radiative_fluxes_f64(&band.apply_radiation, stream) : ne_apply_radiative_fluxes_f32(&band.apply_radiation, stream);
          if (rc) return rc;
        }
      }
      if (d->n_out_fields > 0) {   // parent rows of the launch rows a0..a1 (the first / last band also carries the halo rows)
        NE_CUDA_TRY(cudaEventRecord(hp->computed[c], compute), "host pipeline (record computed)");
        NE_CUDA_TRY(cudaStreamWaitEvent(hp->back, hp->computed[c], 0), "host pipeline (wait computed)");
        const int64_t r0 = (a0 == d->step.assemble.grid.j_lo) ? 0 : a0 + g.hy - 1;
        const int64_t r1 = (a1 == d->step.assemble.grid.j_hi) ? rows_total : a1 + g.hy;
        for (int f = 0; f < d->n_out_fields; ++f)
          NE_CUDA_TRY(cudaMemcpyAsync((char*)d->out_fields[f].host + r0 * d->row_bytes, (const char*)d->out_fields[f].device + r0 * d->row_bytes,
                                      (size_t)((r1 - r0) * d->row_bytes), cudaMemcpyDeviceToHost, hp->back), "host pipeline (D2H)");
      }
    }
    j_next = b + 1;
  }
  if (d->step.diag.n_fields > 0) {
    rc = f64 ? ne_diag_reduce_f64(&d->step.diag, stream) : ne_diag_reduce_f32(&d->step.diag, stream);
    if (rc) return rc;
  }
  if (d->n_out_fields > 0) {   // whoever synchronises the compute stream also waits for the results to be on the host
    NE_CUDA_TRY(cudaEventRecord(hp->drained, hp->back), "host pipeline (record drained)");
    NE_CUDA_TRY(cudaStreamWaitEvent(compute, hp->drained, 0), "host pipeline (wait drained)");
  }
  NE_CUDA_TRY(cudaEventRecord(hp->done, compute), "host pipeline (record done)");
  return NE_OK;
}

}  // namespace ne

extern "C" {

int ne_host_pipeline_create(void** handle, int32_t max_chunks) {
  NE_REQUIRE(handle != nullptr && max_chunks >= 1 && max_chunks <= 4096, "host pipeline: bad arguments");
  ne::HostPipeline* hp = new ne::HostPipeline();
  hp->first = true;
  hp->copy = nullptr;
  hp->back = nullptr;
  hp->done = nullptr;
  hp->drained = nullptr;
  cudaError_t e = cudaGetDevice(&hp->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->copy, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->back, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&hp->done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&hp->drained, cudaEventDisableTiming);
  for (int c = 0; c < max_chunks && e == cudaSuccess; ++c) {
    cudaEvent_t ev;
    e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) hp->landed.push_back(ev);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) hp->computed.push_back(ev);
  }
  if (e != cudaSuccess) {
    const int rc = ne::cuda_error(e, "ne_host_pipeline_create");
    for (cudaEvent_t ev : hp->landed) cudaEventDestroy(ev);
    for (cudaEvent_t ev : hp->computed) cudaEventDestroy(ev);
    if (hp->done) cudaEventDestroy(hp->done);
    if (hp->drained) cudaEventDestroy(hp->drained);
    if (hp->back) cudaStreamDestroy(hp->back);
    if (hp->copy) cudaStreamDestroy(hp->copy);
    delete hp;
    cudaGetLastError();
    return rc;
  }
  *handle = hp;
  return NE_OK;
}

int ne_host_pipeline_destroy(void* handle) {
  ne::HostPipeline* hp = (ne::HostPipeline*)handle;
  if (!hp) return NE_OK;
  for (cudaEvent_t ev : hp->landed) cudaEventDestroy(ev);
  for (cudaEvent_t ev : hp->computed) cudaEventDestroy(ev);
  cudaEventDestroy(hp->done);
  cudaEventDestroy(hp->drained);
  cudaStreamDestroy(hp->back);
  cudaStreamDestroy(hp->copy);
  delete hp;
  return NE_OK;
}

int ne_host_pipelined_step_f64(void* handle, const NeHostStepDesc* d, void* stream) { NE_NVTX();
  return ne::pipelined_step((ne::HostPipeline*)handle, d, stream, true);
}
int ne_host_pipelined_step_f32(void* handle, const NeHostStepDesc* d, void* stream) { NE_NVTX();
  return ne::pipelined_step((ne::HostPipeline*)handle, d, stream, false);
}

}
