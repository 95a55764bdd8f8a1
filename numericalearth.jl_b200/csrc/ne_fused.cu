// ne_fused.cu — fused interface step (interpolation -> a–o solve -> assembly -> radiation), one host
// call, no synchronisation.  For the default plugin tree in Float64 phases 1-2 (both interpolations
// and the solve) run as ONE kernel (ao_fused_tab_kernel, ne_flux_kernels.cu) that never reads the
// interpolated state back from HBM and skips every output left NULL; any other configuration enqueues
// the component kernels back-to-back.  Phase 3-4 (net flux assembly, radiation) need the (i-1, j-1)
// neighbours of the just-computed stresses: they run as ONE further HBM-bound kernel (post_solve_kernel,
// ne_surface_kernels.cu) that also accumulates the optional diagnostics sums.
#include <cstdlib>
#include <cstring>

#include "ne_common.cuh"

namespace ne {
int fused_interp_ao_f64(const NeInterpDesc* atm, const NeInterpDesc* rad, const NeAtmosOceanDesc* d, void* stream);
int post_solve_f64(const NeFusedStepDesc* d, void* stream);   // ne_surface_kernels.cu
int post_solve_f32(const NeFusedStepDesc* d, void* stream);
}

// The radiation series join the atmosphere's interpolation launch when both descriptors describe the same
// interpolation: same launch range, same fractional-index arrays (a binding that regrids both components from one
// source grid stores them once), same source extents / element type and the same time interpolator.  Same arithmetic
// per series, so the merged launch is bit-identical to the two separate ones; it saves one pass over the fractional
// indices, one set of interpolators per point and one launch.
static bool merge_interp(const NeInterpDesc& a, const NeInterpDesc& r, NeInterpDesc& m) {
  const char* off = std::getenv("NE_B200_NO_INTERP_MERGE");
  if (off && off[0] == '1') return false;
  if (r.n_fields <= 0 || a.n_fields + r.n_fields > 9) return false;
  if (std::memcmp(&a.grid, &r.grid, sizeof(NeExchangeGrid)) != 0) return false;
  if (a.frac_i != r.frac_i || a.frac_j != r.frac_j || a.src_dtype != r.src_dtype) return false;
  if (a.src_nx != r.src_nx || a.src_ny != r.src_ny || a.src_hx != r.src_hx || a.src_hy != r.src_hy || a.src_nt != r.src_nt) return false;
  if (a.time.frac != r.time.frac || a.time.frac_dtype != r.time.frac_dtype || a.time.m1 != r.time.m1 ||
      a.time.m2 != r.time.m2 || a.time.same != r.time.same) return false;
  if (r.potential || r.rotation_cos || r.rotation_sin) return false;
  m = a;
  for (int f = 0; f < r.n_fields; ++f) {
    const int k = a.n_fields + f;
    m.n_summands[k] = r.n_summands[f];
    for (int q = 0; q < NE_MAX_SUMMANDS; ++q) m.series[k][q] = r.series[f][q];
    m.out[k] = r.out[f];
  }
  m.n_fields = a.n_fields + r.n_fields;
  return true;
}

namespace ne {
// phase 1 of the step: both interpolations, as one launch when they can be merged
int interp_phase(const NeFusedStepDesc* d, void* stream, bool f64) {
  NeInterpDesc merged;
  if (merge_interp(d->atmosphere, d->radiation, merged))
    return f64 ? ne_interp_state_f64(&merged, stream) : ne_interp_state_f32(&merged, stream);
  if (d->radiation.n_fields > 0) {
    const int rc = f64 ? ne_interp_state_f64(&d->radiation, stream) : ne_interp_state_f32(&d->radiation, stream);
    if (rc) return rc;
  }
  return f64 ? ne_interp_state_f64(&d->atmosphere, stream) : ne_interp_state_f32(&d->atmosphere, stream);
}
}  // namespace ne

extern "C" {

static int interp_and_ao(const NeFusedStepDesc* d, void* stream, bool f64) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  int rc = f64 ? ne::fused_interp_ao_f64(&d->atmosphere, &d->radiation, &d->ao, stream) : 1;
  if (rc < 0) return rc;
  if (rc > 0) {   // not eligible for the single-pass kernel: component kernels
    rc = ne::interp_phase(d, stream, f64);
    if (rc) return rc;
    rc = f64 ? ne_atmosphere_ocean_fluxes_f64(&d->ao, stream) : ne_atmosphere_ocean_fluxes_f32(&d->ao, stream);
    if (rc) return rc;
  }
  return NE_OK;
}

static int fused_step(const NeFusedStepDesc* d, void* stream, bool f64) {
  int rc = interp_and_ao(d, stream, f64);
  if (rc) return rc;
  rc = f64 ? ne::post_solve_f64(d, stream) : ne::post_solve_f32(d, stream);   // one kernel when the descriptors line up
  if (rc <= 0) return rc;
  rc = f64 ? ne_assemble_net_ocean_fluxes_f64(&d->assemble, stream) : ne_assemble_net_ocean_fluxes_f32(&d->assemble, stream);
  if (rc) return rc;
  if (d->apply_radiation.radiation.enabled) {
    rc = f64 ? ne_apply_radiative_fluxes_f64(&d->apply_radiation, stream) : ne_apply_radiative_fluxes_f32(&d->apply_radiation, stream);
    if (rc) return rc;
  }
  if (d->diag.n_fields > 0) return f64 ? ne_diag_reduce_f64(&d->diag, stream) : ne_diag_reduce_f32(&d->diag, stream);
  return NE_OK;
}

int ne_fused_interface_step_f64(const NeFusedStepDesc* d, void* stream) { NE_NVTX(); return fused_step(d, stream, true); }
int ne_fused_interface_step_f32(const NeFusedStepDesc* d, void* stream) { NE_NVTX(); return fused_step(d, stream, false); }
int ne_interp_and_ao_fluxes_f64(const NeFusedStepDesc* d, void* stream) { NE_NVTX(); return interp_and_ao(d, stream, true); }
int ne_interp_and_ao_fluxes_f32(const NeFusedStepDesc* d, void* stream) { NE_NVTX(); return interp_and_ao(d, stream, false); }
}
