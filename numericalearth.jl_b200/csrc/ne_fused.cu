// ne_fused.cu — fused interface step (interpolation -> a–o solve -> assembly -> radiation), one host
// call, no synchronisation.  For the default plugin tree in Float64 phases 1-2 (both interpolations
// and the solve) run as ONE kernel (ao_fused_tab_kernel, ne_flux_kernels.cu) that never reads the
// interpolated state back from HBM and skips every output left NULL; any other configuration enqueues
// the component kernels back-to-back.  Phase 3-4 (net flux assembly, radiation) need the (i-1, j-1)
// neighbours of the just-computed stresses: they run as ONE further HBM-bound kernel (post_solve_kernel,
// ne_surface_kernels.cu) that also accumulates the optional diagnostics sums.
#include "ne_common.cuh"

namespace ne {
int fused_interp_ao_f64(const NeInterpDesc* atm, const NeInterpDesc* rad, const NeAtmosOceanDesc* d, void* stream);
int post_solve_f64(const NeFusedStepDesc* d, void* stream);   // ne_surface_kernels.cu
int post_solve_f32(const NeFusedStepDesc* d, void* stream);
}

extern "C" {

static int interp_and_ao(const NeFusedStepDesc* d, void* stream, bool f64) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  int rc = f64 ? ne::fused_interp_ao_f64(&d->atmosphere, &d->radiation, &d->ao, stream) : 1;
  if (rc < 0) return rc;
  if (rc > 0) {   // not eligible for the single-pass kernel: component kernels
    if (d->radiation.n_fields > 0) {
      rc = f64 ? ne_interp_state_f64(&d->radiation, stream) : ne_interp_state_f32(&d->radiation, stream);
      if (rc) return rc;
    }
    rc = f64 ? ne_interp_state_f64(&d->atmosphere, stream) : ne_interp_state_f32(&d->atmosphere, stream);
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

int ne_fused_interface_step_f64(const NeFusedStepDesc* d, void* stream) { return fused_step(d, stream, true); }
int ne_fused_interface_step_f32(const NeFusedStepDesc* d, void* stream) { return fused_step(d, stream, false); }
int ne_interp_and_ao_fluxes_f64(const NeFusedStepDesc* d, void* stream) { return interp_and_ao(d, stream, true); }
int ne_interp_and_ao_fluxes_f32(const NeFusedStepDesc* d, void* stream) { return interp_and_ao(d, stream, false); }
}
