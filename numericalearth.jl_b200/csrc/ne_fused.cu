// ne_fused.cu — fused interface step (interpolation -> a–o solve -> assembly -> radiation).
// Round-1 first cut: enqueues the component kernels back-to-back on the caller's stream (one host
// call, no synchronisation).  The single-pass kernel that never materialises the interpolated
// atmosphere state replaces this body once the component kernels are parity-green.
#include "ne_common.cuh"

extern "C" {

static int fused_step(const NeFusedStepDesc* d, void* stream, bool f64) {
  NE_REQUIRE(d != nullptr, "null descriptor");
  int rc;
  if (d->radiation.n_fields > 0) {
    rc = f64 ? ne_interp_state_f64(&d->radiation, stream) : ne_interp_state_f32(&d->radiation, stream);
    if (rc) return rc;
  }
  rc = f64 ? ne_interp_state_f64(&d->atmosphere, stream) : ne_interp_state_f32(&d->atmosphere, stream);
  if (rc) return rc;
  rc = f64 ? ne_atmosphere_ocean_fluxes_f64(&d->ao, stream) : ne_atmosphere_ocean_fluxes_f32(&d->ao, stream);
  if (rc) return rc;
  rc = f64 ? ne_assemble_net_ocean_fluxes_f64(&d->assemble, stream) : ne_assemble_net_ocean_fluxes_f32(&d->assemble, stream);
  if (rc) return rc;
  if (d->apply_radiation.radiation.enabled) {
    rc = f64 ? ne_apply_radiative_fluxes_f64(&d->apply_radiation, stream) : ne_apply_radiative_fluxes_f32(&d->apply_radiation, stream);
    if (rc) return rc;
  }
  return NE_OK;
}

int ne_fused_interface_step_f64(const NeFusedStepDesc* d, void* stream) { return fused_step(d, stream, true); }
int ne_fused_interface_step_f32(const NeFusedStepDesc* d, void* stream) { return fused_step(d, stream, false); }
}
