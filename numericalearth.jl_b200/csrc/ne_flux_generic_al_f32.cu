// ne_flux_generic_al_f32.cu — atmosphere–land instantiations of the generic flux kernel, Float32 models.
#include "ne_flux_generic.cuh"

namespace ne {
template int launch_al<float, float, double>(const NeAtmosLandDesc&, cudaStream_t);
template int launch_al<float, float, float>(const NeAtmosLandDesc&, cudaStream_t);
template int launch_al<float, double, double>(const NeAtmosLandDesc&, cudaStream_t);
template int launch_al<float, double, float>(const NeAtmosLandDesc&, cudaStream_t);
}  // namespace ne
