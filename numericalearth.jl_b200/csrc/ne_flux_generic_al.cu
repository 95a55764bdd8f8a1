// ne_flux_generic_al.cu — atmosphere–land instantiations of the generic flux kernel, Float64 models.
#include "ne_flux_generic.cuh"

namespace ne {
template int launch_al<double, double, double>(const NeAtmosLandDesc&, cudaStream_t);
template int launch_al<double, float, double>(const NeAtmosLandDesc&, cudaStream_t);
}  // namespace ne
