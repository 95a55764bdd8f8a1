// ne_flux_generic_asi_f32.cu — explicit instantiations of the generic flux kernels (see ne_flux_generic.cuh).
#include "ne_flux_generic.cuh"

namespace ne {
template int launch_asi<float, double, double>(const NeAtmosSeaIceDesc&, cudaStream_t);
template int launch_asi<float, double, float>(const NeAtmosSeaIceDesc&, cudaStream_t);
template int launch_asi<float, float, double>(const NeAtmosSeaIceDesc&, cudaStream_t);
template int launch_asi<float, float, float>(const NeAtmosSeaIceDesc&, cudaStream_t);
}  // namespace ne
