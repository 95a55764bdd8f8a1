// ne_flux_generic_asi_f64.cu — explicit instantiations of the generic flux kernels (see ne_flux_generic.cuh).
#include "ne_flux_generic.cuh"

namespace ne {
template int launch_asi<double, double, double>(const NeAtmosSeaIceDesc&, cudaStream_t);
template int launch_asi<double, float, double>(const NeAtmosSeaIceDesc&, cudaStream_t);
}  // namespace ne
