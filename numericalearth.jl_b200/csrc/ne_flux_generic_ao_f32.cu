// ne_flux_generic_ao_f32.cu — explicit instantiations of the generic flux kernels (see ne_flux_generic.cuh).
#include "ne_flux_generic.cuh"

namespace ne {
template int launch_ao<float, double, double>(const NeAtmosOceanDesc&, cudaStream_t);
template int launch_ao<float, double, float>(const NeAtmosOceanDesc&, cudaStream_t);
template int launch_ao<float, float, double>(const NeAtmosOceanDesc&, cudaStream_t);
template int launch_ao<float, float, float>(const NeAtmosOceanDesc&, cudaStream_t);
}  // namespace ne
