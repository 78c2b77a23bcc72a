// explicit instantiations of the SVGD fast paths (split so the files compile in parallel)
#include "svgd_kernels.cuh"
namespace bde {
template int launch_pairdist<5>(const float*, int64_t, int64_t, double*, int, void*, int, const BandwidthParams&, cudaStream_t, int);
template int launch_apply<5>(const float*, const float*, float*, const float*, const float*, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
template int launch_apply_fused<5>(float*, const float*, const float*, const float*, int64_t, int64_t, int64_t, const BaseOptParams&, cudaStream_t, const NextDistParams*);
template int launch_pairdist<6>(const float*, int64_t, int64_t, double*, int, void*, int, const BandwidthParams&, cudaStream_t, int);
template int launch_apply<6>(const float*, const float*, float*, const float*, const float*, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
template int launch_apply_fused<6>(float*, const float*, const float*, const float*, int64_t, int64_t, int64_t, const BaseOptParams&, cudaStream_t, const NextDistParams*);
}  // namespace bde
