// svgd_gram.cu — centred-Gram K1 (svgd_gram.cuh): tensor-map encoding and the n = 16 / 20 instantiations.
#include "svgd_gram.cuh"

#include <cudaTypedefs.h>

namespace bde {

// TMA tensor map over the rows of X[n, D] (row stride ld elements): dim0 = columns, dim1 = particles; box =
// box_cols x n.  cuTensorMapEncodeTiled is resolved through the runtime (no link-time dependency on libcuda).
int encode_rows_tensor_map(CUtensorMap* map, const float* X, int n, int64_t D, int64_t ld, int box_cols, int l2_promotion) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        BDE_RETURN_IF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) return BDE_ERR_INVALID_ARG;
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    if (!aligned16(X) || (ld % 4) != 0 || D < 1 || D > 0x7fffffffLL || n < 1 || n > 256 || box_cols > 256) return BDE_ERR_INVALID_ARG;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(n)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(float)};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(n)};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapL2promotion promo = l2_promotion == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : l2_promotion == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                         : l2_promotion == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                             : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? BDE_OK : BDE_ERR_INVALID_ARG;
}

template int launch_pairgram<16>(const float*, int64_t, int64_t, double*, void*, int, const BandwidthParams&, cudaStream_t);
template int launch_pairgram<20>(const float*, int64_t, int64_t, double*, void*, int, const BandwidthParams&, cudaStream_t);

}  // namespace bde
