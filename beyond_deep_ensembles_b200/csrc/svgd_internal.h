// svgd_internal.h — internal (non-ABI) entry points shared between svgd.cu and host_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>

namespace bde {

struct BandwidthParams {
    double l2_reg, kernel_grad_scale, dataset_size, h_override;
    float* K;
    float* A;
    double* info;
    int32_t* sel;
};

int pairdist_impl(const float* X, int n, int64_t D, int64_t ld, double* dist, int accumulate, void* ws,
                  size_t ws_bytes, int fuse, const BandwidthParams& bp, cudaStream_t st);
int apply_impl(const float* X, const float* G, float* out, const float* K, const float* A, int n, int64_t D,
               int64_t ldx, int64_t ldg, int64_t ldo, cudaStream_t st);

}  // namespace bde
