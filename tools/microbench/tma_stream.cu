// tma_stream.cu — microbenchmark (not product code): how fast can one persistent CTA per SM stream [rows x cols]
// tiles of a row-major [rows, D] fp32 matrix into a shared-memory ring, as a function of the tile width (bytes
// per row segment), the copy form (one cp.async.bulk per row | one cp.async.bulk.tensor.2d per tile) and the
// number of issuing lanes?  Consumers only wait for the tile and release it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu -lcuda && ./tma_stream
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// mode 0: one lane issues `rows` 1-D bulk copies per tile; mode 1: one lane issues one 2-D tensor copy per tile;
// mode 2: lanes 0..rows-1 of the producer warp issue one 1-D bulk copy each (rows <= 32)
__global__ void __launch_bounds__(64, 1)
stream_kernel(const float* __restrict__ X, const __grid_constant__ CUtensorMap tmap, int rows, int64_t D, int64_t ld,
              int tc, int stages, int mode, unsigned long long* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[32], empty_bar[32];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], mode == 2 ? rows : 1); mbar_init(&empty_bar[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t ntiles = D / tc;
    const uint32_t row_bytes = tc * 4u, stage_bytes = rows * row_bytes;
    if (tid < 32) {  // producer warp
        int it = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int s = it % stages;
            const uint32_t use = it / stages;
            unsigned char* dst = smem + (size_t)s * stage_bytes;
            if (mode == 2) {
                if (tid < rows) {
                    mbar_wait(&empty_bar[s], (use & 1u) ^ 1u);
                    mbar_expect(&full_bar[s], row_bytes);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + tid * row_bytes)), "l"(X + tid * ld + t * tc), "r"(row_bytes), "r"(smem_u32(&full_bar[s])) : "memory");
                }
            } else if (tid == 0) {
                mbar_wait(&empty_bar[s], (use & 1u) ^ 1u);
                mbar_expect(&full_bar[s], stage_bytes);
                if (mode == 0) {
                    for (int r = 0; r < rows; ++r)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + r * row_bytes)), "l"(X + r * ld + t * tc), "r"(row_bytes), "r"(smem_u32(&full_bar[s])) : "memory");
                } else {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"((int)(t * tc)), "r"(0), "r"(smem_u32(&full_bar[s])) : "memory");
                }
            }
        }
    } else if (tid == 32) {  // consumer: wait, touch one word, release
        int it = 0;
        unsigned long long acc = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int s = it % stages;
            const uint32_t use = it / stages;
            mbar_wait(&full_bar[s], use & 1u);
            acc += *reinterpret_cast<volatile unsigned int*>(smem + (size_t)s * stage_bytes);
            mbar_arrive(&empty_bar[s]);
        }
        if (acc == 0x1234567ull) *sink = acc;
    }
}

int main(int argc, char** argv) {
    const int64_t total_elems = 1000000000;  // 4 GB
    float* X;
    CK(cudaMalloc(&X, total_elems * 4));
    CK(cudaMemset(X, 0, total_elems * 4));
    unsigned long long* sink;
    CK(cudaMalloc(&sink, 8));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int rows_list[] = {10, 20, 40};
    const int tc_list[] = {128, 256, 512, 1024, 2048};
    for (int rows : rows_list) {
        const int64_t ld = total_elems / rows / 1024 * 1024, D = ld;
        for (int tc : tc_list) {
            if (tc > 256 && false) continue;
            const int stage_bytes = rows * tc * 4;
            for (int budget_kb : {100, 200}) {
                int stages = budget_kb * 1024 / stage_bytes;
                if (stages > 32) stages = 32;
                if (stages < 2) continue;
                for (int mode = 0; mode < 3; ++mode) {
                    if (mode == 1 && tc > 256) continue;
                    if (mode == 2 && rows > 32) continue;
                    CUtensorMap map{};
                    if (mode == 1) {
                        const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
                        const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
                        const cuuint32_t box[2] = {(cuuint32_t)tc, (cuuint32_t)rows};
                        const cuuint32_t es[2] = {1, 1};
                        if (cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, X, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                            printf("encode failed\n");
                            continue;
                        }
                    }
                    float best = 1e30f;
                    for (int rep = 0; rep < 4; ++rep) {
                        cudaEventRecord(e0);
                        stream_kernel<<<sms, 64, stages * stage_bytes>>>(X, map, rows, D, ld, tc, stages, mode, sink);
                        cudaEventRecord(e1);
                        CK(cudaEventSynchronize(e1));
                        float ms;
                        cudaEventElapsedTime(&ms, e0, e1);
                        if (rep > 0 && ms < best) best = ms;
                    }
                    const double bytes = (double)(D / tc) * tc * rows * 4;
                    printf("{\"rows\": %d, \"tile_cols\": %d, \"row_segment_bytes\": %d, \"stages\": %d, \"ring_kb\": %d, \"mode\": \"%s\", \"ms\": %.4f, \"GBps\": %.0f}\n",
                           rows, tc, tc * 4, stages, stages * stage_bytes / 1024, mode == 0 ? "bulk1d_one_lane" : (mode == 1 ? "tensor2d" : "bulk1d_lane_per_row"), best, bytes / best / 1e6);
                    fflush(stdout);
                }
            }
        }
    }
    return 0;
}
