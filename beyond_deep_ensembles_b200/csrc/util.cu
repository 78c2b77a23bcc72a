// util.cu — library info, Philox diagnostic entry and the multi-tensor gather/scatter.
#include "elementwise.cuh"

namespace bde {

__global__ void __launch_bounds__(kEwThreads)
philox_normal_kernel(float* __restrict__ out, int64_t count, uint64_t seed, uint64_t stream_id, int64_t quad0,
                     int vec) {
    BDE_QUAD_LOOP(q, count) {
        const float4 z = philox_normal4(seed, stream_id, static_cast<uint64_t>(quad0 + q));
        if (vec)
            store_quad<true>(out, q << 2, count, z);
        else
            store_quad<false>(out, q << 2, count, z);
    }
}

// Gather/scatter between `count` scattered tensors and one flat arena row.
// Replaces parameters_to_vector / cat+stack (svgd.py:83-84) and the per-parameter
// slice+clone scatter (svgd.py:92-97).  One thread per 4 flat elements; the owning tensor is
// found by binary search over the (ascending) flat offsets.
__device__ __forceinline__ int find_tensor(const int64_t* __restrict__ offsets, int count, int64_t e) {
    int lo = 0, hi = count - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(offsets + mid) <= e)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(kEwThreads)
multi_tensor_copy_kernel(float* __restrict__ flat, const uint64_t* __restrict__ ptrs,
                         const int64_t* __restrict__ offsets, const int64_t* __restrict__ sizes, int count,
                         int64_t total, int mode) {
    BDE_QUAD_LOOP(q, total) {
        const int64_t e0 = q << 2;
        const int t = find_tensor(offsets, count, e0);
        const int64_t off = __ldg(offsets + t), sz = __ldg(sizes + t);
        float* tp = reinterpret_cast<float*>(__ldg(ptrs + t));
        const int64_t local = e0 - off;
        if (local >= 0 && local + 4 <= sz && e0 + 4 <= total && ((reinterpret_cast<uintptr_t>(tp + local) & 15u) == 0) &&
            ((reinterpret_cast<uintptr_t>(flat + e0) & 15u) == 0)) {
            if (mode == 2) {
                stg_stream_f4(tp + local, ld_f4(flat + e0));
            } else {
                float4 v = ld_f4(tp + local);
                if (mode == 1) {
                    const float4 a = ld_f4(flat + e0);
                    v = make_float4(__fadd_rn(a.x, v.x), __fadd_rn(a.y, v.y), __fadd_rn(a.z, v.z), __fadd_rn(a.w, v.w));
                }
                stg_stream_f4(flat + e0, v);
            }
        } else {
            for (int k = 0; k < 4; ++k) {
                const int64_t e = e0 + k;
                if (e >= total) break;
                const int tt = find_tensor(offsets, count, e);
                const int64_t o2 = __ldg(offsets + tt), s2 = __ldg(sizes + tt);
                const int64_t l2 = e - o2;
                if (l2 < 0 || l2 >= s2) continue;  // padding between tensors
                float* p2 = reinterpret_cast<float*>(__ldg(ptrs + tt));
                if (mode == 2)
                    p2[l2] = flat[e];
                else if (mode == 1)
                    flat[e] = __fadd_rn(flat[e], p2[l2]);
                else
                    flat[e] = p2[l2];
            }
        }
    }
}

}  // namespace bde

using namespace bde;

extern "C" int bde_version(void) { return 100; }

extern "C" const char* bde_error_string(int code) {
    switch (code) {
        case BDE_OK:
            return "ok";
        case BDE_ERR_INVALID_ARG:
            return "bde: invalid argument";
        case BDE_ERR_ALIGNMENT:
            return "bde: pointer or stride not 16-byte aligned";
        case BDE_ERR_WORKSPACE:
            return "bde: workspace missing or too small";
        case BDE_ERR_UNSUPPORTED_N:
            return "bde: particle count not supported";
        default:
            if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
            return "bde: unknown error";
    }
}

extern "C" int bde_device_sm_count(int* sm_count) {
    if (!sm_count) return BDE_ERR_INVALID_ARG;
    int dev = 0;
    BDE_RETURN_IF_CUDA(cudaGetDevice(&dev));
    BDE_RETURN_IF_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    return BDE_OK;
}

extern "C" int bde_philox_normal(float* out, int64_t count, uint64_t seed, uint64_t stream_id, int64_t elem0,
                                 bde_stream_t stream) {
    if (!out || count < 0 || elem0 < 0 || (elem0 & 3)) return BDE_ERR_INVALID_ARG;
    if (count == 0) return BDE_OK;
    const EwGrid g = ew_grid(count, kEwThreads, kEwCtasPerSm);
    philox_normal_kernel<<<g.blocks, g.threads, 0, static_cast<cudaStream_t>(stream)>>>(out, count, seed, stream_id,
                                                                                       elem0 >> 2, aligned16(out) ? 1 : 0);
    BDE_CHECK_LAUNCH();
    return BDE_OK;
}

extern "C" int bde_multi_tensor_copy(float* flat, const uint64_t* ptrs, const int64_t* offsets, const int64_t* sizes,
                                     int count, int64_t total, int mode, bde_stream_t stream) {
    if (!flat || !ptrs || !offsets || !sizes || count < 1 || total < 0 || mode < 0 || mode > 2)
        return BDE_ERR_INVALID_ARG;
    if (total == 0) return BDE_OK;
    const EwGrid g = ew_grid(total, kEwThreads, kEwCtasPerSm);
    multi_tensor_copy_kernel<<<g.blocks, g.threads, 0, static_cast<cudaStream_t>(stream)>>>(flat, ptrs, offsets, sizes,
                                                                                           count, total, mode);
    BDE_CHECK_LAUNCH();
    return BDE_OK;
}
