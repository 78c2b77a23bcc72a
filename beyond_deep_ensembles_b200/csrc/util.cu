// util.cu — library info, Philox diagnostic entry and the multi-tensor gather/scatter.
#include <string>

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

// Gather/scatter between scattered tensors and one flat arena row.
// Replaces parameters_to_vector / cat+stack (svgd.py:83-84) and the per-parameter
// slice+clone scatter (svgd.py:92-97).  The table of up to kMtcChunk tensors travels in the
// kernel parameters (no device-side table to upload); one thread per 4 flat elements, the
// owning tensor is found by binary search over the ascending flat offsets.
constexpr int kMtcChunk = 112;
struct MtcTable {
    uint64_t ptr[kMtcChunk];
    int64_t off[kMtcChunk];
    int64_t size[kMtcChunk];
    int count;
    int64_t begin, end;  // flat range covered by this chunk
};

__device__ __forceinline__ int find_tensor(const MtcTable& t, int64_t e) {
    int lo = 0, hi = t.count - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t.off[mid] <= e)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

// inv_scale / found_inf (both device scalars, may be null): the gather modes multiply by *inv_scale and raise
// *found_inf to 1 on any non-finite SOURCE value — torch's _amp_foreach_non_finite_check_and_unscale_ (what
// GradScaler.unscale_ runs, algo.py:65-73) folded into the gather, so unscaling costs no extra pass.
__device__ __forceinline__ float4 mtc_unscale(float4 v, const float* inv_scale, float* found_inf) {
    if (inv_scale == nullptr) return v;
    if (!(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w))) *found_inf = 1.0f;
    const float s = *inv_scale;
    if (s == 1.0f) return v;
    return make_float4(__fmul_rn(v.x, s), __fmul_rn(v.y, s), __fmul_rn(v.z, s), __fmul_rn(v.w, s));
}

__global__ void __launch_bounds__(kEwThreads)
multi_tensor_copy_kernel(float* __restrict__ flat, const __grid_constant__ MtcTable tab, int mode,
                         const float* __restrict__ inv_scale, float* __restrict__ found_inf) {
    const int64_t base = tab.begin & ~static_cast<int64_t>(3);
    BDE_QUAD_LOOP(q, tab.end - base) {
        const int64_t e0 = base + (q << 2);
        const int t = find_tensor(tab, e0);
        const int64_t off = tab.off[t], sz = tab.size[t];
        float* tp = reinterpret_cast<float*>(tab.ptr[t]);
        const int64_t local = e0 - off;
        if (local >= 0 && local + 4 <= sz && ((reinterpret_cast<uintptr_t>(tp + local) & 15u) == 0) &&
            ((reinterpret_cast<uintptr_t>(flat + e0) & 15u) == 0)) {
            if (mode == 2) {
                stg_stream_f4(tp + local, ld_f4(flat + e0));
            } else {
                float4 v = mtc_unscale(ld_f4(tp + local), inv_scale, found_inf);
                if (mode == 1) {
                    const float4 a = ld_f4(flat + e0);
                    v = make_float4(__fadd_rn(a.x, v.x), __fadd_rn(a.y, v.y), __fadd_rn(a.z, v.z), __fadd_rn(a.w, v.w));
                }
                stg_stream_f4(flat + e0, v);
            }
        } else {
            for (int k = 0; k < 4; ++k) {
                const int64_t e = e0 + k;
                if (e < tab.begin || e >= tab.end) continue;
                const int tt = find_tensor(tab, e);
                const int64_t l2 = e - tab.off[tt];
                if (l2 < 0 || l2 >= tab.size[tt]) continue;  // padding between tensors
                float* p2 = reinterpret_cast<float*>(tab.ptr[tt]);
                if (mode == 2) {
                    p2[l2] = flat[e];
                } else {
                    float v = p2[l2];
                    if (inv_scale != nullptr) {
                        if (!isfinite(v)) *found_inf = 1.0f;
                        const float s = *inv_scale;
                        if (s != 1.0f) v = __fmul_rn(v, s);
                    }
                    flat[e] = (mode == 1) ? __fadd_rn(flat[e], v) : v;
                }
            }
        }
    }
}

}  // namespace bde

namespace bde {
Tuning& tuning() {
    static Tuning t;
    return t;
}
}  // namespace bde

using namespace bde;

extern "C" int bde_version(void) { return 100; }

extern "C" int bde_tune(const char* key, int value) {
    if (!key || value < 0) return BDE_ERR_INVALID_ARG;
    const std::string k(key);
    if (k == "pairdist_ctas_per_sm") tuning().pairdist_ctas_per_sm = value;
    else if (k == "apply_ctas_per_sm") tuning().apply_ctas_per_sm = value;
    else if (k == "ew_ctas_per_sm") tuning().ew_ctas_per_sm = value;
    else if (k == "apply_variant") tuning().apply_variant = value;
    else if (k == "pairdist_variant") tuning().pairdist_variant = value;
    else if (k == "ew_variant") tuning().ew_variant = value;
    else if (k == "apply_tile_sets") tuning().apply_tile_sets = value;
    else if (k == "swag_batch") tuning().swag_batch = value;
    else if (k == "batch_prefetch") tuning().batch_prefetch = value;
    else if (k == "batch_splits") tuning().batch_splits = value;
    else if (k == "gram_pairing") tuning().gram_pairing = value;
    else if (k == "gram_fold") tuning().gram_fold = value;
    else if (k == "gram_l2_promotion") tuning().gram_l2_promotion = value;
    else if (k == "ring_kb") tuning().ring_kb = value;
    else if (k == "gram_guard_x1000") tuning().gram_guard_x1000 = value;
    else return BDE_ERR_INVALID_ARG;
    return BDE_OK;
}

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
    return launch_ew(philox_normal_kernel, count, static_cast<cudaStream_t>(stream), out, count, seed, stream_id,
                     elem0 >> 2, aligned16(out) ? 1 : 0);
}

static int mtc_launch(float* flat, const uint64_t* ptrs_host, const int64_t* offsets_host, const int64_t* sizes_host,
                      int count, int mode, const float* inv_scale, float* found_inf, bde_stream_t stream) {
    if (!flat || !ptrs_host || !offsets_host || !sizes_host || count < 1 || mode < 0 || mode > 2)
        return BDE_ERR_INVALID_ARG;
    for (int i = 0; i < count; ++i) {
        if (sizes_host[i] < 0 || offsets_host[i] < 0) return BDE_ERR_INVALID_ARG;
        if (i > 0 && offsets_host[i] < offsets_host[i - 1] + sizes_host[i - 1]) return BDE_ERR_INVALID_ARG;
    }
    for (int c0 = 0; c0 < count; c0 += kMtcChunk) {
        MtcTable tab;
        tab.count = (count - c0 < kMtcChunk) ? count - c0 : kMtcChunk;
        for (int i = 0; i < tab.count; ++i) {
            tab.ptr[i] = ptrs_host[c0 + i];
            tab.off[i] = offsets_host[c0 + i];
            tab.size[i] = sizes_host[c0 + i];
        }
        tab.begin = tab.off[0];
        tab.end = tab.off[tab.count - 1] + tab.size[tab.count - 1];
        if (tab.end <= tab.begin) continue;
        const int rc = launch_ew(multi_tensor_copy_kernel, tab.end - (tab.begin & ~static_cast<int64_t>(3)),
                                 static_cast<cudaStream_t>(stream), flat, tab, mode, inv_scale, found_inf);
        if (rc != BDE_OK) return rc;
    }
    return BDE_OK;
}

extern "C" int bde_multi_tensor_copy(float* flat, const uint64_t* ptrs_host, const int64_t* offsets_host,
                                     const int64_t* sizes_host, int count, int mode, bde_stream_t stream) {
    return mtc_launch(flat, ptrs_host, offsets_host, sizes_host, count, mode, nullptr, nullptr, stream);
}

extern "C" int bde_multi_tensor_unscale_copy(float* flat, const uint64_t* ptrs_host, const int64_t* offsets_host,
                                             const int64_t* sizes_host, int count, int mode, const float* inv_scale,
                                             float* found_inf, bde_stream_t stream) {
    if (!inv_scale || !found_inf || mode > 1) return BDE_ERR_INVALID_ARG;
    return mtc_launch(flat, ptrs_host, offsets_host, sizes_host, count, mode, inv_scale, found_inf, stream);
}
