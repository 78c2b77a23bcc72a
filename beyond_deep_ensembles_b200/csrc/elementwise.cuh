// elementwise.cuh — quad (4-element) load/store helpers for the streaming elementwise kernels.
// Every kernel walks the vector in aligned groups of four elements (the Philox granularity);
// a full group on 16-byte-aligned pointers moves with one 128-bit access, anything else
// (ragged tail, unaligned views) falls back to guarded scalar accesses in the same kernel.
#pragma once
#include <map>
#include <mutex>

#include "common.cuh"

namespace bde {

// VEC = all participating pointers are 16-byte aligned
template <bool VEC, bool NC>
__device__ __forceinline__ float4 load_quad(const float* __restrict__ p, int64_t base, int64_t n) {
    if (VEC && base + 4 <= n) return NC ? ldg_stream_f4(p + base) : ld_f4(p + base);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (base + 0 < n) v.x = p[base + 0];
    if (base + 1 < n) v.y = p[base + 1];
    if (base + 2 < n) v.z = p[base + 2];
    if (base + 3 < n) v.w = p[base + 3];
    return v;
}

template <bool VEC>
__device__ __forceinline__ void store_quad(float* __restrict__ p, int64_t base, int64_t n, float4 v) {
    if (VEC && base + 4 <= n) {
        stg_stream_f4(p + base, v);
        return;
    }
    if (base + 0 < n) p[base + 0] = v.x;
    if (base + 1 < n) p[base + 1] = v.y;
    if (base + 2 < n) p[base + 2] = v.z;
    if (base + 3 < n) p[base + 3] = v.w;
}

#define BDE_QUAD_LOOP(q, n_elems)                                                                      \
    for (int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x, _nq = ((n_elems) + 3) >> 2, \
                 _st = static_cast<int64_t>(gridDim.x) * blockDim.x;                                   \
         q < _nq; q += _st)

// apply a scalar functor lane-wise to float4s
#define BDE_LANES(expr_x, expr_y, expr_z, expr_w) make_float4(expr_x, expr_y, expr_z, expr_w)

// block-level fp64 sum -> cta_val (shared), returns after __syncthreads
__device__ __forceinline__ void block_sum_fp64(double v, double* cta_val) {
    __shared__ double warp_part[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) warp_part[warp] = v;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        double s = lane < nw ? warp_part[lane] : 0.0;
        s = warp_sum(s);
        if (lane == 0) *cta_val = s;
    }
    __syncthreads();
}

constexpr int kEwThreads = 256;
constexpr int kMaxCtasEw = 148 * 8 * 2;

// Resident CTAs per SM of a kernel at kEwThreads threads, cached per kernel.
inline int ew_occupancy(const void* kernel) {
    static std::mutex mu;
    static std::map<const void*, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(kernel);
    if (it != cache.end()) return it->second;
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kEwThreads, 0) != cudaSuccess || v < 1) v = 1;
    cache[kernel] = v;
    return v;
}

// One full wave of persistent CTAs (SMs x measured occupancy), grid-stride over the quads:
// a grid larger than what is resident leaves a partial second wave (measured: -25 % bandwidth).
template <typename... KArgs, typename... Args>
inline int launch_ew(void (*kernel)(KArgs...), int64_t n_elems, cudaStream_t st, Args... args) {
    const EwGrid g = ew_grid(n_elems, kEwThreads, ew_occupancy(reinterpret_cast<const void*>(kernel)));
    kernel<<<g.blocks, g.threads, 0, st>>>(args...);
    BDE_CHECK_LAUNCH();
    return BDE_OK;
}

}  // namespace bde
