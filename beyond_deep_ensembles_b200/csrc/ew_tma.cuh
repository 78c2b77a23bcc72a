// ew_tma.cuh — TMA-staged streaming form of the elementwise kernels (K3, K5, K6, K7, K10).
//
// One persistent CTA per SM.  A producer lane streams kEwTile-element tiles of every INPUT
// vector into a ring of shared-memory stages with cp.async.bulk (SASS UBLKCP, completion on
// "full" mbarriers); 16 consumer warps pull one quad (4 elements) per thread out of the stage,
// hand the stage back ("empty" mbarrier) and only then do the arithmetic and the 128-bit
// no-allocate stores.  The bytes in flight per SM are set by the ring (~190 KB) instead of by
// register occupancy, and the loads are decoupled from the (division-heavy) per-element math.
// Outputs may alias inputs (in-place state updates): a tile is consumed by exactly one CTA and
// the producer only ever prefetches OTHER tiles.
//
// Op interface (a plain struct passed by value):
//     static constexpr int NIN, NOUT;
//     __device__ void operator()(const float4 (&in)[NIN], float4 (&out)[NOUT], int64_t quad) const;
// `quad` is the index of the 4-element group inside the vector (the Philox counter offset).
#pragma once
#include "elementwise.cuh"

namespace bde {

constexpr int kEwTile = 2048;          // floats per input vector per stage (8 KB)
constexpr int kEwConsumers = kEwTile / 4;  // 512 threads, one quad each
constexpr int kEwRingBytes = 192 * 1024;

__host__ __device__ constexpr int ew_tma_stages(int nin) {
    return kEwRingBytes / (nin * kEwTile * 4) > 8 ? 8 : kEwRingBytes / (nin * kEwTile * 4);
}

template <int NIN, int NOUT>
struct EwPtrs {
    const float* in[NIN];
    float* out[NOUT];
};

// block-level fp64 sum of the consumer threads only (the producer warp does not take part)
__device__ __forceinline__ void consumer_sum_fp64(double v, double* cta_val, double* warp_part) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) warp_part[warp] = v;
    asm volatile("bar.sync 1, %0;" ::"n"(kEwConsumers) : "memory");
    if (warp == 0) {
        double s = lane < kEwConsumers / 32 ? warp_part[lane] : 0.0;
        s = warp_sum(s);
        if (lane == 0) *cta_val = s;
    }
}

// REDUCE: Op additionally provides `double value(const float4 (&in)[NIN]) const`, summed in fp64 over
// the vector with the deterministic last-CTA reduction; the last CTA calls op.finish(total).
template <class Op, bool REDUCE>
__global__ void __launch_bounds__(kEwConsumers + 32, 1)
ew_tma_kernel(EwPtrs<Op::NIN, Op::NOUT> p, int64_t D, Op op, void* ws) {
    constexpr int NIN = Op::NIN, NOUT = Op::NOUT;
    constexpr int STAGES = ew_tma_stages(NIN);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);  // [STAGES][NIN][kEwTile]
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ double warp_part[32];
    __shared__ double cta_val;
    __shared__ double total;

    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kEwConsumers / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t d4 = D & ~static_cast<int64_t>(3);
    const int64_t ntiles = (d4 + kEwTile - 1) / kEwTile;
    double vsum = 0.0;

    if (tid >= kEwConsumers) {
        if (tid == kEwConsumers) {  // one elected lane drives the copy engine
            int it = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const int s = it % STAGES;
                const uint32_t use = static_cast<uint32_t>(it / STAGES);
                mbar_wait(&empty_bar[s], (use & 1u) ^ 1u);
                const int64_t col0 = t * kEwTile;
                const int64_t w = (d4 - col0 < kEwTile) ? d4 - col0 : kEwTile;
                const uint32_t bytes = static_cast<uint32_t>(w) * 4u;
                mbar_arrive_expect_tx(&full_bar[s], NIN * bytes);
                float* dst = tiles + static_cast<size_t>(s) * NIN * kEwTile;
#pragma unroll
                for (int k = 0; k < NIN; ++k) tma_load_1d(dst + k * kEwTile, p.in[k] + col0, bytes, &full_bar[s]);
            }
        }
    } else {
        const int lane = tid & 31;
        int it = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int s = it % STAGES;
            const uint32_t use = static_cast<uint32_t>(it / STAGES);
            const int64_t col0 = t * kEwTile;
            const int64_t w = (d4 - col0 < kEwTile) ? d4 - col0 : kEwTile;
            const bool active = 4 * tid < w;
            mbar_wait(&full_bar[s], use & 1u);
            const float* src = tiles + static_cast<size_t>(s) * NIN * kEwTile + 4 * tid;
            float4 in[NIN];
            if (active) {
#pragma unroll
                for (int k = 0; k < NIN; ++k) in[k] = *reinterpret_cast<const float4*>(src + k * kEwTile);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);  // operands are in registers: release the stage
            if (active) {
                float4 o[NOUT > 0 ? NOUT : 1];
                const int64_t b = col0 + 4 * tid;
                op(in, o, b >> 2);
                if constexpr (REDUCE) vsum += op.value(in);
#pragma unroll
                for (int k = 0; k < NOUT; ++k)
                    if (p.out[k]) stg_stream_f4(p.out[k] + b, o[k]);
            }
        }
        // ragged tail (D % 4 elements): one thread, guarded scalar accesses
        if (blockIdx.x == 0 && tid == 0 && d4 < D) {
            float4 in[NIN];
#pragma unroll
            for (int k = 0; k < NIN; ++k) in[k] = load_quad<false, false>(p.in[k], d4, D);
            float4 o[NOUT > 0 ? NOUT : 1];
            op(in, o, d4 >> 2);
            if constexpr (REDUCE) vsum += op.value_tail(in, static_cast<int>(D - d4));
#pragma unroll
            for (int k = 0; k < NOUT; ++k)
                if (p.out[k]) store_quad<false>(p.out[k], d4, D, o[k]);
        }
        if constexpr (REDUCE) consumer_sum_fp64(vsum, &cta_val, warp_part);
    }
    if constexpr (REDUCE) {
        __syncthreads();
        if (grid_reduce_fp64(&cta_val, 1, ws, &total)) {
            if (tid == 0) op.finish(total);
        }
    }
}

template <class Op, bool REDUCE = false>
inline int launch_ew_tma(const EwPtrs<Op::NIN, Op::NOUT>& p, int64_t D, const Op& op, void* ws, cudaStream_t st) {
    constexpr int smem = ew_tma_stages(Op::NIN) * Op::NIN * kEwTile * 4;
    static_assert(ew_tma_stages(Op::NIN) >= 2, "too many input vectors for the ring");
    static bool configured = false;
    if (!configured) {
        BDE_RETURN_IF_CUDA(cudaFuncSetAttribute(ew_tma_kernel<Op, REDUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int64_t ntiles = ((D & ~static_cast<int64_t>(3)) + kEwTile - 1) / kEwTile;
    int64_t grid = sm_count_cached();
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    ew_tma_kernel<Op, REDUCE><<<static_cast<unsigned>(grid), kEwConsumers + 32, smem, st>>>(p, D, op, ws);
    BDE_CHECK_LAUNCH();
    return BDE_OK;
}

// `vec` = every pointer 16-byte aligned; `prefer` = this kernel measured faster in the staged form
//  — then it is used once every SM has a few tiles.  bde_tune("ew_variant") forces.
inline bool use_ew_tma(int64_t D, bool vec, bool prefer) {
    const int v = tuning().ew_variant;
    if (!vec || v == 1) return false;
    if (v == 2) return D >= 4;
    return prefer && D >= static_cast<int64_t>(4) * kEwTile * sm_count_cached();
}

}  // namespace bde
