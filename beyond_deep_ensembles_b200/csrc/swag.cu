// swag.cu — SWAG running moments / deviation ring buffer / low-rank + diagonal sampling.
// Reference arithmetic: src/algos/swag.py:91-114 and torch LowRankMultivariateNormal.rsample.
// Operation order and rounding follow the reference op by op (explicit *_rn intrinsics, no
// FMA contraction) so that the update matches eager fp32 PyTorch.
#include "ew_tma.cuh"

namespace bde {

// K3 ------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kEwThreads)
swag_update_kernel(const float* __restrict__ theta, float* __restrict__ mean, float* __restrict__ sq,
                   float* __restrict__ dev_row, int64_t D, float fu, float fu1) {
    BDE_QUAD_LOOP(q, D) {
        const int64_t b = q << 2;
        const float4 t = load_quad<VEC, true>(theta, b, D);
        const float4 m = load_quad<VEC, false>(mean, b, D);
        const float4 s = load_quad<VEC, false>(sq, b, D);
        // swag.py:101  mean = (updates*mean + params) / (updates + 1)
        auto upd_mean = [&](float mv, float tv) { return __fdiv_rn(__fadd_rn(__fmul_rn(fu, mv), tv), fu1); };
        // swag.py:102  sq = (updates*sq + params**2) / (updates + 1)
        auto upd_sq = [&](float sv, float tv) { return __fdiv_rn(__fadd_rn(__fmul_rn(fu, sv), __fmul_rn(tv, tv)), fu1); };
        const float4 mn = BDE_LANES(upd_mean(m.x, t.x), upd_mean(m.y, t.y), upd_mean(m.z, t.z), upd_mean(m.w, t.w));
        const float4 sn = BDE_LANES(upd_sq(s.x, t.x), upd_sq(s.y, t.y), upd_sq(s.z, t.z), upd_sq(s.w, t.w));
        // swag.py:104  deviations[:, -1] = params - mean_new
        const float4 dv = BDE_LANES(__fsub_rn(t.x, mn.x), __fsub_rn(t.y, mn.y), __fsub_rn(t.z, mn.z), __fsub_rn(t.w, mn.w));
        store_quad<VEC>(mean, b, D, mn);
        store_quad<VEC>(sq, b, D, sn);
        store_quad<VEC>(dev_row, b, D, dv);
    }
}

// K3, TMA-staged (ew_tma.cuh): inputs theta, mean, sq; outputs mean, sq, dev_row
struct SwagUpdateOp {
    static constexpr int NIN = 3, NOUT = 3;
    float fu, fu1;
    __device__ __forceinline__ void operator()(const float4 (&in)[NIN], float4 (&out)[NOUT], int64_t) const {
        const float4 t = in[0], m = in[1], s = in[2];
        auto upd_mean = [&](float mv, float tv) { return __fdiv_rn(__fadd_rn(__fmul_rn(fu, mv), tv), fu1); };
        auto upd_sq = [&](float sv, float tv) { return __fdiv_rn(__fadd_rn(__fmul_rn(fu, sv), __fmul_rn(tv, tv)), fu1); };
        const float4 mn = BDE_LANES(upd_mean(m.x, t.x), upd_mean(m.y, t.y), upd_mean(m.z, t.z), upd_mean(m.w, t.w));
        out[0] = mn;
        out[1] = BDE_LANES(upd_sq(s.x, t.x), upd_sq(s.y, t.y), upd_sq(s.z, t.z), upd_sq(s.w, t.w));
        out[2] = BDE_LANES(__fsub_rn(t.x, mn.x), __fsub_rn(t.y, mn.y), __fsub_rn(t.z, mn.z), __fsub_rn(t.w, mn.w));
    }
};

// K4 ------------------------------------------------------------------------------------
constexpr int kMaxSwagRank = 256;

template <bool VEC, int KU>
__global__ void __launch_bounds__(kEwThreads)
swag_sample_kernel(const float* __restrict__ mean, const float* __restrict__ sq, const float* __restrict__ dev, int K,
                   int head, int64_t D, int64_t ld, const float* __restrict__ eps_k, const float* __restrict__ eps_d,
                   uint64_t seed, uint64_t stream_id, int64_t quad0, float inv_norm_den, float* __restrict__ theta) {
    // zc[k] = z[k] / sqrt(2(K-1)) for the LOGICAL column k (0 = oldest); rowoff[k] = physical row offset
    __shared__ float zc[kMaxSwagRank];
    __shared__ int64_t rowoff[kMaxSwagRank];
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float z;
        if (eps_k) {
            z = eps_k[k];
        } else {
            const float4 z4 = philox_normal4(seed, stream_id ^ 0x5741ull, static_cast<uint64_t>(k >> 2));
            z = (k & 3) == 0 ? z4.x : (k & 3) == 1 ? z4.y : (k & 3) == 2 ? z4.z : z4.w;
        }
        zc[k] = __fdiv_rn(z, inv_norm_den);  // cov_factor = deviations / sqrt(2(K-1)), swag.py:113
        rowoff[k] = static_cast<int64_t>((head + k) % K) * ld;
    }
    __syncthreads();

    BDE_QUAD_LOOP(q, D) {
        const int64_t b = q << 2;
        const float4 m = load_quad<VEC, true>(mean, b, D);
        const float4 s = load_quad<VEC, true>(sq, b, D);
        float4 low = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k0 = 0; k0 < K; k0 += KU) {
            float4 d[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u)
                if (k0 + u < K) d[u] = load_quad<VEC, true>(dev + rowoff[k0 + u], b, D);
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                if (k0 + u < K) {
                    const float z = zc[k0 + u];
                    low.x = fmaf(d[u].x, z, low.x);
                    low.y = fmaf(d[u].y, z, low.y);
                    low.z = fmaf(d[u].z, z, low.z);
                    low.w = fmaf(d[u].w, z, low.w);
                }
            }
        }
        float4 e;
        if (eps_d) {
            e = load_quad<VEC, true>(eps_d, b, D);
        } else {
            e = philox_normal4(seed, stream_id, static_cast<uint64_t>(quad0 + q));
        }
        // swag.py:112  diag = 0.5 * (relu(sq - mean**2) + 1e-6);  rsample: loc + W@eps_W + diag.sqrt()*eps_D
        auto fin = [&](float mv, float sv, float lv, float ev) {
            float v = __fsub_rn(sv, __fmul_rn(mv, mv));
            v = fmaxf(v, 0.0f);
            v = __fmul_rn(0.5f, __fadd_rn(v, 1e-6f));
            return __fadd_rn(__fadd_rn(mv, lv), __fmul_rn(__fsqrt_rn(v), ev));
        };
        const float4 o = BDE_LANES(fin(m.x, s.x, low.x, e.x), fin(m.y, s.y, low.y, e.y), fin(m.z, s.z, low.z, e.z),
                                   fin(m.w, s.w, low.w, e.w));
        store_quad<VEC>(theta, b, D, o);
    }
}

// K4 batched: S draws from the SAME posterior in one pass — mean, sq and the K deviation rows
// are read once and S weight vectors are written, 4(K+2)D + 4 S D bytes instead of S * 4(K+3)D.  Draw s is
// bit-identical to bde_swag_sample with stream_id + s: the same fmaf chain over k, the same Philox counters.
constexpr int kSwagBatchMax = 16;   // draws per launch (4 accumulator registers each)

template <bool VEC, int SB>
__global__ void __launch_bounds__(kEwThreads, SB >= 16 ? 2 : 3)
swag_sample_batch_kernel(const float* __restrict__ mean, const float* __restrict__ sq, const float* __restrict__ dev, int K,
                         int head, int64_t D, int64_t ld, int S, const float* __restrict__ eps_k,
                         const float* __restrict__ eps_d, int64_t ld_eps, uint64_t seed, uint64_t stream_id, int64_t quad0,
                         float inv_norm_den, float* __restrict__ theta, int64_t ld_out) {
    // zc[k][s] = z_s[k] / sqrt(2(K-1)) for the logical column k of draw s (0 for the unused slots s >= S)
    __shared__ __align__(16) float zc[kMaxSwagRank][SB];
    __shared__ int64_t rowoff[kMaxSwagRank];
    for (int e = threadIdx.x; e < K * SB; e += blockDim.x) {
        const int k = e / SB, sidx = e - k * SB;
        float z = 0.0f;
        if (sidx < S) {
            if (eps_k) {
                z = eps_k[static_cast<int64_t>(sidx) * K + k];
            } else {
                const float4 z4 = philox_normal4(seed, (stream_id + sidx) ^ 0x5741ull, static_cast<uint64_t>(k >> 2));
                z = (k & 3) == 0 ? z4.x : (k & 3) == 1 ? z4.y : (k & 3) == 2 ? z4.z : z4.w;
            }
            z = __fdiv_rn(z, inv_norm_den);
        }
        zc[k][sidx] = z;
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x) rowoff[k] = static_cast<int64_t>((head + k) % K) * ld;
    __syncthreads();

    BDE_QUAD_LOOP(q, D) {
        const int64_t b = q << 2;
        const float4 m = load_quad<VEC, true>(mean, b, D);
        const float4 sv = load_quad<VEC, true>(sq, b, D);
        float4 low[SB];
#pragma unroll
        for (int sidx = 0; sidx < SB; ++sidx) low[sidx] = make_float4(0.f, 0.f, 0.f, 0.f);
        constexpr int KU = 5;   // deviation rows in flight per thread, as in the single-draw kernel
        for (int k0 = 0; k0 < K; k0 += KU) {
            float4 d[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u)
                if (k0 + u < K) d[u] = load_quad<VEC, true>(dev + rowoff[k0 + u], b, D);
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                if (k0 + u < K) {
#pragma unroll
                    for (int s4 = 0; s4 < SB; s4 += (SB >= 4 ? 4 : 2)) {
                        float z[4];
                        if constexpr (SB >= 4) {
                            const float4 zz = *reinterpret_cast<const float4*>(&zc[k0 + u][s4]);
                            z[0] = zz.x, z[1] = zz.y, z[2] = zz.z, z[3] = zz.w;
                        } else {
                            const float2 zz = *reinterpret_cast<const float2*>(&zc[k0 + u][s4]);
                            z[0] = zz.x, z[1] = zz.y, z[2] = z[3] = 0.0f;
                        }
#pragma unroll
                        for (int t = 0; t < (SB >= 4 ? 4 : 2); ++t) {
                            low[s4 + t].x = fmaf(d[u].x, z[t], low[s4 + t].x);
                            low[s4 + t].y = fmaf(d[u].y, z[t], low[s4 + t].y);
                            low[s4 + t].z = fmaf(d[u].z, z[t], low[s4 + t].z);
                            low[s4 + t].w = fmaf(d[u].w, z[t], low[s4 + t].w);
                        }
                    }
                }
            }
        }
        // diag = 0.5 * (relu(sq - mean**2) + 1e-6), shared by all draws (swag.py:112)
        auto sdev = [&](float mv, float s2) {
            float v = __fsub_rn(s2, __fmul_rn(mv, mv));
            v = fmaxf(v, 0.0f);
            return __fsqrt_rn(__fmul_rn(0.5f, __fadd_rn(v, 1e-6f)));
        };
        const float4 sd = BDE_LANES(sdev(m.x, sv.x), sdev(m.y, sv.y), sdev(m.z, sv.z), sdev(m.w, sv.w));
#pragma unroll
        for (int sidx = 0; sidx < SB; ++sidx) {
            if (sidx < S) {
                float4 e;
                if (eps_d) {
                    e = load_quad<VEC, true>(eps_d + sidx * ld_eps, b, D);
                } else {
                    e = philox_normal4(seed, stream_id + sidx, static_cast<uint64_t>(quad0 + q));
                }
                auto fin = [&](float mv, float lv, float dv, float ev) { return __fadd_rn(__fadd_rn(mv, lv), __fmul_rn(dv, ev)); };
                const float4 o = BDE_LANES(fin(m.x, low[sidx].x, sd.x, e.x), fin(m.y, low[sidx].y, sd.y, e.y),
                                           fin(m.z, low[sidx].z, sd.z, e.z), fin(m.w, low[sidx].w, sd.w, e.w));
                store_quad<VEC>(theta + sidx * ld_out, b, D, o);
            }
        }
    }
}

// Fast path of K4 batched: every pointer 16-byte aligned, all SB draws of the pass in use, whole quads only (the launcher
// sends a ragged tail and partial passes to the general kernel above).  Same arithmetic, bit for bit — the low-rank chain
// runs as packed FFMA2 (IEEE per lane, same k order), the epilogue as packed add / mul without contraction — but none of
// the per-draw predicates, guarded loads and 64-bit address recomputation of the general form: measured 0.99 -> see
// profiles/r02_bench_n1.json (16 draws at ResNet-50 size).
// (KU = deviation rows in flight per thread; 5 measured best against 2 / 3 / 4 with the L2 prefetch in place: 0.64 / 0.70 /
// 0.67 / 0.71 ms, profiles/r02_batch_samplers.jsonl)
template <int SB, bool INJ, int KU = 5>
__global__ void __launch_bounds__(kEwThreads, SB >= 16 ? 2 : 3)
swag_sample_batch_fast_kernel(const float* __restrict__ mean, const float* __restrict__ sq, const float* __restrict__ dev, int K,
                              int head, int64_t nquads, int64_t ld, const float* __restrict__ eps_k,
                              const float* __restrict__ eps_d, int64_t ld_eps, uint64_t seed, uint64_t stream_id, int64_t quad0,
                              float inv_norm_den, float* __restrict__ theta, int64_t ld_out, int pf_dist, int splits) {
    // `splits` (1 by default) neighbouring CTAs share the quads of one CTA slot and take SB draws each (draws part * SB ..
    // part * SB + SB - 1 of the pass).  The partner CTAs are co-resident and walk the same quads at the same time, so the
    // second reader of a deviation row finds it in L2 (ncu: +19 % DRAM reads at 2 splits); see the launcher for why it is off.
    const int part = static_cast<int>(blockIdx.x) % splits;
    const int64_t slot = blockIdx.x / splits, nslots = gridDim.x / splits;
    stream_id += static_cast<uint64_t>(part) * SB;
    theta += static_cast<int64_t>(part) * SB * ld_out;
    if (eps_k) eps_k += static_cast<int64_t>(part) * SB * K;
    if (INJ) eps_d += static_cast<int64_t>(part) * SB * ld_eps;
    __shared__ __align__(16) float zc[kMaxSwagRank][SB];
    __shared__ int64_t rowoff[kMaxSwagRank];
    for (int e = threadIdx.x; e < K * SB; e += blockDim.x) {
        const int k = e / SB, sidx = e - k * SB;
        float z;
        if (eps_k) {
            z = eps_k[static_cast<int64_t>(sidx) * K + k];
        } else {
            const float4 z4 = philox_normal4(seed, (stream_id + sidx) ^ 0x5741ull, static_cast<uint64_t>(k >> 2));
            z = (k & 3) == 0 ? z4.x : (k & 3) == 1 ? z4.y : (k & 3) == 2 ? z4.z : z4.w;
        }
        zc[k][sidx] = __fdiv_rn(z, inv_norm_den);
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x) rowoff[k] = static_cast<int64_t>((head + k) % K) * ld;
    __syncthreads();

    const PhiloxKeys pk = philox_round_keys(seed);
    const int64_t stride = nslots * blockDim.x;
    for (int64_t q = slot * blockDim.x + threadIdx.x; q < nquads; q += stride) {
        const int64_t b = q << 2;
        // The 128 registers of the 16 accumulator pairs leave two CTAs (4 warps per scheduler) to hide the K + 2 loads of a
        // quad (ncu: long-scoreboard stalls 2.3 per issue).  The lines of the quad this thread visits `pf_dist` iterations
        // from now are therefore requested into L2 here — ~2300 instructions ahead of their use, no register held.
        const int64_t qn = q + pf_dist * stride;
        if (pf_dist && qn < nquads) {
            const int64_t bn = qn << 2;
            prefetch_l2(mean + bn);
            prefetch_l2(sq + bn);
            for (int k = 0; k < K; ++k) prefetch_l2(dev + rowoff[k] + bn);
        }
        const V4 m = ldg_stream_v4(mean + b);
        const V4 sv = ldg_stream_v4(sq + b);
        f32x2 lo[SB], hi[SB];
#pragma unroll
        for (int sidx = 0; sidx < SB; ++sidx) lo[sidx] = hi[sidx] = 0ull;
        for (int k0 = 0; k0 < K; k0 += KU) {
            V4 d[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u)
                if (k0 + u < K) d[u] = ldg_stream_v4(dev + rowoff[k0 + u] + b);
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                if (k0 + u < K) {
#pragma unroll
                    for (int s4 = 0; s4 < SB; s4 += (SB >= 4 ? 4 : 2)) {
                        float z[4];
                        if constexpr (SB >= 4) {
                            const float4 zz = *reinterpret_cast<const float4*>(&zc[k0 + u][s4]);
                            z[0] = zz.x, z[1] = zz.y, z[2] = zz.z, z[3] = zz.w;
                        } else {
                            const float2 zz = *reinterpret_cast<const float2*>(&zc[k0 + u][s4]);
                            z[0] = zz.x, z[1] = zz.y, z[2] = z[3] = 0.0f;
                        }
#pragma unroll
                        for (int t = 0; t < (SB >= 4 ? 4 : 2); ++t) {
                            lo[s4 + t] = fma2s(z[t], d[u].lo, lo[s4 + t]);   // fmaf(d, z, low) per lane, k ascending
                            hi[s4 + t] = fma2s(z[t], d[u].hi, hi[s4 + t]);
                        }
                    }
                }
            }
        }
        float m0, m1, m2, m3, s0, s1, s2, s3;
        unpack2(m.lo, m0, m1);
        unpack2(m.hi, m2, m3);
        unpack2(sv.lo, s0, s1);
        unpack2(sv.hi, s2, s3);
        auto sdev = [&](float mv, float sq2) {   // sqrt(0.5 (relu(sq - mean^2) + 1e-6)), swag.py:112
            float v = __fsub_rn(sq2, __fmul_rn(mv, mv));
            v = fmaxf(v, 0.0f);
            return __fsqrt_rn(__fmul_rn(0.5f, __fadd_rn(v, 1e-6f)));
        };
        const float sd0 = sdev(m0, s0), sd1 = sdev(m1, s1), sd2 = sdev(m2, s2), sd3 = sdev(m3, s3);
        float* out = theta + b;
        const float* ein = INJ ? eps_d + b : nullptr;
#pragma unroll
        for (int sidx = 0; sidx < SB; ++sidx) {
            float4 z;
            if constexpr (INJ) {
                z = ldg_stream_f4(ein);
                ein += ld_eps;
            } else {
                z = philox_normal4(pk, stream_id + sidx, static_cast<uint64_t>(quad0 + q));
            }
            // (mean + low) + sqrt(diag) * eps, each operation rounded on its own.  The products stay scalar: ptxas contracts a
            // packed mul followed by a packed add into FFMA2 even when both carry .rn (seen in SASS), which changes the bits.
            V4 o;
            o.lo = add2_rn(add2_rn(m.lo, lo[sidx]), pack2(__fmul_rn(sd0, z.x), __fmul_rn(sd1, z.y)));
            o.hi = add2_rn(add2_rn(m.hi, hi[sidx]), pack2(__fmul_rn(sd2, z.z), __fmul_rn(sd3, z.w)));
            stg_stream_v4(out, o);
            out += ld_out;
        }
    }
}

}  // namespace bde

using namespace bde;

// One full wave of CTAs like launch_ew, rounded down to a multiple of `splits` (partner CTAs share quads).
template <typename... KArgs, typename... Args>
static int launch_swag_fast(void (*kernel)(KArgs...), int splits, int64_t nquads, cudaStream_t st, Args... args) {
    const int per_sm = ew_occupancy(reinterpret_cast<const void*>(kernel));
    int64_t slots = (nquads + kEwThreads - 1) / kEwThreads;
    const int64_t cap = (static_cast<int64_t>(sm_count_cached()) * per_sm) / splits;
    if (slots > cap) slots = cap;
    if (slots < 1) slots = 1;
    kernel<<<static_cast<unsigned>(slots * splits), kEwThreads, 0, st>>>(args...);
    BDE_CHECK_LAUNCH();
    return BDE_OK;
}

extern "C" int bde_swag_update(const float* theta, float* mean, float* sq, float* dev_row, int64_t D, int64_t updates,
                               bde_stream_t stream) {
    if (!theta || !mean || !sq || !dev_row || D < 0 || updates < 1) return BDE_ERR_INVALID_ARG;
    if (D == 0) return BDE_OK;
    const bool vec = aligned16(theta) && aligned16(mean) && aligned16(sq) && aligned16(dev_row);
    const float fu = static_cast<float>(updates), fu1 = static_cast<float>(updates + 1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (use_ew_tma(D, vec, false)) {
        EwPtrs<3, 3> p;
        p.in[0] = theta;
        p.in[1] = mean;
        p.in[2] = sq;
        p.out[0] = mean;
        p.out[1] = sq;
        p.out[2] = dev_row;
        return launch_ew_tma<SwagUpdateOp>(p, D, SwagUpdateOp{fu, fu1}, nullptr, st);
    }
    int rc_;
    if (vec)
        rc_ = launch_ew(swag_update_kernel<true>, D, st, theta, mean, sq, dev_row, D, fu, fu1);
    else
        rc_ = launch_ew(swag_update_kernel<false>, D, st, theta, mean, sq, dev_row, D, fu, fu1);
    return rc_;
}

extern "C" int bde_swag_sample(const float* mean, const float* sq, const float* dev, int K, int head, int64_t D,
                               int64_t ld, const float* eps_k, const float* eps_d, uint64_t seed, uint64_t stream_id,
                               int64_t elem0, float* theta, bde_stream_t stream) {
    if (!mean || !sq || !dev || !theta || K < 1 || K > kMaxSwagRank || head < 0 || head >= K || D < 0 || ld < D ||
        elem0 < 0 || (elem0 & 3))
        return BDE_ERR_INVALID_ARG;
    if (D == 0) return BDE_OK;
    const bool vec = aligned16(mean) && aligned16(sq) && aligned16(dev) && aligned16(theta) && (ld % 4 == 0) &&
                     (!eps_d || aligned16(eps_d));
    const float den = static_cast<float>(sqrt(2.0 * (K - 1)));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc_;
    if (vec)
        rc_ = launch_ew(swag_sample_kernel<true, 5>, D, st, mean, sq, dev, K, head, D, ld, eps_k, eps_d, seed,
                                                                    stream_id, elem0 >> 2, den, theta);
    else
        rc_ = launch_ew(swag_sample_kernel<false, 5>, D, st, mean, sq, dev, K, head, D, ld, eps_k, eps_d, seed,
                                                                     stream_id, elem0 >> 2, den, theta);
    return rc_;
}

template <bool VEC>
static int launch_swag_batch(int sb, int64_t D, cudaStream_t st, const float* mean, const float* sq, const float* dev, int K,
                             int head, int64_t ld, int S, const float* eps_k, const float* eps_d, int64_t ld_eps,
                             uint64_t seed, uint64_t stream_id, int64_t quad0, float den, float* theta, int64_t ld_out) {
    switch (sb) {
        case 2:
            return launch_ew(swag_sample_batch_kernel<VEC, 2>, D, st, mean, sq, dev, K, head, D, ld, S, eps_k, eps_d, ld_eps, seed,
                             stream_id, quad0, den, theta, ld_out);
        case 4:
            return launch_ew(swag_sample_batch_kernel<VEC, 4>, D, st, mean, sq, dev, K, head, D, ld, S, eps_k, eps_d, ld_eps, seed,
                             stream_id, quad0, den, theta, ld_out);
        case 8:
            return launch_ew(swag_sample_batch_kernel<VEC, 8>, D, st, mean, sq, dev, K, head, D, ld, S, eps_k, eps_d, ld_eps, seed,
                             stream_id, quad0, den, theta, ld_out);
        default:
            return launch_ew(swag_sample_batch_kernel<VEC, 16>, D, st, mean, sq, dev, K, head, D, ld, S, eps_k, eps_d, ld_eps, seed,
                             stream_id, quad0, den, theta, ld_out);
    }
}

extern "C" int bde_swag_sample_batch(const float* mean, const float* sq, const float* dev, int K, int head, int64_t D,
                                     int64_t ld, int S, const float* eps_k, const float* eps_d, int64_t ld_eps,
                                     uint64_t seed, uint64_t stream_id, int64_t elem0, float* theta, int64_t ld_out,
                                     bde_stream_t stream) {
    if (!mean || !sq || !dev || !theta || K < 1 || K > kMaxSwagRank || head < 0 || head >= K || D < 0 || ld < D || S < 0 ||
        ld_out < D || (eps_d && ld_eps < D) || elem0 < 0 || (elem0 & 3))
        return BDE_ERR_INVALID_ARG;
    if (D == 0 || S == 0) return BDE_OK;
    const bool vec = aligned16(mean) && aligned16(sq) && aligned16(dev) && aligned16(theta) && (ld % 4 == 0) &&
                     (ld_out % 4 == 0) && (!eps_d || (aligned16(eps_d) && ld_eps % 4 == 0));
    const float den = static_cast<float>(sqrt(2.0 * (K - 1)));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tb = tuning().swag_batch;
    const int pf = tuning().batch_prefetch == 0 ? 1 : (tuning().batch_prefetch >= 9 ? 0 : tuning().batch_prefetch);
    const int per_pass = (tb == 2 || tb == 4 || tb == 8) ? tb : kSwagBatchMax;
    for (int s0 = 0; s0 < S; s0 += per_pass) {   // up to 16 draws per pass over the moments
        const int c = S - s0 < per_pass ? S - s0 : per_pass;
        const int sb = c <= 2 ? 2 : (c <= 4 ? 4 : (c <= 8 ? 8 : 16));
        const float* ek = eps_k ? eps_k + static_cast<int64_t>(s0) * K : nullptr;
        const float* ed = eps_d ? eps_d + s0 * ld_eps : nullptr;
        float* out = theta + s0 * ld_out;
        const int64_t nq = D >> 2;
        if (vec && c == sb && nq > 0 && tuning().swag_batch != 1) {   // whole passes of aligned data: fast kernel for the full quads ...
            int rf = BDE_OK;
            // bde_tune("batch_splits"): a pass can be split over 2 / 4 partner CTAs (fewer draws and registers per thread, shared
            // reads through L2).  Measured on B200 (profiles/r02_batch_samplers.jsonl): 16 draws x1 0.636 ms, x2 0.702, x4 0.835 —
            // the redundant per-quad work (+12 % instructions) costs more than 3 CTAs per SM gain, so the default is one CTA.
            int splits = tuning().batch_splits ? tuning().batch_splits : 1;
            while (splits > 1 && sb / splits < 2) splits >>= 1;
            const int sbk = sb / splits;
#define BDE_SWAG_FAST(SB_)                                                                                                     \
    rf = ed ? launch_swag_fast(swag_sample_batch_fast_kernel<SB_, true>, splits, nq, st, mean, sq, dev, K, head, nq, ld, ek, ed, ld_eps, \
                               seed, stream_id + s0, elem0 >> 2, den, out, ld_out, pf, splits)                                 \
            : launch_swag_fast(swag_sample_batch_fast_kernel<SB_, false>, splits, nq, st, mean, sq, dev, K, head, nq, ld, ek, ed, ld_eps, \
                               seed, stream_id + s0, elem0 >> 2, den, out, ld_out, pf, splits)
            switch (sbk) {
                case 2: BDE_SWAG_FAST(2); break;
                case 4: BDE_SWAG_FAST(4); break;
                case 8: BDE_SWAG_FAST(8); break;
                default: BDE_SWAG_FAST(16); break;
            }
#undef BDE_SWAG_FAST
            if (rf != BDE_OK) return rf;
            const int64_t d4 = nq << 2;
            if (d4 < D) {   // ... and the general kernel for the last D % 4 elements
                const int rt = launch_swag_batch<false>(sb, D - d4, st, mean + d4, sq + d4, dev + d4, K, head, ld, c, ek,
                                                        ed ? ed + d4 : nullptr, ld_eps, seed, stream_id + s0, (elem0 + d4) >> 2, den,
                                                        out + d4, ld_out);
                if (rt != BDE_OK) return rt;
            }
            continue;
        }
        const int rc_ = vec ? launch_swag_batch<true>(sb, D, st, mean, sq, dev, K, head, ld, c, ek, ed, ld_eps, seed,
                                                      stream_id + s0, elem0 >> 2, den, out, ld_out)
                            : launch_swag_batch<false>(sb, D, st, mean, sq, dev, K, head, ld, c, ek, ed, ld_eps, seed,
                                                       stream_id + s0, elem0 >> 2, den, out, ld_out);
        if (rc_ != BDE_OK) return rc_;
    }
    return BDE_OK;
}
