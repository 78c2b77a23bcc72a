// svgd_gram.cuh — K1 for n = 16 / 20 particles: pair distances from the CENTRED Gram matrix.
//
// Why: the direct form  d_ij = sum_c (x_ic - x_jc)^2  (svgd.py:15, torch.cdist for n <= 25) costs one FADD and one
// FFMA per pair and column: 380 fp32 lane operations per 80-byte column at n = 20.  A B200 issues 128 lanes x 148 SMs
// x 1.965 GHz = 37 T lane-ops/s, i.e. the direct form cannot stream faster than 7.8 TB/s at 100 % FMA-pipe
// utilisation; it measured 4.0 TB/s (profiles/r01_ncu_summary.md, session 39).  This kernel computes
//     y_r = x_r - x_0          (r = 1 .. n-1, one FADD per row and column)
//     S_ab = sum_c y_ac y_bc   (1 <= a <= b <= n-1, one FFMA per entry and column: n(n-1)/2 entries)
//     d_0j = S_jj,   d_ij = S_ii + S_jj - 2 S_ij
// which is ~(190 + 48) instead of 380 lane operations per column.  Subtracting particle 0 first removes the common
// offset of the particles, so the classic cancellation of the Gram form (|x|^2 >> |x_i - x_j|^2) does not occur as
// long as no pair is much closer to each other than to particle 0.  That condition is CHECKED on the device from
// the reduced sums (S_ii + S_jj <= kGramGuard * d_ij for every pair); when it fails the kernel raises the `redo`
// flag in the workspace header, leaves K / A untouched, and the direct kernel that the host has already enqueued
// behind it (it exits at once when the flag is clear) recomputes the exact distances.  Duplicated particles
// (d_ij = 0 exactly) therefore still give exact zeros.
//
// Mechanics: persistent CTA per SM, 8 consumer warps + a producer whose elected lane fetches the tile [n x 256
// columns] with ONE cp.async.bulk.tensor.2d (TMA tensor map over X[n, D]; out-of-range columns are zero-filled, so
// ragged ends need no extra code).  Measured on B200 (tools/microbench/tma_stream.cu, profiles/r02_tma_stream.jsonl):
// one lane issuing 20 separate 1 KB cp.async.bulk row copies per tile tops out at 2.9-4.2 TB/s (the copy instruction
// itself is the limit; the round-1 kernels at n = 16 / 20 sat exactly there), the tensor-map form streams 6.8-7.3
// TB/s.  Folding the TMA issue into a consumer warp was tried and starves the ring (4.0 TB/s: the issuing lane is
// blocked behind its own wait for data).  The n(n-1)/2 Gram entries are
// split over four warp groups: T(H1), T(H2), H1 x H2a, H1 x H2b (H1 / H2 = first / second half of the centred
// rows); a warp owns one group and 32 of the tile's 64 column quads.  Per-thread fp32 partial sums are flushed to
// per-warp fp64 accumulators every 64 tiles; CTA sums are combined by the deterministic last-CTA reduction (and,
// for a peer-attached workspace, across ranks), then the last CTA converts S to d, checks the guard and runs K1b.
#pragma once
#include "svgd_kernels.cuh"

namespace bde {

constexpr double kGramGuard = 32.0;      // max (S_ii + S_jj) / d_ij for which the Gram result is accepted
constexpr int kGramTileCols = 256;
constexpr int kGramConsumers = 256;      // 8 warps
constexpr int kGramFlushTiles = 64;
// automatic selection: below this many columns the direct kernels win — the Gram form carries ~20 us of fixed cost (tensor-map
// encode, the conditional exact-redo launch, the register hand-over): measured crossover 4-8 M columns at n = 16 and 20
// (profiles/r02_k1_small.jsonl, L2 flushed between launches)
constexpr int64_t kGramMinColumns = 6000000;

template <int N>
struct GramGeom {
    static constexpr int M = N - 1;             // centred rows y_1 .. y_{n-1}
    static constexpr int H1 = (M + 1) / 2;      // 10 at n = 20, 8 at n = 16
    static constexpr int H2 = M - H1;           //  9 /  7
    static constexpr int PA = (H2 + 1) / 2;     //  5 /  4   (H2a)
    static constexpr int PB = H2 - PA;          //  4 /  3   (H2b)
    static constexpr int E0 = H1 * (H1 + 1) / 2;
    static constexpr int E1 = H2 * (H2 + 1) / 2;
    static constexpr int E2 = H1 * PA;
    static constexpr int E3 = H1 * PB;
    static constexpr int OFF1 = E0, OFF2 = E0 + E1, OFF3 = E0 + E1 + E2;
    static constexpr int ENTRIES = E0 + E1 + E2 + E3;   // = M (M + 1) / 2 = pair_count(N)
    static constexpr int EMAX = E0 > E2 ? E0 : E2;
    static_assert(ENTRIES == M * (M + 1) / 2, "entry partition");
    static constexpr int STAGE_BYTES = N * kGramTileCols * 4;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 12 ? 12 : (200 * 1024) / STAGE_BYTES;
};
__host__ __device__ constexpr int tri_index(int a, int b, int r) { return a * r - a * (a - 1) / 2 + (b - a); }  // a <= b < r
// slot of S(i, j), 0 <= i <= j < M, in the flat entry list
template <int N>
__host__ __device__ constexpr int gram_slot(int i, int j) {
    using GG = GramGeom<N>;
    if (j < GG::H1) return tri_index(i, j, GG::H1);
    if (i >= GG::H1) return GG::OFF1 + tri_index(i - GG::H1, j - GG::H1, GG::H2);
    const int jj = j - GG::H1;
    return jj < GG::PA ? GG::OFF2 + i * GG::PA + jj : GG::OFF3 + i * GG::PB + (jj - GG::PA);
}
__host__ __device__ constexpr int gram_group_entries(int n, int g) {
    const int m = n - 1, h1 = (m + 1) / 2, h2 = m - h1, pa = (h2 + 1) / 2, pb = h2 - pa;
    return g == 0 ? h1 * (h1 + 1) / 2 : (g == 1 ? h2 * (h2 + 1) / 2 : (g == 2 ? h1 * pa : h1 * pb));
}

__device__ __forceinline__ V4 sub4(const V4& a, const V4& b) {
    V4 r;
    r.lo = sub2(a.lo, b.lo);
    r.hi = sub2(a.hi, b.hi);
    return r;
}
__device__ __forceinline__ void gram_fma(f32x2& acc, const V4& a, const V4& b) {
    acc = fma2(a.lo, b.lo, acc);
    acc = fma2(a.hi, b.hi, acc);
}

// One column quad of one tile for warp group G: `row(r)` returns the raw quad of particle r, `loaded()` is called
// once every row this group needs sits in registers (the ring stage is released there).
template <int N, int G, typename Row, typename Loaded>
__device__ __forceinline__ void gram_accumulate(Row&& row, Loaded&& loaded, f32x2 (&acc)[gram_group_entries(N, G)]) {
    using GG = GramGeom<N>;
    const V4 c = row(0);
    if constexpr (G <= 1) {
        constexpr int R = G == 0 ? GG::H1 : GG::H2;
        constexpr int BASE = G == 0 ? 1 : 1 + GG::H1;
        V4 y[R];
#pragma unroll
        for (int r = 0; r < R; ++r) y[r] = row(BASE + r);
        loaded();
#pragma unroll
        for (int r = 0; r < R; ++r) y[r] = sub4(y[r], c);
#pragma unroll
        for (int a = 0; a < R; ++a)
#pragma unroll
            for (int b = a; b < R; ++b) gram_fma(acc[tri_index(a, b, R)], y[a], y[b]);
    } else {
        constexpr int P = G == 2 ? GG::PA : GG::PB;
        constexpr int PBASE = 1 + GG::H1 + (G == 2 ? 0 : GG::PA);
        V4 p[P], u[GG::H1];
#pragma unroll
        for (int b = 0; b < P; ++b) p[b] = row(PBASE + b);
#pragma unroll
        for (int a = 0; a < GG::H1; ++a) u[a] = row(1 + a);
        loaded();
#pragma unroll
        for (int b = 0; b < P; ++b) p[b] = sub4(p[b], c);
#pragma unroll
        for (int a = 0; a < GG::H1; ++a) {
            const V4 ya = sub4(u[a], c);
#pragma unroll
            for (int b = 0; b < P; ++b) gram_fma(acc[a * P + b], ya, p[b]);
        }
    }
}

// warp -> (group, column half).  pairing 0 (default): the two warps of an SM sub-partition (w, w + 4) run a heavy and
// a light group (T(H1) with T(H2), H1xH2a with H1xH2b), which balances the FMA-pipe load of the four
// sub-partitions; 1 ("gram_pairing" tuning knob, A/B runs): both run the same group (one code body per
// sub-partition's instruction cache).  Warp 0 is in group 0 either way.
__device__ __forceinline__ void gram_warp_role(int wid, int pairing, int& g, int& half) {
    if (pairing == 0) {
        g = (wid & 1) * 2 + (wid >> 2);
        half = (wid >> 1) & 1;
    } else {
        g = wid & 3;
        half = wid >> 2;
    }
}

template <int N, int G>
__device__ __forceinline__ void gram_consumer(int64_t ntiles, const float* __restrict__ tiles, uint64_t* full_bar,
                                              uint64_t* empty_bar, double* __restrict__ wacc, int qi, int nst) {
    using GG = GramGeom<N>;
    constexpr int E = gram_group_entries(N, G);
    constexpr int TC = kGramTileCols;
    f32x2 acc[E];
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = 0ull;
    const int lane = threadIdx.x & 31;
    int s = 0;            // ring slot and its use count, advanced without div / mod
    uint32_t use = 0;
    int64_t t = blockIdx.x;
    bool more = true;
    while (more) {  // one flush site: every kGramFlushTiles tiles
        for (int f = 0; f < kGramFlushTiles && t < ntiles; ++f, t += gridDim.x, s = (s + 1 == nst) ? 0 : s + 1, use += (s == 0)) {
            mbar_wait(&full_bar[s], use & 1u);
            const float* sx = tiles + static_cast<size_t>(s) * N * TC + 4 * qi;
            gram_accumulate<N, G>([&](int r) { return lds_v4(sx + r * TC); },
                                  [&] {
                                      __syncwarp();
                                      if (lane == 0) mbar_arrive(&empty_bar[s]);
                                  },
                                  acc);
        }
        more = t < ntiles;
        flush_pairs<E>(acc, wacc, lane);
    }
}

template <int N, int G>
__device__ __forceinline__ void gram_dispatch(int g, int64_t ntiles, const float* tiles, uint64_t* full_bar,
                                              uint64_t* empty_bar, double* wacc, int qi, int nst) {
    if (g == G) {
        gram_consumer<N, G>(ntiles, tiles, full_bar, empty_bar, wacc, qi, nst);
    } else {
        if constexpr (G + 1 < 4) gram_dispatch<N, G + 1>(g, ntiles, tiles, full_bar, empty_bar, wacc, qi, nst);
    }
}

// S (reduced over CTAs and ranks, flat entry list) -> squared distances; returns in *redo whether some pair fails
// the cancellation guard.  Called by all threads of the last CTA; dist is written only for an accepted result.
template <int N>
__device__ __forceinline__ bool gram_to_dist(const double* __restrict__ total, double* __restrict__ dist, double guard,
                                             int tid, int nthreads) {
    __shared__ int s_redo;
    if (tid == 0) s_redo = 0;
    __syncthreads();
    for (int e = tid; e < N * N; e += nthreads) {
        const int i = e / N, j = e - i * N;
        if (i < j) {
            const double sjj = total[gram_slot<N>(j - 1, j - 1)];
            double d, mag;
            if (i == 0) {
                d = sjj;
                mag = sjj;
            } else {
                const double sii = total[gram_slot<N>(i - 1, i - 1)];
                d = sii + sjj - 2.0 * total[gram_slot<N>(i - 1, j - 1)];
                mag = sii + sjj;
            }
            if (!(mag <= guard * d)) s_redo = 1;   // also catches d <= 0 with mag > 0 and NaN
        }
    }
    __syncthreads();
    const bool redo = s_redo != 0;
    if (!redo) {
        for (int e = tid; e < N * N; e += nthreads) {
            const int i = e / N, j = e - i * N;
            double d = 0.0;
            if (i != j) {
                const int a = i < j ? i : j, b = i < j ? j : i;
                const double sbb = total[gram_slot<N>(b - 1, b - 1)];
                d = (a == 0) ? sbb : total[gram_slot<N>(a - 1, a - 1)] + sbb - 2.0 * total[gram_slot<N>(a - 1, b - 1)];
                if (d < 0.0) d = 0.0;
            }
            dist[e] = d;
        }
    }
    __syncthreads();
    return redo;
}

// Register budget.  A ninth (producer) warp puts three warps on one SM sub-partition, which caps every warp at 168
// registers; the n = 20 consumers want ~185.  SHIFT: the CTA is launched with a whole producer WARPGROUP (warps 8-11,
// only lane 0 of warp 8 works) at 168 registers per thread; the producer group hands its registers back
// (setmaxnreg.dec 40) and the two consumer groups grow to 232 (setmaxnreg.inc), the CUTLASS warp-specialisation
// pattern.  Before the common tail both sides return to 168.  !SHIFT: 8 consumer warps + one producer warp, plain.
constexpr int kGramRegsLaunch = 168, kGramRegsConsumer = 232, kGramRegsProducer = 40;
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

template <int N, bool SHIFT>
__global__ void __launch_bounds__(kGramConsumers + (SHIFT ? 128 : 32), 1)
svgd_pairgram_kernel(const __grid_constant__ CUtensorMap tmap, int64_t D, double* __restrict__ dist, void* ws,
                     int fuse_bandwidth, BandwidthParams bp, double guard, int pairing, int nst) {
    using GG = GramGeom<N>;
    constexpr int TC = kGramTileCols;
    constexpr int STAGES = GG::STAGES;
    constexpr int CWARPS = kGramConsumers / 32;
    constexpr int EN = GG::ENTRIES;
    constexpr int nthreads = kGramConsumers + (SHIFT ? 128 : 32);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);  // [STAGES][N][TC]
    __shared__ double wacc[CWARPS][GG::EMAX];
    __shared__ double cta_vals[EN];
    __shared__ double total[EN];
    __shared__ double sd[N * N];
    __shared__ double sk[N * N];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];

    const int tid = threadIdx.x;
    for (int k = tid; k < CWARPS * GG::EMAX; k += nthreads) (&wacc[0][0])[k] = 0.0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CWARPS);
        }
        mbar_fence_init();
        tma_prefetch_map(&tmap);
    }
    __syncthreads();

    const int64_t ntiles = (D + TC - 1) / TC;
    const int wid = tid >> 5;
    if (wid < CWARPS) {
        if constexpr (SHIFT) reg_inc<kGramRegsConsumer>();
        int g, half;
        gram_warp_role(wid, pairing, g, half);
        gram_dispatch<N, 0>(g, ntiles, tiles, full_bar, empty_bar, wacc[wid], half * 32 + (tid & 31), nst);
        if constexpr (SHIFT) reg_dec<kGramRegsLaunch>();
    } else {
        if constexpr (SHIFT) reg_dec<kGramRegsProducer>();
        if (tid == kGramConsumers) {   // the elected producer lane drives the TMA
            int s = 0;
            uint32_t use = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, s = (s + 1 == nst) ? 0 : s + 1, use += (s == 0)) {
                mbar_wait(&empty_bar[s], (use & 1u) ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], GG::STAGE_BYTES);
                tma_load_2d(tiles + static_cast<size_t>(s) * N * TC, &tmap, static_cast<int>(t * TC), 0, &full_bar[s]);
            }
        }
        __syncwarp();
        if constexpr (SHIFT) reg_inc<kGramRegsLaunch>();   // waits until the consumers have shrunk back
    }
    __syncthreads();
    griddep_launch();   // streaming done: dependents may become resident during the tail

    // entry e of group grp was accumulated by that group's two warps (one per column half)
    for (int e = tid; e < EN; e += nthreads) {
        const int grp = e < GG::OFF1 ? 0 : (e < GG::OFF2 ? 1 : (e < GG::OFF3 ? 2 : 3));
        const int k = e - (grp == 0 ? 0 : (grp == 1 ? GG::OFF1 : (grp == 2 ? GG::OFF2 : GG::OFF3)));
        double sacc = 0.0;
        for (int wv = 0; wv < CWARPS; ++wv) {
            int wg, wh;
            gram_warp_role(wv, pairing, wg, wh);
            if (wg == grp) sacc += wacc[wv][k];
        }
        cta_vals[e] = sacc;
    }
    __syncthreads();
    if (!grid_reduce_fp64(cta_vals, EN, ws, total)) return;
    const bool redo = gram_to_dist<N>(total, dist, guard, tid, nthreads);
    if (tid == 0) reinterpret_cast<WsHeader*>(ws)->redo = redo ? 1 : 0;
    if (!redo && fuse_bandwidth && !peer_exchange_failed(ws)) bandwidth_device<pow2_ceil(N * N)>(dist, N, bp, sd, sk);
}

// host side -------------------------------------------------------------------------------------------------------

template <int N>
int launch_pairgram(const float* X, int64_t D, int64_t ld, double* dist, void* ws, int fuse, const BandwidthParams& bp,
                    cudaStream_t st) {
    using GG = GramGeom<N>;
    CUtensorMap map;
    const int rc = encode_rows_tensor_map(&map, X, N, D, ld, kGramTileCols, tuning().gram_l2_promotion);
    if (rc != BDE_OK) return rc;
    constexpr int smem = GG::STAGES * GG::STAGE_BYTES;
    static bool configured = false;
    if (!configured) {
        BDE_RETURN_IF_CUDA(cudaFuncSetAttribute(svgd_pairgram_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        BDE_RETURN_IF_CUDA(cudaFuncSetAttribute(svgd_pairgram_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const double guard = tuning().gram_guard_x1000 > 0 ? 1e-3 * tuning().gram_guard_x1000 : kGramGuard;
    const int64_t ntiles = (D + kGramTileCols - 1) / kGramTileCols;
    int64_t grid = sm_count_cached();
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    int nst = GG::STAGES;
    if (tuning().ring_kb > 0) {
        const int want = tuning().ring_kb * 1024 / GG::STAGE_BYTES;
        if (want < nst) nst = want > 2 ? want : 2;
    }
    const bool shift = tuning().gram_fold == 0 ? (N > 16) : tuning().gram_fold == 1;   // knob: 1 = SHIFT, 2 = plain
    if (shift)
        svgd_pairgram_kernel<N, true><<<static_cast<unsigned>(grid), kGramConsumers + 128, smem, st>>>(map, D, dist, ws, fuse, bp, guard, tuning().gram_pairing, nst);
    else
        svgd_pairgram_kernel<N, false><<<static_cast<unsigned>(grid), kGramConsumers + 32, smem, st>>>(map, D, dist, ws, fuse, bp, guard, tuning().gram_pairing, nst);
    BDE_CHECK_LAUNCH();
    return BDE_OK;
}

}  // namespace bde
