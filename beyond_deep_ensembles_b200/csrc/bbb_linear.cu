// bbb_linear.cu — SURVEY.md §8 f4: the local-reparameterisation forward of the reference's BBBLinear
// (src/algos/bbb_layers.py:61-88, CUDA branch) as ONE tensor-core kernel.
//
// Reference arithmetic (bbb_layers.py:66-79):
//     mean = b_mu + x  @ W_mu^T
//     var  = clamp(softplus(b_rho)^2, 1e-4) + clamp(x^2, 1e-4) @ clamp(softplus(W_rho)^2, 1e-4)^T
//     out  = (mean + sqrt(var) * eps) / mc_sample
// which eager PyTorch runs as ~12 launches (3 stacks, pow, 2 clamps, softplus, baddbmm, sqrt, normal_, mul, add).
//
// Here: D^T[o, b] = sum_k W[o, k] x[b, k] on the 5th-generation tensor cores.  A CTA owns one [128 out-features x 32 k]
// block of the weights (grid = out/128 x in/32 x batch tiles: 144 CTAs at the CivilComments head 768 x 768) and
//   1. fetches its raw W_mu / W_rho / x tiles with three tensor-map TMA loads (128-byte swizzle, the layout the MMA
//      wants; out-of-range rows / columns arrive as zeros),
//   2. rewrites them IN PLACE (element-wise, so the swizzle never has to be computed) into tf32 operand triples:
//      v = h + m + l with h = the top 11 significand bits of v, m = the next 11, l = the rest — each EXACTLY a tf32
//      number, so every tensor-core product is exact and the sum h*h + h*m + m*h + m*m + h*l + l*h carries the fp32
//      product to 2^-33 (a plain hi / lo "3xTF32" split truncates lo and measured 2.4e-6 absolute at K = 768, outside
//      the 1e-5 / 1e-6 tolerance) — for W_mu, for sigma^2 = clamp(softplus(W_rho)^2, 1e-4), for x and clamp(x^2, 1e-4),
//   3. issues 48 tcgen05.mma (kind::tf32, M = 128, N = batch tile, K = 8) from one elected thread into two TMEM
//      accumulators (mean and variance), tcgen05.commit on an mbarrier,
//   4. reads the accumulators back with tcgen05.ld and stores the split-K partial tile to the workspace; the LAST CTA
//      of an out-feature tile (ticket) adds the in/32 partials in fixed order (fp64, deterministic) and applies the
//      bias, sqrt and noise epilogue (injected eps or Philox keyed by the element index).
// The backward pass stays in PyTorch (plain library GEMMs on the saved activations; util.py).
#include <cuda.h>

#include "common.cuh"

namespace bde {

constexpr int kBlM = 128;          // out-features per CTA (UMMA M)
constexpr int kBlK = 32;           // k per CTA = one 128-byte swizzle row of fp32
constexpr int kBlThreads = 128;    // 4 warps: TMEM lanes 32 w .. 32 w + 31
constexpr int kBlMaxNB = 128;      // batch tile (UMMA N), multiple of 16

__device__ __forceinline__ void tma_load_2d_sw(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    tma_load_2d(smem_dst, map, c0, c1, bar);
}
__device__ __forceinline__ uint64_t umma_desc_k_sw128(const void* smem_tile, int k_byte_offset) {
    // K-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14), leading byte
    // offset (unused for swizzled K-major, canonical value 1) in [16,30), stride byte offset = 8 rows x 128 B = 1024 B
    // (>> 4) in [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
    const uint32_t addr = smem_u32(smem_tile) + static_cast<uint32_t>(k_byte_offset);
    uint64_t d = static_cast<uint64_t>((addr >> 4) & 0x3fffu);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    // cute::UMMA::InstrDescriptor: D format F32 (1) in [4,6), A / B format TF32 (2) in [7,10) / [10,13), both K-major,
    // N >> 3 in [17,23), M >> 4 in [24,29)
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ float4 hi4(float4 v) { return make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w)); }
__device__ __forceinline__ float4 sub4f(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
// v = h + m + l, each exactly representable in tf32 (11 significand bits); written at chunk c of the three tiles
__device__ __forceinline__ void split3(float4 v, float* const (&tiles)[3], int c) {
    const float4 h = hi4(v);
    const float4 r = sub4f(v, h);
    const float4 m = hi4(r);
    reinterpret_cast<float4*>(tiles[0])[c] = h;
    reinterpret_cast<float4*>(tiles[1])[c] = m;
    reinterpret_cast<float4*>(tiles[2])[c] = sub4f(r, m);
}
__device__ __forceinline__ float var_of_rho(float rho) {   // clamp(softplus(rho)^2, 1e-4), bbb_layers.py:69
    const float s = softplus_ref(rho);
    return fmaxf(__fmul_rn(s, s), 1e-4f);
}

struct BblParams {
    const float* b_mu;      // [out] or null
    const float* b_rho;     // [out] or null
    const float* eps;       // [batch, out] injected noise or null (Philox)
    float* out;             // [batch, out]
    float* act_std;         // [batch, out] or null: sqrt(var), saved for the backward pass
    float* eps_out;         // [batch, out] or null: the noise that was used
    float* partials;        // workspace: [mtiles][btiles][ksplits][2][NB][128]
    unsigned int* tickets;  // workspace: [mtiles * btiles], zero between launches
    int batch, in_features, out_features, nb, ksplits;
    int two_phase;          // 1: partials only, bbb_linear_reduce_kernel sums them (many k splits)
    float mc;               // BBBLinear.mc_sample: the output is divided by it (bbb_layers.py:88)
    uint64_t seed, stream_id;
};

// fixed-order sum over the k splits of one (out-feature tile, batch tile), batch columns [bc_begin, bc_end), then the
// epilogue; thread = out-feature row.  Called by the last CTA of the tile (few splits) or by the reduce kernel (many)
template <int NB, int BC>
__device__ __forceinline__ void bbl_reduce_epilogue(const BblParams& p, int mt, int bt, int nbt, int tid, int bc_begin, int bc_end) {
    const int o0 = mt * kBlM, b0 = bt * NB, o = o0 + tid;
    const float* base = p.partials + ((static_cast<size_t>(mt) * nbt + bt) * p.ksplits * 2) * NB * kBlM;
    float bmean = 0.f, bvar = 0.f;
    if (o < p.out_features && p.b_mu) {
        bmean = p.b_mu[o];
        bvar = var_of_rho(p.b_rho[o]);
    }
    // BC batch columns reduced together: 2 x BC independent loads in flight per k split
    for (int bc = bc_begin; bc < bc_end && b0 + bc < p.batch; bc += BC) {
        double sm[BC], sv[BC];
#pragma unroll
        for (int j = 0; j < BC; ++j) sm[j] = sv[j] = 0.0;
        // four k splits per round: 4 x 2 x 16 independent L2 loads in flight per thread, summed in split order
        for (int s0 = 0; s0 < p.ksplits; s0 += 4) {
            float vm[4][BC], vv[4][BC];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool on = s0 + u < p.ksplits;
                const float* ps = base + static_cast<size_t>(on ? s0 + u : s0) * 2 * NB * kBlM + tid;
#pragma unroll
                for (int j = 0; j < BC; ++j) {
                    vm[u][j] = on ? __ldcg(ps + static_cast<size_t>(bc + j) * kBlM) : 0.f;
                    vv[u][j] = on ? __ldcg(ps + (static_cast<size_t>(NB) + bc + j) * kBlM) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < BC; ++j) {
                    sm[j] += static_cast<double>(vm[u][j]);
                    sv[j] += static_cast<double>(vv[u][j]);
                }
        }
#pragma unroll
        for (int j = 0; j < BC; ++j) {
            const int b = bc + j;
            if (o < p.out_features && b0 + b < p.batch) {
                const float mean = __fadd_rn(bmean, static_cast<float>(sm[j]));          // baddbmm: add + matmul
                const float sd = __fsqrt_rn(__fadd_rn(bvar, static_cast<float>(sv[j])));
                const int64_t e = static_cast<int64_t>(b0 + b) * p.out_features + o;
                float z;
                if (p.eps) {
                    z = p.eps[e];
                } else {
                    const float4 z4 = philox_normal4(p.seed, p.stream_id, static_cast<uint64_t>(e >> 2));
                    z = (e & 3) == 0 ? z4.x : (e & 3) == 1 ? z4.y : (e & 3) == 2 ? z4.z : z4.w;
                }
                p.out[e] = __fdiv_rn(__fadd_rn(mean, __fmul_rn(sd, z)), p.mc);
                if (p.act_std) p.act_std[e] = sd;
                if (p.eps_out) p.eps_out[e] = z;
            }
        }
    }
}

template <int NB>
__global__ void __launch_bounds__(kBlThreads, 1)
bbb_linear_fwd_kernel(const __grid_constant__ CUtensorMap map_wmu, const __grid_constant__ CUtensorMap map_wrho,
                      const __grid_constant__ CUtensorMap map_x, const __grid_constant__ BblParams p) {
    // operand tiles, 1024-byte aligned (128-byte swizzle atoms): W_mu hi / lo, sigma^2 hi / lo: [128][32]; x hi / lo, x^2 hi / lo: [NB][32]
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // launch reserves 1 KB of slack
    float* a_mu[3];   // W_mu      as h / m / l tiles [128][32]
    float* a_v[3];    // sigma^2   as h / m / l
    float* b_x[3];    // x         as h / m / l tiles [NB][32]
    float* b_q[3];    // clamp(x^2)
    {
        float* t = reinterpret_cast<float*>(smem);
#pragma unroll
        for (int i = 0; i < 3; ++i) a_mu[i] = t + i * kBlM * kBlK;
#pragma unroll
        for (int i = 0; i < 3; ++i) a_v[i] = t + (3 + i) * kBlM * kBlK;
        t += 6 * kBlM * kBlK;
#pragma unroll
        for (int i = 0; i < 3; ++i) b_x[i] = t + i * NB * kBlK;
#pragma unroll
        for (int i = 0; i < 3; ++i) b_q[i] = t + (3 + i) * NB * kBlK;
    }
    __shared__ __align__(8) uint64_t load_bar, mma_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ bool is_last;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int mt = blockIdx.x, ks = blockIdx.y, bt = blockIdx.z;
    const int o0 = mt * kBlM, b0 = bt * NB;
    constexpr int TMEM_COLS = 2 * NB < 32 ? 32 : 2 * NB;   // power of two >= 32 for NB in {16, 32, 64, 128}

    if (tid == 0) {
        mbar_init(&load_bar, 1);
        mbar_init(&mma_bar, 1);
        mbar_fence_init();
        tma_prefetch_map(&map_wmu);
        tma_prefetch_map(&map_wrho);
        tma_prefetch_map(&map_x);
    }
    if (warp == 0) {   // one warp allocates the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    // this CTA's k blocks: ks, ks + ksplits, ... (one block when in / 32 <= ksplits, the CivilComments head case)
    const int nkb = (p.in_features + kBlK - 1) / kBlK;
    int it = 0;
    for (int kb_idx = ks; kb_idx < nkb; kb_idx += p.ksplits, ++it) {
    const int k0 = kb_idx * kBlK;
    const uint32_t phase = static_cast<uint32_t>(it & 1);
    if (tid == 0) {   // raw tiles: W_mu -> a_mu[0], W_rho -> a_v[0], x -> b_x[0] (rewritten in place below)
        mbar_arrive_expect_tx(&load_bar, (2 * kBlM + NB) * kBlK * 4);
        tma_load_2d_sw(a_mu[0], &map_wmu, k0, o0, &load_bar);
        tma_load_2d_sw(a_v[0], &map_wrho, k0, o0, &load_bar);
        tma_load_2d_sw(b_x[0], &map_x, k0, b0, &load_bar);
    }
    mbar_wait(&load_bar, phase);

    // element-wise rewrite at the SAME (swizzled) positions.  chunk c = 16 bytes at tile offset 16 c; its row is c / 8
    // and its logical k-chunk is (c % 8) ^ (row % 8) (128-byte swizzle), needed only to mask k >= in_features
    const bool ragged_k = k0 + kBlK > p.in_features;
#pragma unroll 2
    for (int c = tid; c < kBlM * 8; c += kBlThreads) {
        float4 mu = reinterpret_cast<float4*>(a_mu[0])[c];
        float4 rho = reinterpret_cast<float4*>(a_v[0])[c];
        float4 var = make_float4(var_of_rho(rho.x), var_of_rho(rho.y), var_of_rho(rho.z), var_of_rho(rho.w));
        const int row = c >> 3;
        const bool row_ok = o0 + row < p.out_features;
        if (ragged_k || !row_ok) {
            const int kc = ((c & 7) ^ (row & 7)) * 4 + k0;
            const float m0 = (row_ok && kc + 0 < p.in_features) ? 1.f : 0.f, m1 = (row_ok && kc + 1 < p.in_features) ? 1.f : 0.f;
            const float m2 = (row_ok && kc + 2 < p.in_features) ? 1.f : 0.f, m3 = (row_ok && kc + 3 < p.in_features) ? 1.f : 0.f;
            mu = make_float4(mu.x * m0, mu.y * m1, mu.z * m2, mu.w * m3);
            var = make_float4(var.x * m0, var.y * m1, var.z * m2, var.w * m3);
        }
        split3(mu, a_mu, c);
        split3(var, a_v, c);
    }
    for (int c = tid; c < NB * 8; c += kBlThreads) {
        float4 x = reinterpret_cast<float4*>(b_x[0])[c];
        const int row = c >> 3;
        float4 q = make_float4(fmaxf(x.x * x.x, 1e-4f), fmaxf(x.y * x.y, 1e-4f), fmaxf(x.z * x.z, 1e-4f), fmaxf(x.w * x.w, 1e-4f));
        if (ragged_k) {   // clamp(x^2) of a zero-filled column is 1e-4, not 0
            const int kc = ((c & 7) ^ (row & 7)) * 4 + k0;
            q = make_float4(kc + 0 < p.in_features ? q.x : 0.f, kc + 1 < p.in_features ? q.y : 0.f,
                            kc + 2 < p.in_features ? q.z : 0.f, kc + 3 < p.in_features ? q.w : 0.f);
        }
        split3(x, b_x, c);
        split3(q, b_q, c);
    }
    // generic-proxy writes -> visible to the tensor core's async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (tid == 0) {   // one thread issues every MMA of the CTA: 4 k-steps of 8 x (6 + 6) exact tf32 products
        constexpr uint32_t idesc = umma_idesc_tf32(kBlM, NB);
        const uint32_t d_mean = tmem_base, d_var = tmem_base + NB;
#pragma unroll
        for (int kk = 0; kk < kBlK / 8; ++kk) {
            const int kb = kk * 32;   // 8 tf32 = 32 bytes further along the swizzled row
            // (A part, B part): the six products that matter, smallest first
            constexpr int PA[6] = {2, 0, 1, 1, 0, 0}, PB[6] = {0, 2, 1, 0, 1, 0};
#pragma unroll
            for (int t = 0; t < 6; ++t) {
                umma_tf32(d_mean, umma_desc_k_sw128(a_mu[PA[t]], kb), umma_desc_k_sw128(b_x[PB[t]], kb), idesc, (it | kk | t) > 0);
                umma_tf32(d_var, umma_desc_k_sw128(a_v[PA[t]], kb), umma_desc_k_sw128(b_q[PB[t]], kb), idesc, (it | kk | t) > 0);
            }
        }
        umma_commit(&mma_bar);   // arrives when every MMA above has written TMEM (implies fence::before_thread_sync)
    }
    // every MMA of this block has read its operands (and written TMEM) before the tiles are overwritten / read back
    mbar_wait(&mma_bar, phase);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }   // k blocks
    if (p.two_phase) griddep_launch();   // the reduce kernel's CTAs may take their places; they wait for this grid to finish

    // accumulators -> split-K partial tile in the workspace: [2][NB][128], thread = out-feature row (TMEM lane)
    float* part = p.partials + (((static_cast<size_t>(mt) * gridDim.z + bt) * p.ksplits + ks) * 2) * NB * kBlM;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
    for (int which = 0; which < 2; ++which) {
#pragma unroll
        for (int c0 = 0; c0 < NB; c0 += 16) {
            float v[16];
            tmem_ld16(lane_base + which * NB + c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) part[(static_cast<size_t>(which) * NB + c0 + j) * kBlM + tid] = v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __threadfence();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    if (p.two_phase) return;   // the reduce kernel (launched behind this grid) sums the partial tiles
    if (tid == 0) {
        const unsigned int prev = atomicAdd(&p.tickets[mt * gridDim.z + bt], 1u);
        is_last = prev == static_cast<unsigned int>(p.ksplits) - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    bbl_reduce_epilogue<NB, 16>(p, mt, bt, gridDim.z, tid, 0, NB);
    if (tid == 0) p.tickets[mt * gridDim.z + bt] = 0u;   // workspace reusable by the next launch
}

// ------------------------------------------------------------------------------------------------------------------
// Rank-1 VI linear layer (src/algos/rank1.py:50-64): out = (linear(x * s, W) * r) + bias with s = mu_s + eps_s *
// softplus(rho_s) (one value per INPUT feature) and r likewise per OUTPUT feature — the prologue / epilogue scaling of
// SURVEY §8 f4 fused around the same tensor-core product: s is sampled and multiplied into the x operand while the tile is
// rewritten into its tf32 triples, r and the bias are applied in the fixed-order split-K epilogue.
// ------------------------------------------------------------------------------------------------------------------
struct R1Params {
    const float* s_mu;      // [in]
    const float* s_rho;
    const float* r_mu;      // [out]
    const float* r_rho;
    const float* bias;      // [out] or null
    const float* eps_s;     // [in] injected or null (Philox stream sid_s)
    const float* eps_r;     // [out] injected or null (Philox stream sid_r)
    float* out;             // [batch, out]
    float* lin;             // [batch, out]: linear(x * s, W) before r and bias (backward needs it)
    float* s_out;           // [in], [out]: the sampled vectors and the noise they used
    float* r_out;
    float* eps_s_out;
    float* eps_r_out;
    float* partials;        // workspace: [mtiles][btiles][ksplits][NB][128]
    unsigned int* tickets;
    int batch, in_features, out_features, ksplits, two_phase;
    uint64_t seed, sid_s, sid_r;
};

__device__ __forceinline__ float philox_normal1(uint64_t seed, uint64_t stream_id, int64_t e) {
    const float4 z4 = philox_normal4(seed, stream_id, static_cast<uint64_t>(e >> 2));
    return (e & 3) == 0 ? z4.x : (e & 3) == 1 ? z4.y : (e & 3) == 2 ? z4.z : z4.w;
}

template <int NB, int BC>
__device__ __forceinline__ void r1_reduce_epilogue(const R1Params& p, int mt, int bt, int nbt, int tid, int bc_begin, int bc_end) {
    const int o0 = mt * kBlM, b0 = bt * NB, o = o0 + tid;
    const float* base = p.partials + (static_cast<size_t>(mt) * nbt + bt) * p.ksplits * NB * kBlM;
    float rv = 0.f, bias = 0.f;
    if (o < p.out_features) {
        const float e = p.eps_r ? p.eps_r[o] : philox_normal1(p.seed, p.sid_r, o);
        rv = __fadd_rn(p.r_mu[o], __fmul_rn(e, softplus_ref(p.r_rho[o])));
        if (p.bias) bias = p.bias[o];
        if (bt == 0 && bc_begin == 0) {
            p.r_out[o] = rv;
            p.eps_r_out[o] = e;
        }
    }
    for (int bc = bc_begin; bc < bc_end && b0 + bc < p.batch; bc += BC) {
        double sm[BC];
#pragma unroll
        for (int j = 0; j < BC; ++j) sm[j] = 0.0;
        for (int s0 = 0; s0 < p.ksplits; s0 += 4) {
            float vm[4][BC];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool on = s0 + u < p.ksplits;
                const float* ps = base + static_cast<size_t>(on ? s0 + u : s0) * NB * kBlM + tid;
#pragma unroll
                for (int j = 0; j < BC; ++j) vm[u][j] = on ? __ldcg(ps + static_cast<size_t>(bc + j) * kBlM) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < BC; ++j) sm[j] += static_cast<double>(vm[u][j]);
        }
#pragma unroll
        for (int j = 0; j < BC; ++j) {
            const int b = bc + j;
            if (o < p.out_features && b0 + b < p.batch) {
                const int64_t e = static_cast<int64_t>(b0 + b) * p.out_features + o;
                const float lin = static_cast<float>(sm[j]);
                p.lin[e] = lin;
                const float y = __fmul_rn(lin, rv);                       // self.layer(input * s) * r
                p.out[e] = p.bias ? __fadd_rn(y, bias) : y;               // output += bias[component]
            }
        }
    }
}

template <int NB>
__global__ void __launch_bounds__(kBlThreads, 1)
rank1_linear_fwd_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                        const __grid_constant__ R1Params p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* a_w[3];    // W as h / m / l tiles [128][32]
    float* b_x[3];    // x * s as h / m / l tiles [NB][32]
    {
        float* t = reinterpret_cast<float*>(smem);
#pragma unroll
        for (int i = 0; i < 3; ++i) a_w[i] = t + i * kBlM * kBlK;
        t += 3 * kBlM * kBlK;
#pragma unroll
        for (int i = 0; i < 3; ++i) b_x[i] = t + i * NB * kBlK;
    }
    __shared__ __align__(8) uint64_t load_bar, mma_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ bool is_last;
    __shared__ float s_blk[kBlK];

    const int tid = threadIdx.x, warp = tid >> 5;
    const int mt = blockIdx.x, ks = blockIdx.y, bt = blockIdx.z;
    const int o0 = mt * kBlM, b0 = bt * NB;
    constexpr int TMEM_COLS = NB < 32 ? 32 : NB;

    if (tid == 0) {
        mbar_init(&load_bar, 1);
        mbar_init(&mma_bar, 1);
        mbar_fence_init();
        tma_prefetch_map(&map_w);
        tma_prefetch_map(&map_x);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    const int nkb = (p.in_features + kBlK - 1) / kBlK;
    int it = 0;
    for (int kb_idx = ks; kb_idx < nkb; kb_idx += p.ksplits, ++it) {
        const int k0 = kb_idx * kBlK;
        const uint32_t phase = static_cast<uint32_t>(it & 1);
        if (tid == 0) {
            mbar_arrive_expect_tx(&load_bar, (kBlM + NB) * kBlK * 4);
            tma_load_2d_sw(a_w[0], &map_w, k0, o0, &load_bar);
            tma_load_2d(b_x[0], &map_x, k0, b0, &load_bar);
        }
        if (tid < kBlK) {   // s for this k block: mean + eps * softplus(rho) (util.py:170-171), 0 beyond in_features
            const int k = k0 + tid;
            float sv = 0.f;
            if (k < p.in_features) {
                const float e = p.eps_s ? p.eps_s[k] : philox_normal1(p.seed, p.sid_s, k);
                sv = __fadd_rn(p.s_mu[k], __fmul_rn(e, softplus_ref(p.s_rho[k])));
                if (mt == 0 && bt == 0) {
                    p.s_out[k] = sv;
                    p.eps_s_out[k] = e;
                }
            }
            s_blk[tid] = sv;
        }
        mbar_wait(&load_bar, phase);
        __syncthreads();   // s_blk visible
        for (int c = tid; c < kBlM * 8; c += kBlThreads) split3(reinterpret_cast<float4*>(a_w[0])[c], a_w, c);
        for (int c = tid; c < NB * 8; c += kBlThreads) {
            const float4 x = reinterpret_cast<float4*>(b_x[0])[c];
            const int kc = ((c & 7) ^ ((c >> 3) & 7)) * 4;   // logical k of this swizzled chunk inside the block
            const float4 xs = make_float4(__fmul_rn(x.x, s_blk[kc]), __fmul_rn(x.y, s_blk[kc + 1]), __fmul_rn(x.z, s_blk[kc + 2]),
                                          __fmul_rn(x.w, s_blk[kc + 3]));   // input * s, rounded to fp32 like the reference
            split3(xs, b_x, c);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(kBlM, NB);
#pragma unroll
            for (int kk = 0; kk < kBlK / 8; ++kk) {
                const int kb = kk * 32;
                constexpr int PA[6] = {2, 0, 1, 1, 0, 0}, PB[6] = {0, 2, 1, 0, 1, 0};
#pragma unroll
                for (int t = 0; t < 6; ++t)
                    umma_tf32(tmem_base, umma_desc_k_sw128(a_w[PA[t]], kb), umma_desc_k_sw128(b_x[PB[t]], kb), idesc, (it | kk | t) > 0);
            }
            umma_commit(&mma_bar);
        }
        mbar_wait(&mma_bar, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (p.two_phase) griddep_launch();

    float* part = p.partials + ((static_cast<size_t>(mt) * gridDim.z + bt) * p.ksplits + ks) * NB * kBlM;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < NB; c0 += 16) {
        float v[16];
        tmem_ld16(lane_base + c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) part[static_cast<size_t>(c0 + j) * kBlM + tid] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __threadfence();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    if (p.two_phase) return;   // the reduce kernel (launched behind this grid) sums the partial tiles
    if (tid == 0) {
        const unsigned int prev = atomicAdd(&p.tickets[mt * gridDim.z + bt], 1u);
        is_last = prev == static_cast<unsigned int>(p.ksplits) - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    r1_reduce_epilogue<NB, 16>(p, mt, bt, gridDim.z, tid, 0, NB);
    if (tid == 0) p.tickets[mt * gridDim.z + bt] = 0u;
}

// second phase for many k splits: grid (out tiles, NB / 4 column groups, batch tiles) — the partial tiles (L2-resident) are
// summed by many small CTAs instead of one CTA per tile; launched programmatically behind the product kernel
constexpr int kBlReduceCols = 4;
template <int NB>
__global__ void __launch_bounds__(kBlThreads) bbb_linear_reduce_kernel(const __grid_constant__ BblParams p) {
    griddep_wait();
    bbl_reduce_epilogue<NB, kBlReduceCols>(p, blockIdx.x, blockIdx.z, gridDim.z, threadIdx.x, blockIdx.y * kBlReduceCols,
                                           (blockIdx.y + 1) * kBlReduceCols);
}
template <int NB>
__global__ void __launch_bounds__(kBlThreads) rank1_linear_reduce_kernel(const __grid_constant__ R1Params p) {
    griddep_wait();
    r1_reduce_epilogue<NB, kBlReduceCols>(p, blockIdx.x, blockIdx.z, gridDim.z, threadIdx.x, blockIdx.y * kBlReduceCols,
                                          (blockIdx.y + 1) * kBlReduceCols);
}

// tensor map over a row-major fp32 matrix [rows, cols] (row stride ld), box = 32 columns x box_rows, 128-byte swizzle
static int encode_sw128_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

}  // namespace bde

using namespace bde;

#include <cudaTypedefs.h>

static int bde::encode_sw128_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        BDE_RETURN_IF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) return BDE_ERR_INVALID_ARG;
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(float)};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlK), static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? BDE_OK : BDE_ERR_INVALID_ARG;
}

static int bbl_nb(int batch) { return batch <= 16 ? 16 : (batch <= 32 ? 32 : (batch <= 64 ? 64 : 128)); }

// split-K factor: one k block per CTA while that helps to fill the GPU (~two CTAs per SM in total), otherwise a CTA walks
// several k blocks.  Up to kBlTicketSplits splits the last CTA of a tile sums the partial tiles itself (ticket); beyond
// that the sum is a second, programmatically launched kernel spread over (tiles x NB / 4) CTAs — one CTA reading
// ksplits x 2 x NB x 128 floats through its own L2 port was the dominant cost at batch >= 64
constexpr int kBlTicketSplits = 4;
static int bbl_ksplits(int batch, int in_features, int out_features) {
    const int nb = bbl_nb(batch);
    const int nkb = (in_features + kBlK - 1) / kBlK;
    const int mtiles = (out_features + kBlM - 1) / kBlM, btiles = (batch + nb - 1) / nb;
    int ks = (2 * 148 + mtiles * btiles - 1) / (mtiles * btiles);
    if (ks > nkb) ks = nkb;
    return ks < 1 ? 1 : ks;
}

template <typename Params>
static int launch_reduce(void (*kernel)(const Params), const Params& p, dim3 grid, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kBlThreads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    BDE_RETURN_IF_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
    return BDE_OK;
}

extern "C" int bde_bbb_linear_workspace_bytes(int batch, int in_features, int out_features, size_t* bytes) {
    if (!bytes || batch < 1 || in_features < 1 || out_features < 1) return BDE_ERR_INVALID_ARG;
    const int nb = bbl_nb(batch);
    const size_t mtiles = (out_features + kBlM - 1) / kBlM, btiles = (batch + nb - 1) / nb, ks = bbl_ksplits(batch, in_features, out_features);
    *bytes = 256 + ((mtiles * btiles * sizeof(unsigned int) + 255) & ~static_cast<size_t>(255)) +
             mtiles * btiles * ks * 2 * nb * kBlM * sizeof(float);
    return BDE_OK;
}

template <int NB>
static int launch_bbl(const CUtensorMap& mw, const CUtensorMap& mr, const CUtensorMap& mx, const BblParams& p, dim3 grid, cudaStream_t st) {
    constexpr int smem = (6 * kBlM + 6 * NB) * kBlK * 4 + 1024;
    static bool configured = false;
    if (!configured) {
        BDE_RETURN_IF_CUDA(cudaFuncSetAttribute(bbb_linear_fwd_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    bbb_linear_fwd_kernel<NB><<<grid, kBlThreads, smem, st>>>(mw, mr, mx, p);
    BDE_CHECK_LAUNCH();
    if (p.two_phase) return launch_reduce(bbb_linear_reduce_kernel<NB>, p, dim3(grid.x, NB / kBlReduceCols, grid.z), st);
    return BDE_OK;
}

extern "C" int bde_bbb_linear_fwd(const float* x, int64_t ldx, int batch, int in_features, int out_features, const float* w_mu,
                                  const float* w_rho, const float* b_mu, const float* b_rho, const float* eps, uint64_t seed,
                                  uint64_t stream_id, double mc_sample, float* out, float* act_std, float* eps_out,
                                  void* workspace, size_t workspace_bytes, bde_stream_t stream) {
    if (!x || !w_mu || !w_rho || !out || !workspace || batch < 1 || in_features < 1 || out_features < 1 || ldx < in_features ||
        (b_mu == nullptr) != (b_rho == nullptr))
        return BDE_ERR_INVALID_ARG;
    if (!aligned16(x) || !aligned16(w_mu) || !aligned16(w_rho) || (ldx % 4) != 0 || (in_features % 4) != 0) return BDE_ERR_ALIGNMENT;
    size_t need = 0;
    bde_bbb_linear_workspace_bytes(batch, in_features, out_features, &need);
    if (workspace_bytes < need) return BDE_ERR_WORKSPACE;
    const int nb = bbl_nb(batch);
    const int mtiles = (out_features + kBlM - 1) / kBlM, btiles = (batch + nb - 1) / nb;
    const int ks = bbl_ksplits(batch, in_features, out_features);
    CUtensorMap mw, mr, mx;
    int rc = encode_sw128_map(&mw, w_mu, out_features, in_features, in_features, kBlM);
    if (rc == BDE_OK) rc = encode_sw128_map(&mr, w_rho, out_features, in_features, in_features, kBlM);
    if (rc == BDE_OK) rc = encode_sw128_map(&mx, x, batch, in_features, ldx, nb);
    if (rc != BDE_OK) return rc;
    BblParams p{};
    p.b_mu = b_mu, p.b_rho = b_rho, p.eps = eps, p.out = out, p.act_std = act_std, p.eps_out = eps_out;
    p.tickets = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + 256);
    p.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256 +
                                          ((static_cast<size_t>(mtiles) * btiles * sizeof(unsigned int) + 255) & ~static_cast<size_t>(255)));
    p.batch = batch, p.in_features = in_features, p.out_features = out_features, p.nb = nb, p.ksplits = ks;
    p.two_phase = ks > kBlTicketSplits;
    p.mc = static_cast<float>(mc_sample > 0.0 ? mc_sample : 1.0);
    p.seed = seed, p.stream_id = stream_id;
    const dim3 grid(mtiles, ks, btiles);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (nb) {
        case 16: return launch_bbl<16>(mw, mr, mx, p, grid, st);
        case 32: return launch_bbl<32>(mw, mr, mx, p, grid, st);
        case 64: return launch_bbl<64>(mw, mr, mx, p, grid, st);
        default: return launch_bbl<128>(mw, mr, mx, p, grid, st);
    }
}

template <int NB>
static int launch_r1(const CUtensorMap& mw, const CUtensorMap& mx, const R1Params& p, dim3 grid, cudaStream_t st) {
    constexpr int smem = (3 * kBlM + 3 * NB) * kBlK * 4 + 1024;
    static bool configured = false;
    if (!configured) {
        BDE_RETURN_IF_CUDA(cudaFuncSetAttribute(rank1_linear_fwd_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    rank1_linear_fwd_kernel<NB><<<grid, kBlThreads, smem, st>>>(mw, mx, p);
    BDE_CHECK_LAUNCH();
    if (p.two_phase) return launch_reduce(rank1_linear_reduce_kernel<NB>, p, dim3(grid.x, NB / kBlReduceCols, grid.z), st);
    return BDE_OK;
}

extern "C" int bde_rank1_linear_fwd(const float* x, int64_t ldx, int batch, int in_features, int out_features, const float* W,
                                    const float* s_mu, const float* s_rho, const float* r_mu, const float* r_rho,
                                    const float* bias, const float* eps_s, const float* eps_r, uint64_t seed, uint64_t sid_s,
                                    uint64_t sid_r, float* out, float* lin, float* s_out, float* r_out, float* eps_s_out,
                                    float* eps_r_out, void* workspace, size_t workspace_bytes, bde_stream_t stream) {
    if (!x || !W || !s_mu || !s_rho || !r_mu || !r_rho || !out || !lin || !s_out || !r_out || !eps_s_out || !eps_r_out ||
        !workspace || batch < 1 || in_features < 1 || out_features < 1 || ldx < in_features)
        return BDE_ERR_INVALID_ARG;
    if (!aligned16(x) || !aligned16(W) || (ldx % 4) != 0 || (in_features % 4) != 0) return BDE_ERR_ALIGNMENT;
    size_t need = 0;
    bde_bbb_linear_workspace_bytes(batch, in_features, out_features, &need);   // same tiling, half the partials
    if (workspace_bytes < need) return BDE_ERR_WORKSPACE;
    const int nb = bbl_nb(batch);
    const int mtiles = (out_features + kBlM - 1) / kBlM, btiles = (batch + nb - 1) / nb;
    const int ks = bbl_ksplits(batch, in_features, out_features);
    CUtensorMap mw, mx;
    int rc = encode_sw128_map(&mw, W, out_features, in_features, in_features, kBlM);
    if (rc == BDE_OK) rc = encode_sw128_map(&mx, x, batch, in_features, ldx, nb);
    if (rc != BDE_OK) return rc;
    R1Params p{};
    p.s_mu = s_mu, p.s_rho = s_rho, p.r_mu = r_mu, p.r_rho = r_rho, p.bias = bias, p.eps_s = eps_s, p.eps_r = eps_r;
    p.out = out, p.lin = lin, p.s_out = s_out, p.r_out = r_out, p.eps_s_out = eps_s_out, p.eps_r_out = eps_r_out;
    p.tickets = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + 256);
    p.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256 +
                                          ((static_cast<size_t>(mtiles) * btiles * sizeof(unsigned int) + 255) & ~static_cast<size_t>(255)));
    p.batch = batch, p.in_features = in_features, p.out_features = out_features, p.ksplits = ks;
    p.two_phase = ks > kBlTicketSplits;
    p.seed = seed, p.sid_s = sid_s, p.sid_r = sid_r;
    const dim3 grid(mtiles, ks, btiles);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (nb) {
        case 16: return launch_r1<16>(mw, mx, p, grid, st);
        case 32: return launch_r1<32>(mw, mx, p, grid, st);
        case 64: return launch_r1<64>(mw, mx, p, grid, st);
        default: return launch_r1<128>(mw, mx, p, grid, st);
    }
}
