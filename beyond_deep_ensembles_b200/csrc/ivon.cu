// ivon.cu — iVON posterior sampling, gradient accumulation and Hessian-EMA update.
// Reference arithmetic: src/algos/ivorn.py:66-89 (update), :102-115 (sample), :120-127.
// Each fp32 operation of the reference is reproduced in the same order with explicit
// round-to-nearest intrinsics (no FMA contraction), python-side scalars are folded in
// double and rounded to fp32 exactly where eager PyTorch rounds them.
#include "ew_tma.cuh"

namespace bde {

// ivorn.py:108  delta = 1 / (N * precision.clamp(min=1e-4)).sqrt() * normal
// injected noise (parity runs): the reference's op order with IEEE sqrt / divide;
// in-kernel Philox noise (production): one MUFU.RSQ (rel. error < 2.4e-7, inside rtol 1e-5).
// Both kernel forms share this function, so a result never depends on which form ran.
__device__ __forceinline__ float ivon_delta(bool injected, float n_eff, float pv, float ev) {
    const float x = __fmul_rn(n_eff, fmaxf(pv, 1e-4f));
    if (injected) return __fmul_rn(__fdiv_rn(1.0f, __fsqrt_rn(x)), ev);
    return __fmul_rn(rsqrt_approx(x), ev);
}

// K5 ------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kEwThreads)
ivon_sample_kernel(const float* __restrict__ mean, const float* __restrict__ prec, float* __restrict__ delta_sum,
                   float* __restrict__ theta, int64_t D, float n_eff, int first, int deterministic,
                   const float* __restrict__ eps, uint64_t seed, uint64_t stream_id, int64_t quad0) {
    BDE_QUAD_LOOP(q, D) {
        const int64_t b = q << 2;
        const float4 m = load_quad<VEC, true>(mean, b, D);
        float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!deterministic) {
            const float4 p = load_quad<VEC, true>(prec, b, D);
            float4 e;
            if (eps)
                e = load_quad<VEC, true>(eps, b, D);
            else
                e = philox_normal4(seed, stream_id, static_cast<uint64_t>(quad0 + q));
            const bool inj = eps != nullptr;
            auto f = [&](float pv, float ev) { return ivon_delta(inj, n_eff, pv, ev); };
            dl = BDE_LANES(f(p.x, e.x), f(p.y, e.y), f(p.z, e.z), f(p.w, e.w));
        }
        const float4 th = BDE_LANES(__fadd_rn(m.x, dl.x), __fadd_rn(m.y, dl.y), __fadd_rn(m.z, dl.z), __fadd_rn(m.w, dl.w));
        store_quad<VEC>(theta, b, D, th);
        float4 ds = dl;
        if (!first) {
            const float4 o = load_quad<VEC, false>(delta_sum, b, D);
            ds = BDE_LANES(__fadd_rn(o.x, dl.x), __fadd_rn(o.y, dl.y), __fadd_rn(o.z, dl.z), __fadd_rn(o.w, dl.w));
        }
        store_quad<VEC>(delta_sum, b, D, ds);
    }
}

// K5 batched: S consecutive draws in one pass — mean and precision are read once, delta_sum is
// read / written once, S weight vectors are written: (8 or 16 + 4 S) D bytes instead of 20 S D.  Draw s equals
// bde_ivon_sample with stream_id + s * stream_stride bit for bit (same 1/sqrt factor, same running delta_sum order).
template <bool VEC>
__global__ void __launch_bounds__(kEwThreads)
ivon_sample_batch_kernel(const float* __restrict__ mean, const float* __restrict__ prec, float* __restrict__ delta_sum,
                         float* __restrict__ theta, int64_t ld_out, int64_t D, int S, float n_eff, int first,
                         int deterministic, const float* __restrict__ eps, int64_t ld_eps, uint64_t seed,
                         uint64_t stream_id, uint64_t stream_stride, int64_t quad0) {
    const PhiloxKeys pk = philox_round_keys(seed);
    BDE_QUAD_LOOP(q, D) {
        const int64_t b = q << 2;
        const float4 m = load_quad<VEC, true>(mean, b, D);
        const bool inj = eps != nullptr;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);   // 1 / sqrt(N max(prec, 1e-4)), shared by all draws
        if (!deterministic) {
            const float4 p = load_quad<VEC, true>(prec, b, D);
            auto f = [&](float pv) {
                const float x = __fmul_rn(n_eff, fmaxf(pv, 1e-4f));
                return inj ? __fdiv_rn(1.0f, __fsqrt_rn(x)) : rsqrt_approx(x);
            };
            c = BDE_LANES(f(p.x), f(p.y), f(p.z), f(p.w));
        }
        float4 ds = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!first) ds = load_quad<VEC, false>(delta_sum, b, D);
        for (int sidx = 0; sidx < S; ++sidx) {
            float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!deterministic) {
                float4 e;
                if (inj)
                    e = load_quad<VEC, true>(eps + sidx * ld_eps, b, D);
                else
                    e = philox_normal4(pk, stream_id + sidx * stream_stride, static_cast<uint64_t>(quad0 + q));
                dl = BDE_LANES(__fmul_rn(c.x, e.x), __fmul_rn(c.y, e.y), __fmul_rn(c.z, e.z), __fmul_rn(c.w, e.w));
            }
            const float4 th = BDE_LANES(__fadd_rn(m.x, dl.x), __fadd_rn(m.y, dl.y), __fadd_rn(m.z, dl.z), __fadd_rn(m.w, dl.w));
            store_quad<VEC>(theta + sidx * ld_out, b, D, th);
            if (first && sidx == 0)
                ds = dl;
            else
                ds = BDE_LANES(__fadd_rn(ds.x, dl.x), __fadd_rn(ds.y, dl.y), __fadd_rn(ds.z, dl.z), __fadd_rn(ds.w, dl.w));
        }
        store_quad<VEC>(delta_sum, b, D, ds);
    }
}

// Fast path of K5 batched (production form: Philox noise, 16-byte-aligned pointers, whole quads, up to SB draws per pass —
// the launcher splits S into passes of at most 16 and sends a ragged tail, injected noise and the deterministic mode to
// the general kernel above).  Same arithmetic bit for bit; what is gone is the rolled draw loop with its per-draw
// predicates, guarded accesses and 64-bit address arithmetic: the Philox chains of a quad's draws are independent and
// fully unrolled (instruction-level parallelism instead of one serial 10-round chain at a time), the round keys sit in
// uniform registers, mean / delta_sum move as packed pairs (FADD2), and the next quad's lines are requested into L2 while
// this quad's normals are computed.  `count` <= SB draws are produced (uniform early exit of the unrolled loop).
template <int SB, bool FIRST>
__global__ void __launch_bounds__(kEwThreads)
ivon_sample_batch_fast_kernel(const float* __restrict__ mean, const float* __restrict__ prec, float* __restrict__ delta_sum,
                              float* __restrict__ theta, int64_t ld_out, int64_t nquads, int count, float n_eff, uint64_t seed,
                              uint64_t stream_id, uint64_t stream_stride, int64_t quad0, int pf_dist) {
    const PhiloxKeys pk = philox_round_keys(seed);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < nquads; q += stride) {
        const int64_t b = q << 2;
        const int64_t qn = q + pf_dist * stride;
        if (pf_dist && qn < nquads) {
            const int64_t bn = qn << 2;
            prefetch_l2(mean + bn);
            prefetch_l2(prec + bn);
            if (!FIRST) prefetch_l2(delta_sum + bn);
        }
        const V4 m = ldg_stream_v4(mean + b);
        const float4 p = ldg_stream_f4(prec + b);
        auto f = [&](float pv) { return rsqrt_approx(__fmul_rn(n_eff, fmaxf(pv, 1e-4f))); };   // ivon_delta's factor
        const float c0 = f(p.x), c1 = f(p.y), c2 = f(p.z), c3 = f(p.w);
        V4 ds;
        if constexpr (FIRST) {
            ds.lo = ds.hi = 0ull;
        } else {
            ds = ldg_stream_v4(delta_sum + b);
        }
        float* out = theta + b;
        const uint64_t quad = static_cast<uint64_t>(quad0 + q);
#pragma unroll
        for (int sidx = 0; sidx < SB; ++sidx) {
            if (sidx >= SB / 2 && sidx >= count) break;   // passes are sized so that count > SB / 2 (or SB == 2)
            const float4 z = philox_normal4(pk, stream_id + sidx * stream_stride, quad);
            V4 dl, th;
            dl.lo = pack2(__fmul_rn(c0, z.x), __fmul_rn(c1, z.y));
            dl.hi = pack2(__fmul_rn(c2, z.z), __fmul_rn(c3, z.w));
            th.lo = add2_rn(m.lo, dl.lo);
            th.hi = add2_rn(m.hi, dl.hi);
            stg_stream_v4(out, th);
            out += ld_out;
            if (FIRST && sidx == 0) {
                ds = dl;
            } else {
                ds.lo = add2_rn(ds.lo, dl.lo);
                ds.hi = add2_rn(ds.hi, dl.hi);
            }
        }
        stg_stream_v4(delta_sum + b, ds);
    }
}

// K5, TMA-staged (ew_tma.cuh): inputs mean, prec, [delta_sum unless FIRST], [eps if EPS]
template <bool FIRST, bool EPS>
struct IvonSampleOp {
    static constexpr int NIN = 2 + (FIRST ? 0 : 1) + (EPS ? 1 : 0), NOUT = 2;  // out: theta, delta_sum
    float n_eff;
    uint64_t seed, stream_id;
    int64_t quad0;
    __device__ __forceinline__ void operator()(const float4 (&in)[NIN], float4 (&out)[NOUT], int64_t quad) const {
        const float4 m = in[0], p = in[1];
        float4 e;
        if constexpr (EPS)
            e = in[NIN - 1];
        else
            e = philox_normal4(seed, stream_id, static_cast<uint64_t>(quad0 + quad));
        auto f = [&](float pv, float ev) { return ivon_delta(EPS, n_eff, pv, ev); };
        const float4 dl = BDE_LANES(f(p.x, e.x), f(p.y, e.y), f(p.z, e.z), f(p.w, e.w));
        out[0] = BDE_LANES(__fadd_rn(m.x, dl.x), __fadd_rn(m.y, dl.y), __fadd_rn(m.z, dl.z), __fadd_rn(m.w, dl.w));
        if constexpr (FIRST) {
            out[1] = dl;
        } else {
            const float4 o = in[2];
            out[1] = BDE_LANES(__fadd_rn(o.x, dl.x), __fadd_rn(o.y, dl.y), __fadd_rn(o.z, dl.z), __fadd_rn(o.w, dl.w));
        }
    }
};

template <bool FIRST, bool EPS>
static int launch_ivon_sample_tma(const float* mean, const float* prec, float* delta_sum, float* theta, int64_t D,
                                  float n_eff, const float* eps, uint64_t seed, uint64_t stream_id, int64_t quad0,
                                  cudaStream_t st) {
    using Op = IvonSampleOp<FIRST, EPS>;
    EwPtrs<Op::NIN, Op::NOUT> p;
    int k = 0;
    p.in[k++] = mean;
    p.in[k++] = prec;
    if (!FIRST) p.in[k++] = delta_sum;
    if (EPS) p.in[k++] = eps;
    p.out[0] = theta;
    p.out[1] = delta_sum;
    return launch_ew_tma<Op>(p, D, Op{n_eff, seed, stream_id, quad0}, nullptr, st);
}

// K6 ------------------------------------------------------------------------------------
struct IvonAccumulateOp {
    static constexpr int NIN = 2, NOUT = 1;
    __device__ __forceinline__ void operator()(const float4 (&in)[NIN], float4 (&out)[NOUT], int64_t) const {
        const float4 a = in[0], g = in[1];
        out[0] = BDE_LANES(__fadd_rn(a.x, g.x), __fadd_rn(a.y, g.y), __fadd_rn(a.z, g.z), __fadd_rn(a.w, g.w));
    }
};

template <bool VEC>
__global__ void __launch_bounds__(kEwThreads)
ivon_accumulate_kernel(float* __restrict__ acc, const float* __restrict__ grad, int64_t D, int first) {
    BDE_QUAD_LOOP(q, D) {
        const int64_t b = q << 2;
        float4 g = load_quad<VEC, true>(grad, b, D);
        if (!first) {
            const float4 a = load_quad<VEC, false>(acc, b, D);
            g = BDE_LANES(__fadd_rn(a.x, g.x), __fadd_rn(a.y, g.y), __fadd_rn(a.z, g.z), __fadd_rn(a.w, g.w));
        }
        store_quad<VEC>(acc, b, D, g);
    }
}

// K7 ------------------------------------------------------------------------------------
struct IvonScalars {
    float S;        // mc_samples
    float delta0;   // tempering * prior_prec / N_eff
    float beta1;
    float omb1;     // 1 - beta1
    float n_eff;
    float damping;
    float bc1;      // 1 - beta1**t
    float bc2;      // 1 - beta2**t
    float lr;
    float omb2;     // 1 - beta2
    float c2;       // 0.5 * (1 - beta2)**2
};

__device__ __forceinline__ void ivon_update_one(const IvonScalars& c, float acc, float dsum, float& mean, float& mom,
                                                float& prec) {
    const float gradient = __fdiv_rn(acc, c.S);                                                // :79
    const float g_mu = __fadd_rn(__fmul_rn(c.delta0, mean), gradient);                          // :80
    const float m_new = __fadd_rn(__fmul_rn(c.beta1, mom), __fmul_rn(c.omb1, g_mu));            // :81
    float t = __fmul_rn(c.n_eff, prec);                                                         // :82
    t = __fmul_rn(t, dsum);
    t = __fdiv_rn(t, c.S);
    t = __fmul_rn(t, gradient);
    float g_s = __fadd_rn(__fsub_rn(c.delta0, prec), t);
    g_s = __fadd_rn(g_s, c.damping);
    const float cm = __fdiv_rn(m_new, c.bc1);                                                   // :84
    const float cp = __fdiv_rn(prec, c.bc2);                                                    // :85 (old precision)
    const float mean_new = __fsub_rn(mean, __fdiv_rn(__fmul_rn(c.lr, cm), cp));                 // :88
    float u = __fmul_rn(c.c2, g_s);                                                             // :89
    u = __fdiv_rn(u, prec);
    u = __fadd_rn(c.omb2, u);
    u = __fmul_rn(u, g_s);
    const float prec_new = __fadd_rn(prec, u);
    mean = mean_new;
    mom = m_new;
    prec = prec_new;
}

template <bool VEC>
__global__ void __launch_bounds__(kEwThreads)
ivon_update_kernel(const float* __restrict__ acc_grad, const float* __restrict__ delta_sum, float* __restrict__ mean,
                   float* __restrict__ momentum, float* __restrict__ prec, int64_t D, IvonScalars c) {
    BDE_QUAD_LOOP(q, D) {
        const int64_t b = q << 2;
        const float4 a = load_quad<VEC, true>(acc_grad, b, D);
        const float4 ds = load_quad<VEC, true>(delta_sum, b, D);
        float4 m = load_quad<VEC, false>(mean, b, D);
        float4 mo = load_quad<VEC, false>(momentum, b, D);
        float4 p = load_quad<VEC, false>(prec, b, D);
        ivon_update_one(c, a.x, ds.x, m.x, mo.x, p.x);
        ivon_update_one(c, a.y, ds.y, m.y, mo.y, p.y);
        ivon_update_one(c, a.z, ds.z, m.z, mo.z, p.z);
        ivon_update_one(c, a.w, ds.w, m.w, mo.w, p.w);
        store_quad<VEC>(mean, b, D, m);
        store_quad<VEC>(momentum, b, D, mo);
        store_quad<VEC>(prec, b, D, p);
    }
}

// K7, TMA-staged: inputs acc_grad, delta_sum, mean, momentum, prec; outputs mean, momentum, prec
struct IvonUpdateOp {
    static constexpr int NIN = 5, NOUT = 3;
    IvonScalars c;
    __device__ __forceinline__ void operator()(const float4 (&in)[NIN], float4 (&out)[NOUT], int64_t) const {
        const float4 a = in[0], ds = in[1];
        float4 m = in[2], mo = in[3], p = in[4];
        ivon_update_one(c, a.x, ds.x, m.x, mo.x, p.x);
        ivon_update_one(c, a.y, ds.y, m.y, mo.y, p.y);
        ivon_update_one(c, a.z, ds.z, m.z, mo.z, p.z);
        ivon_update_one(c, a.w, ds.w, m.w, mo.w, p.w);
        out[0] = m;
        out[1] = mo;
        out[2] = p;
    }
};

}  // namespace bde

using namespace bde;

extern "C" int bde_ivon_sample(const float* mean, const float* prec, float* delta_sum, float* theta, int64_t D,
                               double n_eff, int first, int deterministic, const float* eps, uint64_t seed,
                               uint64_t stream_id, int64_t elem0, bde_stream_t stream) {
    if (!mean || !prec || !delta_sum || !theta || D < 0 || elem0 < 0 || (elem0 & 3)) return BDE_ERR_INVALID_ARG;
    if (D == 0) return BDE_OK;
    const bool vec = aligned16(mean) && aligned16(prec) && aligned16(delta_sum) && aligned16(theta) &&
                     (!eps || aligned16(eps));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float nf = static_cast<float>(n_eff);
    if (!deterministic && use_ew_tma(D, vec, false)) {
        const int64_t q0 = elem0 >> 2;
        if (first)
            return eps ? launch_ivon_sample_tma<true, true>(mean, prec, delta_sum, theta, D, nf, eps, seed, stream_id, q0, st)
                       : launch_ivon_sample_tma<true, false>(mean, prec, delta_sum, theta, D, nf, eps, seed, stream_id, q0, st);
        return eps ? launch_ivon_sample_tma<false, true>(mean, prec, delta_sum, theta, D, nf, eps, seed, stream_id, q0, st)
                   : launch_ivon_sample_tma<false, false>(mean, prec, delta_sum, theta, D, nf, eps, seed, stream_id, q0, st);
    }
    int rc_;
    if (vec)
        rc_ = launch_ew(ivon_sample_kernel<true>, D, st, mean, prec, delta_sum, theta, D, nf, first,
                                                                 deterministic, eps, seed, stream_id, elem0 >> 2);
    else
        rc_ = launch_ew(ivon_sample_kernel<false>, D, st, mean, prec, delta_sum, theta, D, nf, first,
                                                                  deterministic, eps, seed, stream_id, elem0 >> 2);
    return rc_;
}

extern "C" int bde_ivon_sample_batch(const float* mean, const float* prec, float* delta_sum, float* theta,
                                     int64_t ld_out, int64_t D, int S, double n_eff, int first, int deterministic,
                                     const float* eps, int64_t ld_eps, uint64_t seed, uint64_t stream_id,
                                     uint64_t stream_stride, int64_t elem0, bde_stream_t stream) {
    if (!mean || !prec || !delta_sum || !theta || D < 0 || S < 0 || ld_out < D || (eps && ld_eps < D) || elem0 < 0 ||
        (elem0 & 3))
        return BDE_ERR_INVALID_ARG;
    if (D == 0 || S == 0) return BDE_OK;
    const bool vec = aligned16(mean) && aligned16(prec) && aligned16(delta_sum) && aligned16(theta) && (ld_out % 4 == 0) &&
                     (!eps || (aligned16(eps) && ld_eps % 4 == 0));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float nf = static_cast<float>(n_eff);
    const int64_t nq = D >> 2;
    if (vec && !eps && !deterministic && S >= 2 && nq > 0 && tuning().swag_batch != 1) {
        // production form: passes of up to 16 draws through the fast kernel (delta_sum is carried from pass to pass in
        // draw order, so the running sum is the one S single calls produce) ...
        const int pf = tuning().batch_prefetch == 0 ? 1 : (tuning().batch_prefetch >= 9 ? 0 : tuning().batch_prefetch);
        for (int s0 = 0; s0 < S;) {
            const int c = S - s0 < 16 ? S - s0 : 16;
            const int sb = c <= 2 ? 2 : (c <= 4 ? 4 : (c <= 8 ? 8 : 16));   // c > sb / 2 unless sb == 2 (c = 1: a lone last draw)
            const bool f0 = first && s0 == 0;
            float* out = theta + s0 * ld_out;
            const uint64_t sid = stream_id + static_cast<uint64_t>(s0) * stream_stride;
            int rf = BDE_OK;
#define BDE_IVON_FAST(SB_)                                                                                                          \
    rf = f0 ? launch_ew(ivon_sample_batch_fast_kernel<SB_, true>, nq * 4, st, mean, prec, delta_sum, out, ld_out, nq, c, nf, seed, sid, \
                        stream_stride, elem0 >> 2, pf)                                                                              \
            : launch_ew(ivon_sample_batch_fast_kernel<SB_, false>, nq * 4, st, mean, prec, delta_sum, out, ld_out, nq, c, nf, seed, sid, \
                        stream_stride, elem0 >> 2, pf)
            switch (sb) {
                case 2: BDE_IVON_FAST(2); break;
                case 4: BDE_IVON_FAST(4); break;
                case 8: BDE_IVON_FAST(8); break;
                default: BDE_IVON_FAST(16); break;
            }
#undef BDE_IVON_FAST
            if (rf != BDE_OK) return rf;
            s0 += c;
        }
        const int64_t d4 = nq << 2;
        if (d4 < D)     // ... and the general kernel for all S draws of the last D % 4 elements
            return launch_ew(ivon_sample_batch_kernel<false>, D - d4, st, mean + d4, prec + d4, delta_sum + d4, theta + d4, ld_out,
                             D - d4, S, nf, first, deterministic, eps, ld_eps, seed, stream_id, stream_stride, (elem0 + d4) >> 2);
        return BDE_OK;
    }
    if (vec)
        return launch_ew(ivon_sample_batch_kernel<true>, D, st, mean, prec, delta_sum, theta, ld_out, D, S, nf, first,
                         deterministic, eps, ld_eps, seed, stream_id, stream_stride, elem0 >> 2);
    return launch_ew(ivon_sample_batch_kernel<false>, D, st, mean, prec, delta_sum, theta, ld_out, D, S, nf, first,
                     deterministic, eps, ld_eps, seed, stream_id, stream_stride, elem0 >> 2);
}

extern "C" int bde_ivon_accumulate(float* acc, const float* grad, int64_t D, int first, bde_stream_t stream) {
    if (!acc || !grad || D < 0) return BDE_ERR_INVALID_ARG;
    if (D == 0) return BDE_OK;
    const bool vec = aligned16(acc) && aligned16(grad);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!first && use_ew_tma(D, vec, false)) {
        EwPtrs<2, 1> p;
        p.in[0] = acc;
        p.in[1] = grad;
        p.out[0] = acc;
        return launch_ew_tma<IvonAccumulateOp>(p, D, IvonAccumulateOp{}, nullptr, st);
    }
    int rc_;
    if (vec)
        rc_ = launch_ew(ivon_accumulate_kernel<true>, D, st, acc, grad, D, first);
    else
        rc_ = launch_ew(ivon_accumulate_kernel<false>, D, st, acc, grad, D, first);
    return rc_;
}

extern "C" int bde_ivon_update(const float* acc_grad, const float* delta_sum, float* mean, float* momentum,
                               float* prec, int64_t D, int mc_samples, int64_t step, double lr, double beta1,
                               double beta2, double prior_prec, double n_eff, double tempering, double damping,
                               bde_stream_t stream) {
    if (!acc_grad || !delta_sum || !mean || !momentum || !prec || D < 0 || mc_samples < 1 || step < 1)
        return BDE_ERR_INVALID_ARG;
    if (D == 0) return BDE_OK;
    IvonScalars c;
    c.S = static_cast<float>(mc_samples);
    c.delta0 = static_cast<float>(tempering * prior_prec / n_eff);
    c.beta1 = static_cast<float>(beta1);
    c.omb1 = static_cast<float>(1.0 - beta1);
    c.n_eff = static_cast<float>(n_eff);
    c.damping = static_cast<float>(damping);
    c.bc1 = static_cast<float>(1.0 - pow(beta1, static_cast<double>(step)));
    c.bc2 = static_cast<float>(1.0 - pow(beta2, static_cast<double>(step)));
    c.lr = static_cast<float>(lr);
    c.omb2 = static_cast<float>(1.0 - beta2);
    c.c2 = static_cast<float>(0.5 * (1.0 - beta2) * (1.0 - beta2));
    const bool vec = aligned16(acc_grad) && aligned16(delta_sum) && aligned16(mean) && aligned16(momentum) && aligned16(prec);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (use_ew_tma(D, vec, true)) {
        EwPtrs<5, 3> p;
        p.in[0] = acc_grad;
        p.in[1] = delta_sum;
        p.in[2] = mean;
        p.in[3] = momentum;
        p.in[4] = prec;
        p.out[0] = mean;
        p.out[1] = momentum;
        p.out[2] = prec;
        return launch_ew_tma<IvonUpdateOp>(p, D, IvonUpdateOp{c}, nullptr, st);
    }
    int rc_;
    if (vec)
        rc_ = launch_ew(ivon_update_kernel<true>, D, st, acc_grad, delta_sum, mean, momentum, prec, D, c);
    else
        rc_ = launch_ew(ivon_update_kernel<false>, D, st, acc_grad, delta_sum, mean, momentum, prec, D, c);
    return rc_;
}
