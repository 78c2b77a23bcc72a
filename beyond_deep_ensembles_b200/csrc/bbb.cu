// bbb.cu — Bayes-by-Backprop / Rank-1 VI: reparameterised Gaussian sample (fwd + bwd),
// prior KL value with analytic gradient, L2 term of the deterministic parameters.
// Reference arithmetic: src/algos/util.py:170-183 (GaussianParameter), src/algos/bbb.py:18-37
// (priors) and :69-80 (KL / L2 collection).
#include "elementwise.cuh"

namespace bde {

// K8 fwd: w = mu + eps * softplus(rho) ---------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kEwThreads)
gauss_sample_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ rho, float* __restrict__ w, int64_t P,
                        const float* __restrict__ eps, uint64_t seed, uint64_t stream_id, int64_t quad0) {
    BDE_QUAD_LOOP(q, P) {
        const int64_t b = q << 2;
        const float4 m = load_quad<VEC, true>(mu, b, P);
        const float4 r = load_quad<VEC, true>(rho, b, P);
        float4 e;
        if (eps)
            e = load_quad<VEC, true>(eps, b, P);
        else
            e = philox_normal4(seed, stream_id, static_cast<uint64_t>(quad0 + q));
        auto f = [&](float mv, float rv, float ev) { return __fadd_rn(mv, __fmul_rn(ev, softplus_ref(rv))); };
        store_quad<VEC>(w, b, P, BDE_LANES(f(m.x, r.x, e.x), f(m.y, r.y, e.y), f(m.z, r.z, e.z), f(m.w, r.w, e.w)));
    }
}

// K8 bwd: grad_rho = (grad_w * eps) * z / (z + 1), z = exp(rho) ----------------------------
template <bool VEC>
__global__ void __launch_bounds__(kEwThreads)
gauss_sample_bwd_kernel(const float* __restrict__ grad_w, const float* __restrict__ rho, float* __restrict__ grad_rho,
                        int64_t P, const float* __restrict__ eps, uint64_t seed, uint64_t stream_id, int64_t quad0) {
    BDE_QUAD_LOOP(q, P) {
        const int64_t b = q << 2;
        const float4 g = load_quad<VEC, true>(grad_w, b, P);
        const float4 r = load_quad<VEC, true>(rho, b, P);
        float4 e;
        if (eps)
            e = load_quad<VEC, true>(eps, b, P);
        else
            e = philox_normal4(seed, stream_id, static_cast<uint64_t>(quad0 + q));
        auto f = [&](float gv, float rv, float ev) {
            const float gs = __fmul_rn(gv, ev);
            if (rv > 20.0f) return gs;
            const float z = expf(rv);
            return __fdiv_rn(__fmul_rn(gs, z), __fadd_rn(z, 1.0f));
        };
        store_quad<VEC>(grad_rho, b, P, BDE_LANES(f(g.x, r.x, e.x), f(g.y, r.y, e.y), f(g.z, r.z, e.z), f(g.w, r.w, e.w)));
    }
}

// shared tail of the value+grad kernels: CTA fp64 partial -> deterministic grid sum -> *value
__device__ __forceinline__ void finish_value(double local, double factor, double* value, void* ws) {
    __shared__ double cta_val;
    __shared__ double total;
    block_sum_fp64(local, &cta_val);
    if (value == nullptr) return;
    if (grid_reduce_fp64(&cta_val, 1, ws, &total)) {
        if (threadIdx.x == 0) *value = total * factor;
    }
}

__device__ __forceinline__ float resolve_scale(float host_scale, const float* dev_scale) {
    return dev_scale ? __fmul_rn(host_scale, __ldg(dev_scale)) : host_scale;
}

// per-element arithmetic shared by the single-tensor and the multi-tensor kernels ------------
// bbb.py:20  0.5*(2*log(sp/s) - 1 + (s/sp)^2 + ((mp - m)/sp)^2), s = softplus(rho); analytic gradient
template <bool GRAD>
__device__ __forceinline__ float kl_gauss_elem(float mv, float rv, float prior_mu, float prior_sigma, float inv_var_p,
                                               float& dmu, float& drho) {
    const float sigma = softplus_ref(rv);
    const float a = __fmul_rn(2.0f, logf(__fdiv_rn(prior_sigma, sigma)));
    const float ratio = __fdiv_rn(sigma, prior_sigma);
    const float dm = __fdiv_rn(__fsub_rn(prior_mu, mv), prior_sigma);
    float kl = __fadd_rn(__fadd_rn(__fsub_rn(a, 1.0f), __fmul_rn(ratio, ratio)), __fmul_rn(dm, dm));
    kl = __fmul_rn(0.5f, kl);
    if (GRAD) {
        dmu = __fmul_rn(__fsub_rn(mv, prior_mu), inv_var_p);
        const float dsig = __fadd_rn(__fdiv_rn(-1.0f, sigma), __fmul_rn(sigma, inv_var_p));
        drho = __fmul_rn(dsig, softplus_grad_ref(rv));
    }
    return kl;
}

struct MixtureConsts {
    float log_pi, log_1mpi;
    float inv_var1, inv_var2;      // 1/sigma^2
    float lognorm1, lognorm2;      // -log(sigma) - 0.5*log(2*pi)
};

// bbb.py:31-34: returns logaddexp(...) (the KL term is its negative) and d(-lse)/dmu
template <bool GRAD>
__device__ __forceinline__ float mixture_lse_elem(float mv, const MixtureConsts& c, float& dkl) {
    // Normal(0, s).log_prob(v) = -v^2/(2 s^2) - log s - log sqrt(2 pi)
    const float l1 = __fadd_rn(__fmul_rn(-0.5f * c.inv_var1, __fmul_rn(mv, mv)), c.lognorm1);
    const float l2 = __fadd_rn(__fmul_rn(-0.5f * c.inv_var2, __fmul_rn(mv, mv)), c.lognorm2);
    const float c1 = fminf(fmaxf(l1, -23.0f), 0.0f);
    const float c2 = fminf(fmaxf(l2, -23.0f), 0.0f);
    const float p1 = __fadd_rn(c.log_pi, c1), p2 = __fadd_rn(c.log_1mpi, c2);
    const float mx = fmaxf(p1, p2), mn = fminf(p1, p2);
    const float lse = __fadd_rn(mx, log1pf(expf(__fsub_rn(mn, mx))));
    if (GRAD) {
        const float w1 = expf(__fsub_rn(p1, lse)), w2 = expf(__fsub_rn(p2, lse));
        const float d1 = (l1 >= -23.0f && l1 <= 0.0f) ? __fmul_rn(-mv, c.inv_var1) : 0.0f;
        const float d2 = (l2 >= -23.0f && l2 <= 0.0f) ? __fmul_rn(-mv, c.inv_var2) : 0.0f;
        dkl = -__fadd_rn(__fmul_rn(w1, d1), __fmul_rn(w2, d2));
    }
    return lse;
}

// K9: Gaussian prior ------------------------------------------------------------------------
template <bool VEC, bool GRAD, bool ACC>
__global__ void __launch_bounds__(kEwThreads)
kl_gauss_kernel(const float* __restrict__ mu, const float* __restrict__ rho, int64_t P, float prior_mu,
                float prior_sigma, double* value, float* __restrict__ grad_mu, float* __restrict__ grad_rho,
                float host_scale, const float* __restrict__ dev_scale, void* ws) {
    const float scale = resolve_scale(host_scale, dev_scale);
    const float inv_var_p = __fdiv_rn(1.0f, __fmul_rn(prior_sigma, prior_sigma));
    double local = 0.0;
    BDE_QUAD_LOOP(q, P) {
        const int64_t b = q << 2;
        const float4 m = load_quad<VEC, true>(mu, b, P);
        const float4 r = load_quad<VEC, true>(rho, b, P);
        float4 gm = make_float4(0.f, 0.f, 0.f, 0.f), gr = gm;
        if (GRAD && ACC) {
            gm = load_quad<VEC, false>(grad_mu, b, P);
            gr = load_quad<VEC, false>(grad_rho, b, P);
        }
        auto f = [&](float mv, float rv, float& gmv, float& grv, bool valid) {
            float dmu = 0.0f, drho = 0.0f;
            const float kl = kl_gauss_elem<GRAD>(mv, rv, prior_mu, prior_sigma, inv_var_p, dmu, drho);
            if (valid) local += static_cast<double>(kl);
            if (GRAD) {
                gmv = ACC ? fmaf(scale, dmu, gmv) : __fmul_rn(scale, dmu);
                grv = ACC ? fmaf(scale, drho, grv) : __fmul_rn(scale, drho);
            }
        };
        f(m.x, r.x, gm.x, gr.x, b + 0 < P);
        f(m.y, r.y, gm.y, gr.y, b + 1 < P);
        f(m.z, r.z, gm.z, gr.z, b + 2 < P);
        f(m.w, r.w, gm.w, gr.w, b + 3 < P);
        if (GRAD) {
            store_quad<VEC>(grad_mu, b, P, gm);
            store_quad<VEC>(grad_rho, b, P, gr);
        }
    }
    finish_value(local, 1.0, value, ws);
}

// K9b: scale-mixture prior --------------------------------------------------------------------

template <bool VEC, bool GRAD, bool ACC>
__global__ void __launch_bounds__(kEwThreads)
kl_mixture_kernel(const float* __restrict__ mu, int64_t P, MixtureConsts c, double* value, float* __restrict__ grad_mu,
                  float host_scale, const float* __restrict__ dev_scale, void* ws) {
    const float scale = resolve_scale(host_scale, dev_scale);
    double local = 0.0;
    BDE_QUAD_LOOP(q, P) {
        const int64_t b = q << 2;
        const float4 m = load_quad<VEC, true>(mu, b, P);
        float4 gm = make_float4(0.f, 0.f, 0.f, 0.f);
        if (GRAD && ACC) gm = load_quad<VEC, false>(grad_mu, b, P);
        auto f = [&](float mv, float& gmv, bool valid) {
            float dkl = 0.0f;
            const float lse = mixture_lse_elem<GRAD>(mv, c, dkl);
            if (valid) local -= static_cast<double>(lse);
            if (GRAD) gmv = ACC ? fmaf(scale, dkl, gmv) : __fmul_rn(scale, dkl);
        };
        f(m.x, gm.x, b + 0 < P);
        f(m.y, gm.y, b + 1 < P);
        f(m.z, gm.z, b + 2 < P);
        f(m.w, gm.w, b + 3 < P);
        if (GRAD) store_quad<VEC>(grad_mu, b, P, gm);
    }
    finish_value(local, 1.0, value, ws);
}

// K10: L2 of deterministic parameters ---------------------------------------------------------
template <bool VEC, bool GRAD, bool ACC>
__global__ void __launch_bounds__(kEwThreads)
l2_kernel(const float* __restrict__ theta, int64_t D, float l2_scale, double value_factor, double* value,
          float* __restrict__ grad, float host_scale, const float* __restrict__ dev_scale, void* ws) {
    const float scale = __fmul_rn(resolve_scale(host_scale, dev_scale), l2_scale);
    double local = 0.0;
    BDE_QUAD_LOOP(q, D) {
        const int64_t b = q << 2;
        const float4 t = load_quad<VEC, true>(theta, b, D);
        local += static_cast<double>(__fmul_rn(t.x, t.x)) + static_cast<double>(__fmul_rn(t.y, t.y)) +
                 static_cast<double>(__fmul_rn(t.z, t.z)) + static_cast<double>(__fmul_rn(t.w, t.w));
        if (GRAD) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ACC) g = load_quad<VEC, false>(grad, b, D);
            g = BDE_LANES(ACC ? fmaf(scale, t.x, g.x) : __fmul_rn(scale, t.x), ACC ? fmaf(scale, t.y, g.y) : __fmul_rn(scale, t.y),
                          ACC ? fmaf(scale, t.z, g.z) : __fmul_rn(scale, t.z), ACC ? fmaf(scale, t.w, g.w) : __fmul_rn(scale, t.w));
            store_quad<VEC>(grad, b, D, g);
        }
    }
    finish_value(local, value_factor, value, ws);
}

// K9/K10 over a list of tensors: bbb.py:69-76 in ONE launch ------------------------------------
// The reference walks every parameter of every group and adds its KL (Gaussian parameters) or its
// l2_scale/2 * ||theta||^2 (deterministic parameters) to one scalar; autograd then runs ~8 backward
// nodes per tensor.  Here the whole list is one virtual vector of column quads: segment t owns the quads
// [qoff[t], qoff[t+1]) (each tensor padded to a multiple of 4 elements), a thread finds its segment by
// binary search and evaluates that segment's term; the value is one deterministic fp64 grid sum.
constexpr int kPtChunk = 56;
constexpr int kPtGauss = 0, kPtMixture = 1, kPtL2 = 2;
struct PriorTable {
    uint64_t a[kPtChunk];        // mu | theta
    uint64_t b[kPtChunk];        // rho | 0
    uint64_t ga[kPtChunk];       // grad of a (0: none)
    uint64_t gb[kPtChunk];       // grad of b (0: none)
    int64_t qoff[kPtChunk + 1];  // first quad of every segment
    int64_t size[kPtChunk];
    float l2[kPtChunk];          // l2_scale of a deterministic tensor
    unsigned char kind[kPtChunk];
    int count;
};

template <bool GRAD, bool ACC>
__global__ void __launch_bounds__(kEwThreads)
prior_terms_kernel(const __grid_constant__ PriorTable tab, float prior_mu, float prior_sigma, MixtureConsts mc, double* value,
                   int accumulate_value, float host_scale, const float* __restrict__ dev_scale, void* ws) {
    const float scale = resolve_scale(host_scale, dev_scale);
    const float inv_var_p = __fdiv_rn(1.0f, __fmul_rn(prior_sigma, prior_sigma));
    double local = 0.0;
    const int64_t nq = tab.qoff[tab.count];
    for (int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < nq;
         q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        int lo = 0, hi = tab.count - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tab.qoff[mid] <= q)
                lo = mid;
            else
                hi = mid - 1;
        }
        const int t = lo;
        const int64_t P = tab.size[t];
        const int64_t b = (q - tab.qoff[t]) << 2;
        const int kind = tab.kind[t];
        const float* pa = reinterpret_cast<const float*>(tab.a[t]);
        const float* pb = reinterpret_cast<const float*>(tab.b[t]);
        float* ga = reinterpret_cast<float*>(tab.ga[t]);
        float* gb = reinterpret_cast<float*>(tab.gb[t]);
        const bool vec = ((tab.a[t] | tab.b[t] | tab.ga[t] | tab.gb[t]) & 15u) == 0;
        const bool want_grad = GRAD && ga != nullptr;
        float4 m, r = make_float4(0.f, 0.f, 0.f, 0.f), gm = r, gr = r;
        m = vec ? load_quad<true, true>(pa, b, P) : load_quad<false, true>(pa, b, P);
        if (kind == kPtGauss) r = vec ? load_quad<true, true>(pb, b, P) : load_quad<false, true>(pb, b, P);
        if (want_grad && ACC) {
            gm = vec ? load_quad<true, false>(ga, b, P) : load_quad<false, false>(ga, b, P);
            if (kind == kPtGauss) gr = vec ? load_quad<true, false>(gb, b, P) : load_quad<false, false>(gb, b, P);
        }
        const float mv[4] = {m.x, m.y, m.z, m.w}, rv[4] = {r.x, r.y, r.z, r.w};
        float gmv[4] = {gm.x, gm.y, gm.z, gm.w}, grv[4] = {gr.x, gr.y, gr.z, gr.w};
        if (kind == kPtGauss) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float dmu = 0.0f, drho = 0.0f;
                const float kl = kl_gauss_elem<GRAD>(mv[e], rv[e], prior_mu, prior_sigma, inv_var_p, dmu, drho);
                if (b + e < P) local += static_cast<double>(kl);
                if (GRAD) {
                    gmv[e] = ACC ? fmaf(scale, dmu, gmv[e]) : __fmul_rn(scale, dmu);
                    grv[e] = ACC ? fmaf(scale, drho, grv[e]) : __fmul_rn(scale, drho);
                }
            }
        } else if (kind == kPtMixture) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float dkl = 0.0f;
                const float lse = mixture_lse_elem<GRAD>(mv[e], mc, dkl);
                if (b + e < P) local -= static_cast<double>(lse);
                if (GRAD) gmv[e] = ACC ? fmaf(scale, dkl, gmv[e]) : __fmul_rn(scale, dkl);
            }
        } else {
            const float l2 = tab.l2[t];
            const float sl = __fmul_rn(scale, l2);
            double sq = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sq += static_cast<double>(__fmul_rn(mv[e], mv[e]));  // padding lanes are 0
                if (GRAD) gmv[e] = ACC ? fmaf(sl, mv[e], gmv[e]) : __fmul_rn(sl, mv[e]);
            }
            local += 0.5 * static_cast<double>(l2) * sq;
        }
        if (want_grad) {
            const float4 o1 = make_float4(gmv[0], gmv[1], gmv[2], gmv[3]);
            if (vec)
                store_quad<true>(ga, b, P, o1);
            else
                store_quad<false>(ga, b, P, o1);
            if (kind == kPtGauss) {
                const float4 o2 = make_float4(grv[0], grv[1], grv[2], grv[3]);
                if (vec)
                    store_quad<true>(gb, b, P, o2);
                else
                    store_quad<false>(gb, b, P, o2);
            }
        }
    }
    __shared__ double cta_val;
    __shared__ double total;
    block_sum_fp64(local, &cta_val);
    if (value == nullptr) return;
    if (grid_reduce_fp64(&cta_val, 1, ws, &total)) {
        if (threadIdx.x == 0) *value = accumulate_value ? *value + total : total;
    }
}

}  // namespace bde

using namespace bde;

static MixtureConsts mixture_consts(double pi, double sigma1, double sigma2);

extern "C" int bde_gauss_sample_fwd(const float* mu, const float* rho, float* w, int64_t P, const float* eps,
                                    uint64_t seed, uint64_t stream_id, int64_t elem0, bde_stream_t stream) {
    if (!mu || !rho || !w || P < 0 || elem0 < 0 || (elem0 & 3)) return BDE_ERR_INVALID_ARG;
    if (P == 0) return BDE_OK;
    const bool vec = aligned16(mu) && aligned16(rho) && aligned16(w) && (!eps || aligned16(eps));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc_;
    if (vec)
        rc_ = launch_ew(gauss_sample_fwd_kernel<true>, P, st, mu, rho, w, P, eps, seed, stream_id, elem0 >> 2);
    else
        rc_ = launch_ew(gauss_sample_fwd_kernel<false>, P, st, mu, rho, w, P, eps, seed, stream_id, elem0 >> 2);
    return rc_;
}

extern "C" int bde_gauss_sample_bwd(const float* grad_w, const float* rho, float* grad_rho, int64_t P,
                                    const float* eps, uint64_t seed, uint64_t stream_id, int64_t elem0,
                                    bde_stream_t stream) {
    if (!grad_w || !rho || !grad_rho || P < 0 || elem0 < 0 || (elem0 & 3)) return BDE_ERR_INVALID_ARG;
    if (P == 0) return BDE_OK;
    const bool vec = aligned16(grad_w) && aligned16(rho) && aligned16(grad_rho) && (!eps || aligned16(eps));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc_;
    if (vec)
        rc_ = launch_ew(gauss_sample_bwd_kernel<true>, P, st, grad_w, rho, grad_rho, P, eps, seed, stream_id, elem0 >> 2);
    else
        rc_ = launch_ew(gauss_sample_bwd_kernel<false>, P, st, grad_w, rho, grad_rho, P, eps, seed, stream_id, elem0 >> 2);
    return rc_;
}

extern "C" int bde_value_workspace_bytes(size_t* bytes) {
    if (!bytes) return BDE_ERR_INVALID_ARG;
    *bytes = grid_reduce_ws_bytes(kMaxCtasEw, 1);
    return BDE_OK;
}

static int check_value_ws(double* value, void* ws, size_t ws_bytes) {
    if (value && (!ws || ws_bytes < grid_reduce_ws_bytes(kMaxCtasEw, 1))) return BDE_ERR_WORKSPACE;
    return BDE_OK;
}

#define BDE_DISPATCH3(KERNEL, nelem, vec, grad, acc, ...)                                          \
    do {                                                                                           \
        if (vec) {                                                                                 \
            if (!grad)                                                                             \
                rc = launch_ew(KERNEL<true, false, false>, nelem, st, __VA_ARGS__);                \
            else if (acc)                                                                          \
                rc = launch_ew(KERNEL<true, true, true>, nelem, st, __VA_ARGS__);                  \
            else                                                                                   \
                rc = launch_ew(KERNEL<true, true, false>, nelem, st, __VA_ARGS__);                 \
        } else {                                                                                   \
            if (!grad)                                                                             \
                rc = launch_ew(KERNEL<false, false, false>, nelem, st, __VA_ARGS__);               \
            else if (acc)                                                                          \
                rc = launch_ew(KERNEL<false, true, true>, nelem, st, __VA_ARGS__);                 \
            else                                                                                   \
                rc = launch_ew(KERNEL<false, true, false>, nelem, st, __VA_ARGS__);                \
        }                                                                                          \
    } while (0)

extern "C" int bde_kl_gauss_value_and_grad(const float* mu, const float* rho, int64_t P, double prior_mu,
                                           double prior_sigma, double* value, float* grad_mu, float* grad_rho,
                                           double grad_scale, const float* grad_scale_dev, int accumulate,
                                           void* workspace, size_t workspace_bytes, bde_stream_t stream) {
    if (!mu || !rho || P < 0 || !(prior_sigma > 0.0) || ((grad_mu == nullptr) != (grad_rho == nullptr)))
        return BDE_ERR_INVALID_ARG;
    int rc = check_value_ws(value, workspace, workspace_bytes);
    if (rc != BDE_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P == 0) {
        if (value) BDE_RETURN_IF_CUDA(cudaMemsetAsync(value, 0, sizeof(double), st));
        return BDE_OK;
    }
    const bool grad = grad_mu != nullptr;
    const bool vec = aligned16(mu) && aligned16(rho) && (!grad || (aligned16(grad_mu) && aligned16(grad_rho)));
    BDE_DISPATCH3(kl_gauss_kernel, P, vec, grad, accumulate != 0, mu, rho, P, static_cast<float>(prior_mu),
                  static_cast<float>(prior_sigma), value, grad_mu, grad_rho, static_cast<float>(grad_scale),
                  grad_scale_dev, workspace);
    return rc;
}

extern "C" int bde_kl_mixture_value_and_grad(const float* mu, int64_t P, double pi, double sigma1, double sigma2,
                                             double* value, float* grad_mu, double grad_scale,
                                             const float* grad_scale_dev, int accumulate, void* workspace,
                                             size_t workspace_bytes, bde_stream_t stream) {
    if (!mu || P < 0 || !(pi > 0.0 && pi < 1.0) || !(sigma1 > 0.0) || !(sigma2 > 0.0)) return BDE_ERR_INVALID_ARG;
    int rc = check_value_ws(value, workspace, workspace_bytes);
    if (rc != BDE_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P == 0) {
        if (value) BDE_RETURN_IF_CUDA(cudaMemsetAsync(value, 0, sizeof(double), st));
        return BDE_OK;
    }
    const MixtureConsts c = mixture_consts(pi, sigma1, sigma2);
    const bool grad = grad_mu != nullptr;
    const bool vec = aligned16(mu) && (!grad || aligned16(grad_mu));
    BDE_DISPATCH3(kl_mixture_kernel, P, vec, grad, accumulate != 0, mu, P, c, value, grad_mu,
                  static_cast<float>(grad_scale), grad_scale_dev, workspace);
    return rc;
}

extern "C" int bde_l2_value_and_grad(const float* theta, int64_t D, double l2_scale, double* value, float* grad,
                                     double grad_scale, const float* grad_scale_dev, int accumulate, void* workspace,
                                     size_t workspace_bytes, bde_stream_t stream) {
    if (!theta || D < 0) return BDE_ERR_INVALID_ARG;
    int rc = check_value_ws(value, workspace, workspace_bytes);
    if (rc != BDE_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (D == 0) {
        if (value) BDE_RETURN_IF_CUDA(cudaMemsetAsync(value, 0, sizeof(double), st));
        return BDE_OK;
    }
    const bool has_grad = grad != nullptr;
    const bool vec = aligned16(theta) && (!has_grad || aligned16(grad));
    BDE_DISPATCH3(l2_kernel, D, vec, has_grad, accumulate != 0, theta, D, static_cast<float>(l2_scale), 0.5 * l2_scale, value,
                  grad, static_cast<float>(grad_scale), grad_scale_dev, workspace);
    return rc;
}

static MixtureConsts mixture_consts(double pi, double sigma1, double sigma2) {
    MixtureConsts c;
    // torch.log(torch.tensor(pi)) is an fp32 log of the fp32-rounded pi
    c.log_pi = logf(static_cast<float>(pi));
    c.log_1mpi = logf(1.0f - static_cast<float>(pi));
    c.inv_var1 = static_cast<float>(1.0 / (sigma1 * sigma1));
    c.inv_var2 = static_cast<float>(1.0 / (sigma2 * sigma2));
    const double half_log_2pi = 0.91893853320467274178;
    c.lognorm1 = static_cast<float>(-log(sigma1) - half_log_2pi);
    c.lognorm2 = static_cast<float>(-log(sigma2) - half_log_2pi);
    return c;
}

extern "C" int bde_prior_terms_value_and_grad(int count, const int32_t* kinds_host, const uint64_t* a_host,
                                              const uint64_t* b_host, const uint64_t* grad_a_host,
                                              const uint64_t* grad_b_host, const int64_t* sizes_host,
                                              const double* l2_scales_host, double prior_p0, double prior_p1,
                                              double prior_p2, double* value, double grad_scale,
                                              const float* grad_scale_dev, int accumulate_grad, void* workspace,
                                              size_t workspace_bytes, bde_stream_t stream) {
    if (count < 0 || (count > 0 && (!kinds_host || !a_host || !sizes_host))) return BDE_ERR_INVALID_ARG;
    int rc = check_value_ws(value, workspace, workspace_bytes);
    if (rc != BDE_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool grad = grad_a_host != nullptr;
    bool any_gauss = false, any_mix = false;
    for (int i = 0; i < count; ++i) {
        const int k = kinds_host[i];
        if (k < kPtGauss || k > kPtL2 || sizes_host[i] < 0 || (sizes_host[i] > 0 && !a_host[i])) return BDE_ERR_INVALID_ARG;
        if (k == kPtGauss && sizes_host[i] > 0 && (!b_host || !b_host[i])) return BDE_ERR_INVALID_ARG;
        if (k == kPtGauss && grad && grad_a_host[i] && (!grad_b_host || !grad_b_host[i])) return BDE_ERR_INVALID_ARG;
        if (k == kPtL2 && !l2_scales_host) return BDE_ERR_INVALID_ARG;
        any_gauss |= k == kPtGauss;
        any_mix |= k == kPtMixture;
    }
    if (any_gauss && any_mix) return BDE_ERR_INVALID_ARG;  // one prior per call
    if (any_gauss && !(prior_p1 > 0.0)) return BDE_ERR_INVALID_ARG;
    if (any_mix && (!(prior_p0 > 0.0 && prior_p0 < 1.0) || !(prior_p1 > 0.0) || !(prior_p2 > 0.0))) return BDE_ERR_INVALID_ARG;
    if (value) BDE_RETURN_IF_CUDA(cudaMemsetAsync(value, 0, sizeof(double), st));
    const MixtureConsts mc = any_mix ? mixture_consts(prior_p0, prior_p1, prior_p2) : MixtureConsts{};
    const float pmu = any_gauss ? static_cast<float>(prior_p0) : 0.0f;
    const float psig = any_gauss ? static_cast<float>(prior_p1) : 1.0f;
    for (int c0 = 0; c0 < count; c0 += kPtChunk) {
        PriorTable tab;
        tab.count = 0;
        int64_t q = 0;
        for (int i = c0; i < count && tab.count < kPtChunk; ++i) {
            if (sizes_host[i] == 0) continue;
            const int t = tab.count++;
            tab.a[t] = a_host[i];
            tab.b[t] = (kinds_host[i] == kPtGauss) ? b_host[i] : 0;
            tab.ga[t] = grad ? grad_a_host[i] : 0;
            tab.gb[t] = (grad && kinds_host[i] == kPtGauss && grad_a_host[i]) ? grad_b_host[i] : 0;
            tab.size[t] = sizes_host[i];
            tab.l2[t] = (kinds_host[i] == kPtL2) ? static_cast<float>(l2_scales_host[i]) : 0.0f;
            tab.kind[t] = static_cast<unsigned char>(kinds_host[i]);
            tab.qoff[t] = q;
            q += (sizes_host[i] + 3) >> 2;
        }
        if (tab.count == 0) continue;
        tab.qoff[tab.count] = q;
        const float hs = static_cast<float>(grad_scale);
        if (!grad)
            rc = launch_ew(prior_terms_kernel<false, false>, q << 2, st, tab, pmu, psig, mc, value, 1, hs, grad_scale_dev, workspace);
        else if (accumulate_grad)
            rc = launch_ew(prior_terms_kernel<true, true>, q << 2, st, tab, pmu, psig, mc, value, 1, hs, grad_scale_dev, workspace);
        else
            rc = launch_ew(prior_terms_kernel<true, false>, q << 2, st, tab, pmu, psig, mc, value, 1, hs, grad_scale_dev, workspace);
        if (rc != BDE_OK) return rc;
    }
    return BDE_OK;
}
