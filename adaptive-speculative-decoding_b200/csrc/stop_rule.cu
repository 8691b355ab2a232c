// Cascade stop rule: backward-induction DP of the reference's optimal_stopping_rule
// (/root/reference/src/algorithms/dp_solver.py:12-71) and the Beta-posterior shrinkage
// bayesian_adjustment (:106-130), evaluated for a batch of requests on the device so the
// stop decision needs no logits/feature round trip, plus the same arithmetic as a host
// function for the scalar Python API (policy seam 3, pipeline.py:236-238,251-256).
//
// Bit-exactness: the reference runs in CPython binary64.  Every operation here is an explicit
// round-to-nearest binary64 intrinsic in the same order with the same `<=` tie rule, so no
// FMA contraction can change a result.
#include <cuda_runtime.h>

#include "asd_internal.h"

namespace asd {

constexpr int kMaxStages = 64;

template <bool kDevice>
struct F64 {
    static __host__ __device__ __forceinline__ double mul(double a, double b) {
#ifdef __CUDA_ARCH__
        return __dmul_rn(a, b);
#else
        volatile double r = a * b;
        return r;
#endif
    }
    static __host__ __device__ __forceinline__ double add(double a, double b) {
#ifdef __CUDA_ARCH__
        return __dadd_rn(a, b);
#else
        volatile double r = a + b;
        return r;
#endif
    }
    static __host__ __device__ __forceinline__ double div(double a, double b) {
#ifdef __CUDA_ARCH__
        return __ddiv_rn(a, b);
#else
        volatile double r = a / b;
        return r;
#endif
    }
};

template <bool D>
__host__ __device__ __forceinline__ double bayes(double p_hat, double n_obs, double alpha, double beta) {
    using F = F64<D>;
    const double pa = F::add(F::mul(n_obs, p_hat), alpha);                     // dp_solver.py:122
    const double pb = F::add(F::mul(n_obs, F::add(1.0, -p_hat)), beta);        // :123
    return F::div(pa, F::add(pa, pb));                                         // :126
}

template <bool D>
__host__ __device__ __forceinline__ int stop_rule_one(const double* p_in, const double* C, int L, double lam,
                                                      int risk, double alpha, double beta, double* J) {
    using F = F64<D>;
    double p_bar[kMaxStages + 1];
    p_bar[0] = 1.0;
    for (int i = 0; i < L; ++i) {
        const double pi = risk ? bayes<D>(p_in[i], 100.0, alpha, beta) : p_in[i];  // :40-41
        p_bar[i + 1] = F::mul(p_bar[i], pi);                                        // :44-46
    }
    J[L] = 0.0;
    int k_star = L - 1;                                                             // :69 default
    for (int i = L - 1; i >= 0; --i) {                                              // :53-66
        const double cost_if_stop = F::add(C[i], F::mul(lam, F::add(1.0, -p_bar[i + 1])));
        const double cost_if_continue = F::add(C[i], J[i + 1]);
        if (cost_if_stop <= cost_if_continue) {
            J[i] = cost_if_stop;
            k_star = i;  // the lowest flagged index wins because i decreases
        } else {
            J[i] = cost_if_continue;
        }
    }
    // k_star = first flagged stage; if none was flagged it is still L-1
    return k_star;
}

__global__ void stop_rule_kernel(const double* __restrict__ p, const double* __restrict__ C, int n, int L, double lam,
                                 const double* __restrict__ lam_rows, int risk, double alpha, double beta,
                                 int* __restrict__ k_star, double* __restrict__ J) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double Jl[kMaxStages + 1];
    const int k = stop_rule_one<true>(p + (size_t)r * L, C + (size_t)r * L, L, lam_rows ? lam_rows[r] : lam, risk,
                                      alpha, beta, Jl);
    k_star[r] = k;
    for (int i = 0; i <= L; ++i) J[(size_t)r * (L + 1) + i] = Jl[i];
}

// ---------------------------------------------------------------------------------------------------------
// Scorer -> stop decision on the device (north star 3): one CTA per request turns the per-token features the
// sampling kernel left in device memory (entropy / max-prob / margin / emitted-token log-prob over the FULL vocabulary)
// into the predictor's feature vector, runs the 256 -> 128 -> 1 MLP, the Bayesian shrinkage and the optimal-stopping
// DP, and emits (acceptance probability, stop?, k*).  The host only reads those three numbers per request: no logits,
// no per-token features and no MLP on the host.  Mirrors FeatureExtractor.extract + QualityPredictor.predict +
// bayesian_adjustment + optimal_stopping_rule as chained by the reference's loop
// (/root/reference/src/serving/pipeline.py:225-256; listing docs/guides/RESEARCH_PROTOCOL.md:366-409).
constexpr int kHidden = 128;

struct CascadeArgs {
    const float* features;     // [n, T, 6]
    const int* n_tokens;       // [n]
    const double* scalars;     // [n, 3] prompt words / 2048, output words / 512, stage / 4
    const float *w1, *b1, *w2, *b2;
    const double* prev_p;      // [n, L]
    const double* C;           // [L]
    double* prob;
    int* stop;
    int* k_star;
    int T, fdim, L, stage_idx, mode, risk;
    double lam, n_obs, alpha, beta;
};

__device__ __forceinline__ double block_sum128(double v, double* scratch) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    return (scratch[0] + scratch[1]) + (scratch[2] + scratch[3]);
}

__global__ void __launch_bounds__(kHidden) cascade_decide_kernel(const CascadeArgs a) {
    __shared__ double scratch[4];
    __shared__ float x[8];
    const int r = blockIdx.x, tid = threadIdx.x;
    const int nt = min(a.n_tokens[r], a.T);
    const float* f = a.features + (size_t)r * a.T * ASD_NUM_FEATURES;
    const double* sc = a.scalars + (size_t)r * 3;
    // ---- feature vector (FeatureExtractor.extract): means over the generated tokens in binary64
    double ent = 0.0, lpm = 0.0, mar = 0.0, lpt = 0.0, lmin = INFINITY;
    const int t0 = max(0, nt - 32);
    for (int t = tid; t < nt; t += kHidden) {
        const float* ft = f + (size_t)t * ASD_NUM_FEATURES;
        if (t >= t0) ent += (double)ft[3];
        lpm += log(fmin(fmax((double)ft[1], 1e-30), 1.0));
        mar += (double)ft[2];
        lpt += (double)ft[5];
        lmin = fmin(lmin, (double)ft[5]);
    }
    ent = block_sum128(ent, scratch);
    lpm = block_sum128(lpm, scratch);
    mar = block_sum128(mar, scratch);
    lpt = block_sum128(lpt, scratch);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, d));
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = lmin;
    __syncthreads();
    lmin = fmin(fmin(scratch[0], scratch[1]), fmin(scratch[2], scratch[3]));
    if (tid == 0) {
        const bool any = nt > 0;
        x[0] = any ? (float)(ent / (double)(nt - t0)) : 0.0f;
        x[1] = (float)sc[0];
        x[2] = (float)sc[1];
        x[3] = any ? (float)(lpm / (double)nt) : -10.0f;
        x[4] = (float)sc[2];
        x[5] = any ? (float)(mar / (double)nt) : 0.0f;
        x[6] = any ? (float)(lpt / (double)nt) : 0.0f;
        x[7] = any ? (float)lmin : 0.0f;
    }
    __syncthreads();
    // ---- MLP: hidden unit = thread; only the first 8 input dimensions are ever non-zero
    float h = a.b1[tid];
    const float* w = a.w1 + (size_t)tid * a.fdim;
#pragma unroll
    for (int d = 0; d < 8; ++d)
        if (d < a.fdim) h = fmaf(w[d], x[d], h);
    h = fmaxf(h, 0.0f);
    const double y = block_sum128((double)(a.w2[tid] * h), scratch);
    if (tid != 0) return;
    const bool last_stage = a.stage_idx >= a.L - 1;
    double prob = 1.0;                                                     // pipeline.py:241-242
    if (!last_stage) {
        const float z = (float)y + a.b2[0];
        prob = (double)(1.0f / (1.0f + expf(-z)));
        if (a.risk) prob = bayes<true>(prob, a.n_obs, a.alpha, a.beta);    // pipeline.py:234-238
    }
    a.prob[r] = prob;
    double p[kMaxStages], J[kMaxStages + 1];
    for (int i = 0; i < a.stage_idx; ++i) p[i] = a.prev_p[(size_t)r * a.L + i];
    p[a.stage_idx] = prob;
    if (a.mode == 1) {   // the reference's loop: DP on the prefix seen so far (pipeline.py:248-256)
        const int k = stop_rule_one<true>(p, a.C, a.stage_idx + 1, a.lam, 0, a.alpha, a.beta, J);
        a.stop[r] = k == a.stage_idx;
        a.k_star[r] = k;
    } else {             // the documented rule: DP over ALL stages, unseen stages at probability 1
        for (int i = a.stage_idx + 1; i < a.L; ++i) p[i] = 1.0;
        const int k = stop_rule_one<true>(p, a.C, a.L, a.lam, 0, a.alpha, a.beta, J);
        a.stop[r] = k <= a.stage_idx;
        a.k_star[r] = a.stage_idx;
    }
}

int launch_cascade_decide(const float* features, const int* n_tokens, int n, int T, const double* scalars, const float* w1,
                          const float* b1, const float* w2, const float* b2, int fdim, const double* prev_p,
                          const double* C, int L, int stage_idx, int mode, double lam, int risk, double n_obs,
                          double alpha, double beta, double* prob, int* stop, int* k_star, cudaStream_t stream) {
    if (L < 1 || L > kMaxStages || stage_idx < 0 || stage_idx >= L) return set_error("asd_cascade_decide: bad stage count");
    if (fdim < 8) return set_error("asd_cascade_decide: feature_dim must be >= 8");
    if (n <= 0) return 0;
    CascadeArgs a;
    a.features = features, a.n_tokens = n_tokens, a.scalars = scalars;
    a.w1 = w1, a.b1 = b1, a.w2 = w2, a.b2 = b2;
    a.prev_p = prev_p, a.C = C, a.prob = prob, a.stop = stop, a.k_star = k_star;
    a.T = T, a.fdim = fdim, a.L = L, a.stage_idx = stage_idx, a.mode = mode, a.risk = risk;
    a.lam = lam, a.n_obs = n_obs, a.alpha = alpha, a.beta = beta;
    cascade_decide_kernel<<<n, kHidden, 0, stream>>>(a);
    ASD_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
}

int launch_stop_rule(const double* p, const double* C, int n, int L, double lam, int risk_adjustment, double alpha,
                     double beta, int* k_star, double* J, cudaStream_t stream, const double* lam_rows) {
    if (L < 1 || L > kMaxStages) return set_error("asd_stop_rule: L must be in [1, %d]", kMaxStages);
    if (n <= 0) return 0;
    const int threads = 128, blocks = (n + threads - 1) / threads;
    stop_rule_kernel<<<blocks, threads, 0, stream>>>(p, C, n, L, lam, lam_rows, risk_adjustment, alpha, beta, k_star, J);
    ASD_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
}

int stop_rule_host(const double* p, const double* C, int L, double lam, int risk_adjustment, double alpha,
                   double beta, double* J) {
    if (L < 1 || L > kMaxStages) return set_error("asd_stop_rule_host: L must be in [1, %d]", kMaxStages);
    return stop_rule_one<false>(p, C, L, lam, risk_adjustment, alpha, beta, J);
}

int stop_rule_rows_host(const double* p, const double* C, const double* lam, int n, int L, int risk_adjustment,
                        double alpha, double beta, int* k_star, double* J) {
    if (L < 1 || L > kMaxStages) return set_error("asd_stop_rule_rows_host: L must be in [1, %d]", kMaxStages);
    if (!p || !C || !lam || !k_star) return set_error("asd_stop_rule_rows_host: NULL argument");
    double Jl[kMaxStages + 1];
    for (int r = 0; r < n; ++r) {
        k_star[r] = stop_rule_one<false>(p + (size_t)r * L, C + (size_t)r * L, L, lam[r], risk_adjustment, alpha, beta, Jl);
        if (J)
            for (int i = 0; i <= L; ++i) J[(size_t)r * (L + 1) + i] = Jl[i];
    }
    return 0;
}

double bayesian_adjustment_host(double p_hat, double n_obs, double alpha, double beta) {
    return bayes<false>(p_hat, n_obs, alpha, beta);
}

}  // namespace asd
