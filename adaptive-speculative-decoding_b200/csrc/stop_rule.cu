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
