// extern "C" surface of libasd_b200.so (declared in include/asd_b200.h).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "../../include/asd_b200.h"
#include "asd_internal.h"

namespace asd {

static thread_local char g_err[512] = "";
thread_local const Tuning* g_tuning = nullptr;
static std::atomic<long long> g_launches{0};

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return -1;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace asd

using namespace asd;

extern "C" {

int asd_abi_version(void) { return ASD_B200_ABI_VERSION; }
const char* asd_last_error(void) { return g_err; }
long long asd_launch_count(void) { return g_launches.load(); }
void asd_reset_launch_count(void) { g_launches.store(0); }

size_t asd_reject_sample_workspace_bytes(int B, int k) { return reject_sample_workspace_bytes(B, k); }
void asd_reject_sample_set_impl(int impl) { g_sampler_impl = impl; }

int asd_reject_sample(const float* target_logits, const float* draft_logits, const int32_t* draft_tokens,
                      const double* u_accept, const double* u_resid, int B, int k, int V, float temperature,
                      uint8_t* accept_mask, int32_t* accepted_len, int32_t* out_tokens, float* out_logprobs,
                      float* features, void* workspace, void* stream) {
    return launch_reject_sample(target_logits, draft_logits, draft_tokens, u_accept, u_resid, B, k, V, temperature,
                                accept_mask, accepted_len, out_tokens, out_logprobs, features, workspace,
                                static_cast<cudaStream_t>(stream));
}

int asd_reject_sample_host(const float* target_logits, const float* draft_logits, const int32_t* draft_tokens,
                           const double* u_accept, const double* u_resid, int B, int k, int V, float temperature,
                           uint8_t* accept_mask, int32_t* accepted_len, int32_t* out_tokens, float* out_logprobs,
                           float* features) {
    if (B <= 0) return 0;
    const size_t rows = (size_t)B * (k + 1), drows = (size_t)B * k;
    const size_t n_t = rows * V * 4, n_d = draft_logits ? drows * V * 4 : 0;
    const size_t ws_bytes = reject_sample_workspace_bytes(B, k);
    size_t offs[16], total = 0;
    const size_t sizes[] = {n_t, n_d, drows * 4, drows * 8, (size_t)B * 8, drows, (size_t)B * 4,
                            rows * 4, rows * 4, rows * ASD_NUM_FEATURES * 4, ws_bytes};
    for (int i = 0; i < 11; ++i) {
        offs[i] = total;
        total += (sizes[i] + 255) / 256 * 256;
    }
    uint8_t* d = nullptr;
    ASD_CUDA(cudaMalloc(&d, total + 256));
    int rc = 0;
    cudaStream_t s = nullptr;
#define H2D(i, src)                                                                                        \
    if (rc == 0 && sizes[i] && cudaMemcpyAsync(d + offs[i], src, sizes[i], cudaMemcpyHostToDevice, s) != cudaSuccess) \
        rc = set_error("asd_reject_sample_host: H2D copy failed");
    H2D(0, target_logits)
    if (draft_logits) { H2D(1, draft_logits) }
    H2D(2, draft_tokens)
    H2D(3, u_accept)
    H2D(4, u_resid)
#undef H2D
    if (rc == 0 && cudaMemsetAsync(d + offs[10], 0, ws_bytes, s) != cudaSuccess) rc = set_error("memset failed");
    if (rc == 0)
        rc = launch_reject_sample((const float*)(d + offs[0]), draft_logits ? (const float*)(d + offs[1]) : nullptr,
                                  (const int*)(d + offs[2]), (const double*)(d + offs[3]),
                                  (const double*)(d + offs[4]), B, k, V, temperature, d + offs[5],
                                  (int*)(d + offs[6]), (int*)(d + offs[7]), (float*)(d + offs[8]),
                                  (float*)(d + offs[9]), d + offs[10], s);
#define D2H(i, dst)                                                                                        \
    if (rc == 0 && sizes[i] && cudaMemcpyAsync(dst, d + offs[i], sizes[i], cudaMemcpyDeviceToHost, s) != cudaSuccess) \
        rc = set_error("asd_reject_sample_host: D2H copy failed");
    D2H(5, accept_mask)
    D2H(6, accepted_len)
    D2H(7, out_tokens)
    D2H(8, out_logprobs)
    D2H(9, features)
#undef D2H
    if (rc == 0) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = set_error("asd_reject_sample_host: %s", cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}

int asd_stop_rule(const double* p, const double* C, int n, int L, double lam, int risk_adjustment, double alpha,
                  double beta, int32_t* k_star, double* J, void* stream) {
    return launch_stop_rule(p, C, n, L, lam, risk_adjustment, alpha, beta, k_star, J,
                            static_cast<cudaStream_t>(stream));
}
int asd_stop_rule_rows(const double* p, const double* C, const double* lam, int n, int L, int risk_adjustment,
                       double alpha, double beta, int32_t* k_star, double* J, void* stream) {
    if (!lam) return set_error("asd_stop_rule_rows: lam is NULL");
    return launch_stop_rule(p, C, n, L, 0.0, risk_adjustment, alpha, beta, k_star, J, static_cast<cudaStream_t>(stream),
                            lam);
}
int asd_cascade_decide(const float* features, const int32_t* n_tokens, int n, int T, const double* scalars,
                       const float* w1, const float* b1, const float* w2, const float* b2, int feature_dim,
                       const double* prev_p, const double* C, int L, int stage_idx, int prefix_mode, double lam,
                       int risk_adjustment, double n_obs, double alpha, double beta, double* prob, int32_t* stop,
                       int32_t* k_star, void* stream) {
    if (!features || !n_tokens || !scalars || !w1 || !b1 || !w2 || !b2 || !prev_p || !C || !prob || !stop || !k_star)
        return set_error("asd_cascade_decide: NULL argument");
    return launch_cascade_decide(features, n_tokens, n, T, scalars, w1, b1, w2, b2, feature_dim, prev_p, C, L, stage_idx,
                                 prefix_mode, lam, risk_adjustment, n_obs, alpha, beta, prob, stop, k_star,
                                 static_cast<cudaStream_t>(stream));
}
int asd_stop_rule_rows_host(const double* p, const double* C, const double* lam, int n, int L, int risk_adjustment,
                            double alpha, double beta, int32_t* k_star, double* J) {
    return stop_rule_rows_host(p, C, lam, n, L, risk_adjustment, alpha, beta, k_star, J);
}
int asd_stop_rule_host(const double* p, const double* C, int L, double lam, int risk_adjustment, double alpha,
                       double beta, double* J) {
    return stop_rule_host(p, C, L, lam, risk_adjustment, alpha, beta, J);
}
double asd_bayesian_adjustment_host(double p_hat, double n_obs, double alpha, double beta) {
    return bayesian_adjustment_host(p_hat, n_obs, alpha, beta);
}

}  // extern "C"
