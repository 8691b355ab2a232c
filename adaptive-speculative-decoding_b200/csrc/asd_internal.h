// Internal declarations shared by the translation units of libasd_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define ASD_NUM_FEATURES 6

namespace asd {

int set_error(const char* fmt, ...);  // records the message, returns -1
void count_launch(int n);

#define ASD_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::asd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// Every kernel of a forward pass asks for the same (maximum) shared-memory carve-out: an SM only re-partitions
// L1/shared memory when it is idle, so kernels with different carve-outs cannot be co-resident and a
// programmatically launched GEMM could not start streaming weights next to the attention CTAs.
// ASD_CARVEOUT=0 in the environment leaves the driver's per-kernel choice (for A/B runs).
template <typename F>
inline void prefer_max_smem(F* fn) {
    static const bool on = []() {
        const char* v = getenv("ASD_CARVEOUT");
        return !(v && v[0] == '0');
    }();
    if (on) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// sampler.cu
size_t reject_sample_workspace_bytes(int B, int k);
int launch_reject_sample(const float* target, const float* draft, const int* draft_tokens, const double* u_accept,
                         const double* u_resid, int B, int k, int V, float temperature, uint8_t* accept_mask,
                         int* accepted_len, int* out_tokens, float* out_logprobs, float* features, void* workspace,
                         cudaStream_t stream);

extern int g_sampler_impl;

// stop_rule.cu
int launch_stop_rule(const double* p, const double* C, int n, int L, double lam, int risk_adjustment, double alpha,
                     double beta, int* k_star, double* J, cudaStream_t stream);
int stop_rule_host(const double* p, const double* C, int L, double lam, int risk_adjustment, double alpha,
                   double beta, double* J);
double bayesian_adjustment_host(double p_hat, double n_obs, double alpha, double beta);

}  // namespace asd
