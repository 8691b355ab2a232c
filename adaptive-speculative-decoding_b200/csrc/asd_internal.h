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

// Per-engine tuning knobs (engine options).  They used to be process globals, so an option set on the draft
// engine changed the target's plans; now every engine owns a Tuning and installs it for the duration of one
// forward() on the calling host thread (thread_local pointer), which also makes concurrent forwards of different
// engines from different host threads independent.
struct Tuning {
    int gemm_early_trigger = 0;
    int gemm_headroom = 1;
    int gemm_recv_dedicated = 1;
    int gemm_next_mb = 0;   // L2 budget (MB) for the next kernel's weights; measured neutral, off
    int glue_pdl = 1;       // launch the glue kernels programmatically (they wait on griddepcontrol)
    int attn_wide = 0;
    int attn_dbg = 0;       // diagnostics (tcgen05 attention): 1 = skip the softmax arithmetic, 2 = also skip the fold
    int gemm_big = 1;       // token counts above the HBM/tensor ridge use the 2-CTA tensor-bound GEMM
};
extern thread_local const Tuning* g_tuning;
inline const Tuning& tuning() {
    static const Tuning defaults;
    return g_tuning ? *g_tuning : defaults;
}
struct TuningScope {
    const Tuning* prev;
    explicit TuningScope(const Tuning* t) : prev(g_tuning) { g_tuning = t; }
    ~TuningScope() { g_tuning = prev; }
};

// cudaFuncSetAttribute is per device: remember which devices a kernel's attributes were set on
struct PerDeviceOnce {
    unsigned long long mask = 0;
    bool need(int* dev_out = nullptr) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev_out) *dev_out = dev;
        const unsigned long long bit = 1ull << (dev & 63);
        if (mask & bit) return false;
        mask |= bit;
        return true;
    }
};

// sampler.cu
size_t reject_sample_workspace_bytes(int B, int k);
int launch_reject_sample(const float* target, const float* draft, const int* draft_tokens, const double* u_accept,
                         const double* u_resid, int B, int k, int V, float temperature, uint8_t* accept_mask,
                         int* accepted_len, int* out_tokens, float* out_logprobs, float* features, void* workspace,
                         cudaStream_t stream);

extern int g_sampler_impl;

// stop_rule.cu
int launch_stop_rule(const double* p, const double* C, int n, int L, double lam, int risk_adjustment, double alpha,
                     double beta, int* k_star, double* J, cudaStream_t stream, const double* lam_rows = nullptr);
int launch_cascade_decide(const float* features, const int* n_tokens, int n, int T, const double* scalars, const float* w1,
                          const float* b1, const float* w2, const float* b2, int fdim, const double* prev_p,
                          const double* C, int L, int stage_idx, int mode, double lam, int risk, double n_obs,
                          double alpha, double beta, double* prob, int* stop, int* k_star, cudaStream_t stream);
int stop_rule_rows_host(const double* p, const double* C, const double* lam, int n, int L, int risk_adjustment,
                        double alpha, double beta, int* k_star, double* J);
int stop_rule_host(const double* p, const double* C, int L, double lam, int risk_adjustment, double alpha,
                   double beta, double* J);
double bayesian_adjustment_host(double p_hat, double n_obs, double alpha, double beta);

}  // namespace asd
