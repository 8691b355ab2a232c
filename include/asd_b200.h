/*
 * asd_b200.h - C ABI of libasd_b200.so, the B200-native (sm_100a) draft-then-verify hot path
 * behind the adaptive-speculative-decoding Python API.
 *
 * The reference (sa2shun/adaptive-speculative-decoding, /root/reference) is 100 % Python and has
 * NO FFI for this path: its engine seam is duck-typed Python that delegates to vLLM
 * (src/serving/real_model_pipeline.py:98-108,135; docs/guides/RESEARCH_PROTOCOL.md:233-304).
 * Each entry point below therefore cites the Python interface it stands behind; the ctypes
 * binding a maintainer of the reference would add is shown in INTEGRATION.md.
 *
 * Conventions: every function returns 0 on success and -1 on error (message from
 * asd_last_error(), thread local).  Unless a name ends in _host, pointers are CUDA device
 * pointers owned by the caller; the library never frees caller memory.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  No torch types appear here.
 */
#ifndef ASD_B200_H
#define ASD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(ASD_BUILDING_DSO)
#define ASD_API __attribute__((visibility("default")))
#else
#define ASD_API
#endif

#define ASD_B200_ABI_VERSION 1
#define ASD_NUM_FEATURES 6 /* lse, p_max, margin, entropy, ln p(draft tok), ln p(resampled tok) */

ASD_API int asd_abi_version(void);
ASD_API const char* asd_last_error(void);
/* number of kernels this library has launched since load / since the last reset (bench "gpu_launches") */
ASD_API long long asd_launch_count(void);
ASD_API void asd_reset_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Fused logits -> softmax -> rejection sampling -> residual resample -> stop-rule features.
 * Replaces the per-token softmax/log + .item() loop of
 * src/training/generate_training_data.py:128-134 and the logprob reductions of
 * FeatureExtractor.extract (docs/guides/RESEARCH_PROTOCOL.md:379-398); the accept/resample rule is
 * the canonical speculative sampling of SURVEY.md Appendix C (no reference implementation).
 *
 *   target_logits fp32 [B, k+1, V]   draft_logits fp32 [B, k, V] (may be NULL if k == 0 or T <= 0)
 *   draft_tokens  i32  [B, k]        u_accept fp64 [B, k]        u_resid fp64 [B]
 *   temperature <= 0 selects greedy verification (accept iff draft token == argmax).
 *   accept_mask u8 [B, k]; accepted_len i32 [B]; out_tokens i32 [B, k+1] (-1 padded);
 *   out_logprobs fp32 [B, k+1] (0 padded); features fp32 [B, k+1, ASD_NUM_FEATURES].
 *   workspace: asd_reject_sample_workspace_bytes(B, k) bytes, zero-filled once before first use.
 * V must be a multiple of 4 and <= 212992; k <= 64; logits 16-byte aligned.  k == 0 samples one
 * token per row from softmax(target/T) (used for the draft model's own sampling).
 */
ASD_API size_t asd_reject_sample_workspace_bytes(int B, int k);
ASD_API int asd_reject_sample(const float* target_logits, const float* draft_logits, const int32_t* draft_tokens,
                      const double* u_accept, const double* u_resid, int B, int k, int V, float temperature,
                      uint8_t* accept_mask, int32_t* accepted_len, int32_t* out_tokens, float* out_logprobs,
                      float* features, void* workspace, void* stream);
/* Same call with HOST buffers: copies in, runs the kernel on the current device, copies out. */
ASD_API int asd_reject_sample_host(const float* target_logits, const float* draft_logits, const int32_t* draft_tokens,
                           const double* u_accept, const double* u_resid, int B, int k, int V, float temperature,
                           uint8_t* accept_mask, int32_t* accepted_len, int32_t* out_tokens, float* out_logprobs,
                           float* features);

/* ---------------------------------------------------------------------------------------------
 * Cascade stop rule.  optimal_stopping_rule / bayesian_adjustment of
 * src/algorithms/dp_solver.py:12-71,106-130, bit-exact (binary64, same operation order, `<=` tie).
 *   p, C fp64 [n, L]; k_star i32 [n]; J fp64 [n, L+1].  L <= 64.
 */
ASD_API int asd_stop_rule(const double* p, const double* C, int n, int L, double lam, int risk_adjustment, double alpha,
                  double beta, int32_t* k_star, double* J, void* stream);
/* scalar host form used by the Python policy seam (pipeline.py:251-256): returns k_star or -1 */
ASD_API int asd_stop_rule_host(const double* p, const double* C, int L, double lam, int risk_adjustment, double alpha,
                       double beta, double* J);
ASD_API double asd_bayesian_adjustment_host(double p_hat, double n_obs, double alpha, double beta);

/* ---------------------------------------------------------------------------------------------
 * Weight-streaming linear layer  Y[m, n] = sum_k X[m, k] * W[n, k]  (bf16 in, fp32 accumulate) on
 * tcgen05/TMEM fed by TMA: the dense contraction of the verify / draft forward that the reference
 * delegates to vLLM (src/serving/real_model_pipeline.py:98-108,135).
 *   x bf16 [M, K]; w bf16 [N, K] (torch nn.Linear layout); K % 8 == 0; 16-byte aligned.
 *   out_mode 0: out fp32 [ksplit_used, M, N] - one slice per K split, the caller (or the fused
 *               add+RMSNorm / RoPE kernels) sums the slices in order;
 *   out_mode 1: out bf16 [M, N] (ksplit forced to 1);
 *   out_mode 2: SwiGLU - w rows interleaved per 128-row tile as 64 gate rows then 64 up rows;
 *               out bf16 [M, N/2] = silu(gate) * up.
 *   ksplit / stages: 0 = let the library choose (single co-resident wave).
 */
ASD_API int asd_linear_bf16(const void* x, const void* w, void* out, int M, int N, int K, int out_mode, int ksplit,
                    int stages, int* ksplit_used, void* stream);
ASD_API int asd_linear_plan(int M, int N, int K, int out_mode, int* ksplit, int* stages, int* token_tile);

#ifdef __cplusplus
}
#endif
#endif /* ASD_B200_H */
