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
#define ASD_NUM_FEATURES 6 /* lse, p_max, margin, entropy, ln p(draft tok), ln p(token emitted at this position) */

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
 *   out_logprobs fp32 [B, k+1] (0 padded); features fp32 [B, k+1, ASD_NUM_FEATURES]; features[.., 5] repeats
 *   out_logprobs (the residual / bonus draw is only made for the position that emits the new token).
 *   workspace: asd_reject_sample_workspace_bytes(B, k) bytes, zero-filled once before first use.
 * V must be a multiple of 4 and <= 212992; k <= 64; logits 16-byte aligned.  k == 0 samples one
 * token per row from softmax(target/T) (used for the draft model's own sampling).
 */
ASD_API size_t asd_reject_sample_workspace_bytes(int B, int k);
/* tuning switch of the streaming sampler (bit flags, default 5: 4 = the draw kernel is gated per sequence, 2 = persistent
 * statistics CTAs, 8 = single acq_rel ticket); every setting implements the same arithmetic contract bit for bit */
ASD_API void asd_reject_sample_set_impl(int impl);
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
/* same with one lambda per row (lam fp64 [n]): OptimalStoppingTable.precompute evaluates its whole
 * lambda x probability grid with ONE call (src/algorithms/dp_solver.py:149-171 loops in Python) */
ASD_API int asd_stop_rule_rows(const double* p, const double* C, const double* lam, int n, int L, int risk_adjustment,
                       double alpha, double beta, int32_t* k_star, double* J, void* stream);
ASD_API int asd_stop_rule_rows_host(const double* p, const double* C, const double* lam, int n, int L, int risk_adjustment,
                            double alpha, double beta, int32_t* k_star, double* J);
/* Scorer -> stop decision in ONE launch, all on the device (north star 3; replaces the chain
 * FeatureExtractor.extract -> QualityPredictor.predict -> bayesian_adjustment -> optimal_stopping_rule of
 * src/serving/pipeline.py:225-256 for n requests that have just finished stage `stage_idx`):
 *   features fp32 [n, T, ASD_NUM_FEATURES]: the rows asd_reject_sample wrote for each request's generated tokens (still
 *     in device memory); n_tokens i32 [n]: how many of the T rows are valid;
 *   scalars fp64 [n, 3]: prompt words / 2048, output words / 512, stage / 4 (RESEARCH_PROTOCOL.md:389-403);
 *   w1 fp32 [128, feature_dim], b1 [128], w2 [128], b2 [1]: the 256 -> 128 -> 1 MLP (RESEARCH_PROTOCOL.md:325-331);
 *   prev_p fp64 [n, L]: acceptance probabilities of the stages before stage_idx; C fp64 [L]: stage costs;
 *   prefix_mode 1: the reference loop's rule (DP over the stages seen so far, stop iff k* == stage_idx, pipeline.py:248-256),
 *   0: DP over all L stages with unseen stages at probability 1 (stop iff k* <= stage_idx).
 * Outputs: prob fp64 [n] (after the optional Bayesian shrinkage with n_obs; 1.0 at the last stage), stop i32 [n],
 * k_star i32 [n].  The host reads three numbers per request; no per-token data leaves the device. */
ASD_API int asd_cascade_decide(const float* features, const int32_t* n_tokens, int n, int T, const double* scalars,
                       const float* w1, const float* b1, const float* w2, const float* b2, int feature_dim,
                       const double* prev_p, const double* C, int L, int stage_idx, int prefix_mode, double lam,
                       int risk_adjustment, double n_obs, double alpha, double beta, double* prob, int32_t* stop,
                       int32_t* k_star, void* stream);
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
 *               out bf16 [M, N/2] = silu(gate) * up;
 *   out_mode 3: out fp32 [M, N], final: the K splits of a tile form a thread-block cluster and are
 *               reduced through distributed shared memory in rank order (deterministic).
 *   ksplit / stages: 0 = let the library choose (single co-resident wave).
 */
ASD_API int asd_linear_bf16(const void* x, const void* w, void* out, int M, int N, int K, int out_mode, int ksplit,
                    int stages, int* ksplit_used, void* stream);
/* Same contraction on the tensor-bound kernel used above the HBM/tensor ridge (M > 256 tokens: BASELINE configs[4],
 * prefill): CTA pairs (tcgen05 cta_group::2, 256 weight rows x <= 256 tokens per tile), persistent, two TMEM
 * accumulators.  out_mode 0: fp32 [ksplit_used, M, N] slices; out_mode 2: SwiGLU bf16 [M, N/2]. */
ASD_API int asd_linear_bf16_tc(const void* x, const void* w, void* out, int M, int N, int K, int out_mode, int ksplit,
                       int stages, int* ksplit_used, void* stream);
ASD_API int asd_linear_plan(int M, int N, int K, int out_mode, int* ksplit, int* stages, int* token_tile);

/* ---------------------------------------------------------------------------------------------
 * Model forward engine (Qwen2 family: RMSNorm, RoPE, GQA attention with QKV bias, SwiGLU MLP).
 * Stands where the reference calls vllm.LLM(model, tensor_parallel_size, dtype="bfloat16",
 * max_model_len=4096).generate(...)  (src/serving/real_model_pipeline.py:98-108,135;
 * Stage.generate in docs/guides/RESEARCH_PROTOCOL.md:233-304).  One handle per (model, TP rank);
 * calls on one handle must be serialised by the caller (the Python Stage holds a lock).
 *
 * Dimensions in asd_model_config are LOCAL to the tensor-parallel rank: n_heads, n_kv_heads and ffn
 * are the global values divided by tp_size (column-parallel QKV / gate|up, row-parallel O / down);
 * hidden and vocab are global (lm_head is replicated).
 *
 * Weight layouts (bf16, row-major, torch nn.Linear [out, in]):
 *   wqkv [(n_heads + 2 n_kv_heads) * head_dim, hidden] = rows of q heads, then k heads, then v heads;
 *   bqkv [(n_heads + 2 n_kv_heads) * head_dim];  wo [hidden, n_heads * head_dim];
 *   wgateup [2 * ceil(ffn/64)*64, hidden]: per 128-row tile 64 gate rows then 64 up rows (zero padded);
 *   wdown [hidden, ffn];  ln1, ln2, final_norm [hidden];  embed, lm_head [vocab, hidden];
 *   inv_freq fp32 [head_dim / 2].
 * KV pool (bf16, caller allocated, asd_engine_kv_pool_bytes):
 *   [n_layers][2 (K, V)][num_pages][n_kv_heads][page_size][head_dim]; page_table i32
 *   [max_seqs, max_pages_per_seq] maps (sequence slot, position / page_size) -> page.
 */
typedef struct asd_model_config {
    int hidden, n_layers, n_heads, n_kv_heads, head_dim, ffn, vocab;
    float rms_eps;
    int max_tokens; /* largest M of one forward call */
    int page_size;
    int tp_rank, tp_size;
} asd_model_config;
typedef struct asd_engine asd_engine_t;

ASD_API asd_engine_t* asd_engine_create(const asd_model_config* cfg); /* NULL on error */
ASD_API void asd_engine_destroy(asd_engine_t* e);
ASD_API int asd_engine_set_layer(asd_engine_t* e, int layer, const void* wqkv, const void* bqkv, const void* wo,
                         const void* wgateup, const void* wdown, const void* ln1, const void* ln2);
ASD_API int asd_engine_set_globals(asd_engine_t* e, const void* embed, const void* final_norm, const void* lm_head,
                           const float* inv_freq);
ASD_API size_t asd_engine_kv_pool_bytes(const asd_model_config* cfg, int num_pages);
ASD_API int asd_engine_set_kv(asd_engine_t* e, void* kv_pool, int num_pages, const int32_t* page_table, int max_seqs,
                      int max_pages_per_seq);
/* comm: ncclComm_t; nccl_allreduce_fn: address of ncclAllReduce from the NCCL the process loaded.
 * Called (fp32 sum, in place) after the O and down projections when tp_size > 1. */
ASD_API int asd_engine_set_allreduce(asd_engine_t* e, void* comm, void* nccl_allreduce_fn);
/* Tensor-parallel boundary over NVLink peer memory (preferred to the NCCL path above): every rank exports
 * CUDA-IPC handles of its two receive buffers and its flag array (3 x 64 bytes), the host all-gathers them
 * (rank-major, world x 3 x 64 bytes) and each rank imports the table.  From then on the O / down projections
 * either finish the all-reduce inside their own epilogue (owner CTAs push {value, epoch} words into the peers'
 * receive buffers; option "tp_fused", default at 2 ranks / small exchanges) or are followed by ONE kernel that
 * all-reduces through P2P loads (one-shot, or two-shot with a home rank per row: option "tp_two_shot"), adds the
 * residual and emits the norm statistics.  All paths give bit-identical residuals on every rank.
 * asd_engine_tp_error returns 1 if a peer ever failed to show up within the spin bound. */
ASD_API int asd_engine_ipc_export(asd_engine_t* e, void* handles_out);
ASD_API int asd_engine_ipc_import(asd_engine_t* e, const void* all_handles);
ASD_API int asd_engine_tp_error(asd_engine_t* e);
/* In-process tensor parallelism (what Stage(tensor_parallel_size = t, gpu_ids = [...]) builds, mirroring
 * configs/qwen3_models.yaml:10-51 and src/serving/real_model_pipeline.py:98-108): the t rank engines were created in
 * ONE process, rank r on its own device.  Enables peer access between the devices and wires every rank's receive
 * buffers and flags into the others (no IPC, no torchrun).  engines[r] must be rank r of a t-way group. */
ASD_API int asd_engine_peer_connect(asd_engine_t** engines, int n);
/* options: "attn_impl" (2 tcgen05 / TMEM kernel - default; 1 mma.sync kernel, also the fallback for head_dim 64 or more
 * than 128 query rows; 0 one-warp cross-check kernel), "tc_qkvo" (1, default: above 256 tokens QKV and O run on the
 * CTA-pair tensor-bound GEMM like gate|up / down / lm_head; 0: weight-streaming kernel), "gemm_big" (0: never use the
 * CTA-pair GEMM), "attn_prefetch_mb" (L2 prefetch of the following GEMMs' weights from the attention kernel, measured
 * slower, default 0), "pdl" (0/1),
 * "reduce" (1 in-cluster split-K reduction with fused residual add, 0 fp32 slices summed by the glue
 * kernels), "fuse_rope" (1: bias + RoPE + q store + paged K/V append in the QKV GEMM epilogue),
 * "fuse_norm" (1: RMSNorm fused into the GEMMs: the O / down epilogues emit bf16(resid * ln_w) and the
 * per-token sum of squares, the consuming QKV / gate|up / lm_head epilogues apply rstd),
 * "attn_target_ctas", "attn_min_split_keys" (keys per flash-decoding split, default 1024), "glue_pdl",
 * "ksplit", "stages" (0 = automatic), "headroom" (shared memory left for the next kernel's CTAs on <= 32-token
 * tiles), "recv_dedicated", "early_trigger", "next_prefetch_mb" (tuning knobs),
 * "tp_fused" (0 never, 1 where it pays, 2 always), "tp_two_shot" (-1 never, 0 where it pays, 1 always), "p2p",
 * "persist" (1: single-rank forwards of <= 128 tokens run as ONE cooperative launch of the persistent forward
 * kernel, forward_persist.cu; 0, default: one launch per GEMM / attention - measured faster),
 * "persist_ahead" (L2 prefetch distance of the persistent kernel's weight producer, 0 = off), "profile" (0/1) */
ASD_API int asd_engine_set_option(asd_engine_t* e, const char* name, int value);
/* With option "profile" = 1 every launch is bracketed by CUDA events on the launching stream;
 * this call synchronises and returns the summed milliseconds and launch counts by kernel class
 * {0 weight-streaming GEMMs, 1 attention, 2 glue (norm/rope/gather), 3 TP all-reduce, 4 lm_head GEMM}
 * since the previous read (bench.py's roofline numbers come from here). nclass >= 5. */
/* Diagnostics: every following GEMM launch (up to max_launches, <= 1024 CTAs each) records per-CTA globaltimer
 * stamps into buf (device, u64): {entry, setup done, upstream grid done, first tile landed, main loop done,
 * cluster barrier 1, scatter done, cluster barrier 2, exit, smid}.  Returns the u64 stride per launch; buf = NULL
 * switches tracing off. */
ASD_API int asd_debug_gemm_trace(unsigned long long* buf, int max_launches);
/* Same for the attention kernel: {entry, upstream grid done, metadata read, loads issued, first K/V tile landed,
 * key loop done, key groups merged, partials stored, ticket taken, exit of the merging CTA}. */
ASD_API int asd_debug_attn_trace(unsigned long long* buf, int max_launches);
/* Diagnostics of the persistent forward kernel: per-CTA globaltimer stamps into buf (device, u64 [CTAs][643]: slot 0
 * kernel entry, slot p + 1 = this CTA's share of phase p done; phase 0 embedding, 1 + 5 l + {0 QKV, 1 attention, 2 O,
 * 3 gate|up, 4 down} of layer l, then logit-row gather and lm_head); returns the number of CTAs; NULL = off. */
ASD_API int asd_engine_persist_trace(asd_engine_t* e, unsigned long long* buf);
ASD_API int asd_engine_profile_read(asd_engine_t* e, float* ms_by_class, int* launches_by_class, int nclass);
/*
 * One forward pass over M tokens (draft step: q_len 1; verify step: q_len k+1; prefill chunk).
 *   tokens, positions, token_slot i32 [M]: token id, absolute position, sequence slot of each token;
 *     the tokens of one sequence are contiguous and in increasing position;
 *   cu_q i32 [nseq+1]: token range of each active sequence; seq_slot i32 [nseq]: its slot;
 *   max_qlen: largest q_len (q_len * n_heads / n_kv_heads <= 128); max_kv_len: upper bound of
 *     (last position + 1) over the sequences (sizes the flash-decoding split, may over-estimate);
 *   K/V of the M tokens are written in place into the paged pool before attention: a rejected
 *     speculative tail is rolled back by simply not advancing the caller's length (asd_kv_commit is
 *     therefore a no-op in this design);
 *   logit_rows i32 [n_logit_rows] selects rows (NULL = all M rows, n_logit_rows == M; 0 = no logits);
 *   logits_out fp32 [n_logit_rows, vocab] with row stride logits_ld elements (0 = vocab).
 * Asynchronous on `stream`; performs no allocation or synchronisation (CUDA-graph capturable after
 * one warm-up call with the same M when tp_size == 1; a tensor-parallel forward numbers its peer-memory exchanges
 * with a host-side epoch and refuses stream capture).  Runs on the device the engine was created on.
 */
ASD_API int asd_engine_forward(asd_engine_t* e, const int32_t* tokens, const int32_t* positions,
                       const int32_t* token_slot, int M, const int32_t* cu_q, const int32_t* seq_slot, int nseq,
                       int max_qlen, int max_kv_len, const int32_t* logit_rows, int n_logit_rows, float* logits_out,
                       long long logits_ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ASD_B200_H */
