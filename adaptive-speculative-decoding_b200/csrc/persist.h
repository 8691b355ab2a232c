// Host-side interface of the persistent forward kernel (forward_persist.cu): ONE cooperative launch runs the whole
// Qwen2 forward (embedding -> L x [QKV, attention, O, gate|up, down] -> lm_head) with one CTA per SM, the weight
// stream running through a shared-memory ring ACROSS the layer's GEMMs.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace asd {

// per-layer operands, one device array per engine (tensor maps live in global memory)
struct alignas(64) PLayer {
    CUtensorMap t_qkv, t_o, t_gu, t_down;    // weight maps, box = 64 x 128 (SWIZZLE_128B)
    const __nv_bfloat16 *bqkv, *ln1, *ln2;
    __nv_bfloat16 *k_cache, *v_cache;
    const void* pad_[3];
};

struct PersistLaunch {
    // model
    const PLayer* layers = nullptr;          // device
    int n_layers = 0, h = 0, nqkv = 0, qdim = 0, ffn = 0, ffp = 0, vocab = 0, nh = 0, nkv = 0, hd = 0, page_size = 0,
        max_pages = 0, Mx = 0;
    float eps = 0.f;
    const __nv_bfloat16 *embed = nullptr, *final_norm = nullptr, *lm_head = nullptr;
    const float* inv_freq = nullptr;
    const int* page_table = nullptr;
    // batch
    int M = 0, nseq = 0, max_qlen = 0, max_kv_len = 0;
    const int *tokens = nullptr, *positions = nullptr, *token_slot = nullptr, *cu_q = nullptr, *seq_slot = nullptr;
    // activations (engine-owned)
    float* resid = nullptr;
    __nv_bfloat16* resid_bf = nullptr;
    float *sumsq = nullptr, *sumsq_sel = nullptr;
    float2* rope_cs = nullptr;
    __nv_bfloat16 *q = nullptr, *attn = nullptr, *act = nullptr, *xsel = nullptr;
    // attention
    int split_keys = 0, nsplit_max = 1;
    float *o_part = nullptr, *ml_part = nullptr;
    int* tickets = nullptr;
    // logits
    const int* logit_rows = nullptr;
    int n_logit_rows = 0;
    float* logits = nullptr;
    long long logits_ld = 0;
    // synchronisation workspace (engine-owned): sync = [kPersistSyncWords] u32, zeroed by the launcher every forward
    unsigned* sync = nullptr;
    float* part_ws = nullptr;     // [CTAs][2][128 * 128] fp32 split-K partials (persist_part_ws_bytes)
    int prefetch_ahead = 0;       // weight blocks (16 KB) per CTA pulled into L2 beyond the shared-memory ring (measured: the
                                  // prefetches take the same SM->L2 request slots as the loads, main loops 1.6x slower; off)
    int* error = nullptr;         // sticky: set to 1 if a device-side wait ran into its bound
    unsigned long long* trace = nullptr;   // optional: [CTAs][kPersistMaxPhases] globaltimer stamps (slot 0 entry, p + 1 phase p done)
};

constexpr int kPersistMaxPhases = 3 + 5 * 128;           // embed + 5 per layer + gather + lm_head
constexpr int kPersistSyncWords = kPersistMaxPhases + 512;   // phase counters, then two partial flags per CTA

// 0 if this forward can run on the persistent kernel (single rank, <= 128 tokens, q_len * group <= 64, ...)
bool persist_supported(int M, int n_logit_rows, int max_qlen, int nh, int nkv, int hd, int page_size, int h, int n_layers);
int persist_num_ctas();
size_t persist_part_ws_bytes();
int persist_launch(const PersistLaunch& L, cudaStream_t stream);

}  // namespace asd
