// Host-side launchers of the non-GEMM kernels of the forward (layers.cu, attention.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace asd {


int launch_add_norm(float* resid, const float* part, int nslices, size_t slice_stride, const int* tokens,
                    const __nv_bfloat16* emb, const __nv_bfloat16* w, __nv_bfloat16* xnorm, int M, int h, float eps,
                    cudaStream_t stream, __nv_bfloat16* resid_bf = nullptr, float* sumsq0 = nullptr);
int launch_tp_allreduce_norm(const float* const* peer_bufs, uint32_t* const* peer_flags, int rank, int world,
                             uint32_t epoch, int* error, float* resid, const __nv_bfloat16* w, __nv_bfloat16* xnorm,
                             int M, int h, float eps, __nv_bfloat16* resid_bf, float* sumsq0, cudaStream_t stream,
                             const float* const* peer_bcast = nullptr, uint32_t* const* peer_rowflags = nullptr);
// sums the slices in order into dst (nullptr: into slice 0)
int launch_reduce_slices(float* part, int nslices, size_t slice_stride, size_t n, cudaStream_t stream,
                         float* dst = nullptr);
int launch_qkv_rope(const float* part, int nslices, size_t slice_stride, const __nv_bfloat16* bias,
                    const int* positions, const int* token_slot, const int* page_table, int max_pages,
                    const float* inv_freq, __nv_bfloat16* q_out, __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, int M,
                    int nh, int nkv, int hd, int page_size, cudaStream_t stream);
int launch_rope_table(const int* positions, const float* inv_freq, float2* cs, int M, int half, cudaStream_t stream);
int launch_gather_rows(const __nv_bfloat16* src, const int* rows, __nv_bfloat16* dst, int n, int h,
                       cudaStream_t stream, const float* ss_src = nullptr, float* ss_dst = nullptr, int parts = 0,
                       int ld = 0);

struct AttnLaunch {
    const __nv_bfloat16* q;
    const __nv_bfloat16* k_cache;
    const __nv_bfloat16* v_cache;
    const int* positions;   // [M]
    const int* token_slot;  // [M]
    const int* cu_q;        // [nseq + 1]
    const int* seq_slot;    // [nseq]
    const int* page_table;  // [slots, max_pages]
    __nv_bfloat16* out;     // [M, nh, hd]
    float* o_part;
    float* ml_part;
    int* tickets;           // [nseq * nkv] zero-initialised, self-resetting
    int M, nseq, max_qlen, nh, nkv, hd, page_size, max_pages, split_keys, nsplit_max, impl;
    // impl 2 (tcgen05 kernel, attention_tc.cu): 2-D TMA map over the whole paged pool viewed as rows of head_dim
    // elements (box = 64 elements x 16 positions) and the first row of this layer's K / V
    const CUtensorMap* kv_map = nullptr;
    long long k_row0 = 0, v_row0 = 0;
    // L2 prefetch of the weights of the GEMMs that follow (O projection, then gate|up), issued by this kernel's CTAs
    // once the QKV GEMM upstream has finished: attention streams ~1.5 TB/s of K/V, the rest of the HBM bandwidth
    // is idle for its whole duration (engine option attn_prefetch_mb; mma.sync kernel only)
    struct WeightPrefetch {
        const CUtensorMap* tmap = nullptr;   // weight map of the GEMM (box = 64 k x 128 rows)
        int ntiles = 0, ksplit = 1, kblocks = 0, kp = 0;   // kp = k-blocks per CTA of that GEMM to prefetch
    } pf[2];
};
// impl 0: one-warp cross-check kernel; 1: mma.sync kernel; 2: tcgen05 kernel when the shape allows (head_dim 128,
// 16-position pages, q_len * group <= 128), else 1
int launch_attention(const AttnLaunch& L, cudaStream_t stream);
int launch_attention_tc(const AttnLaunch& L, cudaStream_t stream);

// load every kernel of the forward into the current context (see layers.cu: preload_layers)
int preload_layers();
int preload_attention();
int preload_attention_tc();
int preload_gemm();
int preload_gemm_tc();

}  // namespace asd
