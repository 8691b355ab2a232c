// Host-side interface of the weight-streaming tcgen05 GEMM (gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace asd {

enum GemmOut { GEMM_OUT_F32 = 0, GEMM_OUT_BF16 = 1, GEMM_OUT_SWIGLU = 2, GEMM_OUT_QKV = 3 };

// Fused RMSNorm.  Producer side (fp32 accumulate epilogue of the O / down projection): besides the fp32
// residual v it writes x = bf16(v * ln_w[n]) (the next norm's weight applied per column, NOT yet divided by
// the rms) and this tile's sum of v^2 per token.  Consumer side: the GEMM reads x and its epilogue scales
// token m by rstd[m] = rsqrt(sum_t sumsq[t][m] / hidden + eps) - algebraically RMSNorm(v) * ln_w @ W^T with the
// same bf16 rounding points as the unfused kernels.
struct NormFusion {
    const float* sumsq_in = nullptr;   // consumer: [parts][ld]
    int parts = 0, ld = 0, hidden = 0;
    float eps = 0.f;
    float* sumsq_out = nullptr;        // producer: [n_tiles][ld]
    __nv_bfloat16* resid_bf = nullptr; // producer: [M][ldo] bf16(v * ln_w)
    const __nv_bfloat16* ln_w = nullptr; // producer: [N] weight of the NEXT RMSNorm
};

// Tensor-parallel row-parallel projection, fused: the owner CTAs of the in-cluster split-K reduction PUSH their
// fp32 partial rows into every peer's receive buffer over NVLink peer memory (slot = source rank) as
// {value, epoch} word pairs - the flag travels with the data, so there is no system-scope fence and no flag round
// trip - and then add the `world` partials in rank order (bit-identical on all ranks) onto the residual as soon as
// the peers' words carry this launch's epoch.  The all-reduce happens inside the GEMM epilogue, row by row, with no
// separate collective launch.  Receive buffers alternate with the epoch parity: a peer can only be one fused GEMM
// ahead (it needs my push of launch e+1 to finish it), so launch e+2 never overwrites rows I still read.
constexpr int kTpRowFlags = 4096;        // per-row flags of the two-shot all-reduce kernel (max tokens per forward)
struct TpFusion {
    float* recv[8] = {};       // this launch's receive buffer of every rank: [world][2 * slot_stride] words (peer-mapped)
    int rank = 0, world = 1;
    uint32_t epoch = 0;
    int* error = nullptr;      // set to 1 if a peer never showed up (bounded spin)
    size_t slot_stride = 0;    // floats between the per-source slots of a receive buffer
};

// extra operands of the fused QKV epilogue (bias + rotate-half RoPE + q store + paged K/V append)
struct QkvEpilogue {
    const float2* cs;               // [M, hd/2] (cos, sin) of each token's position
    const __nv_bfloat16* bias;      // [(nh + 2 nkv) * hd]
    const int* positions;           // [M]
    const int* token_slot;          // [M]
    const int* page_table;          // [slots, max_pages]
    __nv_bfloat16* q_out;           // [M, nh, hd]
    __nv_bfloat16* k_cache;
    __nv_bfloat16* v_cache;
    int max_pages, nh, nkv, hd, page_size;
};

struct GemmPlan {
    int M, N, K, mode;
    int MT, m_tiles, n_tiles, kblocks, ksplit, stages, smem_bytes, reduce, recv_dedicated;
    uint32_t tmem_cols;
};

// The next weight-streaming GEMM of the stream (see GemmArgs::next_*): filled by gemm_set_next() before a launch.
struct NextPrefetch {
    const CUtensorMap* tmap = nullptr;
    int ntiles = 0, ksplit = 1, kblocks = 0, kp = 0;
};
extern thread_local NextPrefetch g_gemm_next;
// prefetch budget -> k-blocks per CTA of `next`; call right before launching the GEMM that precedes `next`
void gemm_set_next(const GemmPlan& next, const CUtensorMap* next_w);
int gemm_token_tile(int M);
// reduce = 1: the K splits of a tile form a thread-block cluster and reduce through DSMEM, so the fp32
// output is final ([M][ldo], one slice) instead of one slice per split
int gemm_plan(GemmPlan* pl, int M, int N, int K, int mode, int force_ksplit, int force_stages, int reduce);
int make_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                   uint32_t box_rows);
// out: GEMM_OUT_F32 -> float [ksplit][M][ldo]; GEMM_OUT_BF16 -> bf16 [M][ldo];
// GEMM_OUT_SWIGLU -> bf16 [M][ldo] with N/2 columns (n_valid = ff)
int gemm_launch(const GemmPlan& pl, const CUtensorMap& tmap_w, const CUtensorMap& tmap_x, void* out, int ldo,
                int n_valid, bool pdl, cudaStream_t stream, bool accumulate = false,
                const QkvEpilogue* qkv = nullptr, const NormFusion* norm = nullptr, const TpFusion* tp = nullptr);

// ---- tensor-bound 2-CTA kernel (gemm_tc.cu): token counts above the HBM/tensor ridge
struct TcPlan {
    int M, N, K, mode;
    int MT, m_tiles, w_tiles, kblocks, ksplit, stages, smem_bytes, pairs;
};
int gemm_tc_plan(TcPlan* pl, int M, int N, int K, int mode, int force_ksplit, int force_stages);
// x tensor map: box rows = MT / 2 (each CTA of the pair loads half of the token tile); w tensor map: box rows = 128.
// GEMM_OUT_F32: out fp32 [ksplit][slice_stride] slices of [M][ldo]; GEMM_OUT_SWIGLU: out bf16 [M][ldo].
int gemm_tc_launch(const TcPlan& pl, const CUtensorMap& tmap_w, const CUtensorMap& tmap_x, void* out, int ldo, int n_valid,
                   size_t slice_stride, bool pdl, cudaStream_t stream, bool accumulate, const NormFusion* norm);

}  // namespace asd
