// Paged-KV attention for the verify step on the legacy tensor path (mma.sync): the FALLBACK of the tcgen05 / TMEM kernel
// in attention_tc.cu (engine option attn_impl = 2, default) for head_dim 64 and for more than 128 query rows per
// (sequence, kv head), and its cross-check (attn_impl = 1).  Every (k+1) new position of a sequence is scored against
// the paged prefix in ONE pass (causal among the new positions); long contexts (>= 1024 keys per piece) are
// split over the KV length (flash-decoding) and merged by the last CTA to finish.
//
//   * work item = (sequence, kv head, kv split).  The query tile is all new tokens x the GQA group
//     of that kv head (rows = q_len * G, e.g. 6 x 5 = 30 for Qwen2.5-32B with k = 5), so each K/V
//     byte is read from HBM once per sequence, not once per query head;
//   * QK^T and PV run on tensor cores through mma.sync m16n8k16 bf16 -> fp32 (about an eighth of the tcgen05 rate:
//     HMMA issue bounds the kernel once the context is long - 0.23 of HBM at a 4096-token prefix against 0.65-0.72
//     for attention_tc.cu), operands staged with cp.async into XOR-swizzled shared memory and read with ldmatrix;
//   * with few query rows (draft steps: 7 rows) the warps of a CTA share the rows and split every
//     64-key tile between them, then merge through shared memory, so that a CTA always has 4+ warps
//     issuing loads; the last CTA of a (sequence, kv head) to finish merges the kv splits (atomic
//     ticket), so there is no separate combine launch;
//   * programmatic dependent launch: the batch description, the CTA's slice of the page table (staged in
//     shared memory) and the K/V tiles of keys older than this step's tokens are fetched BEFORE
//     griddepcontrol.wait, i.e. while the QKV GEMM upstream is still in its epilogue;
//   * K/V were appended in place by the QKV GEMM's epilogue (or qkv_rope_kernel on the unfused path) before
//     this kernel's wait returns; rejected speculative positions are simply overwritten by the next step
//     (rollback = not advancing the length).
// attn_simple_kernel is a one-warp-per-(token, head) restatement used to cross-check the tensor-core
// kernel on the device (engine option attn_impl = 0).
//
// The reference has no attention code (vLLM does it: real_model_pipeline.py:98-108); semantics are
// HF Qwen2 GQA attention with softmax scale 1/sqrt(head_dim).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "asd_internal.h"
#include "layers.h"
#include "mma_sync.cuh"
#include "ptx.cuh"

namespace asd {

// ------------------------------------------------------------------------------------------ simple
__global__ void __launch_bounds__(32) attn_simple_kernel(const __nv_bfloat16* __restrict__ q,
                                                         const __nv_bfloat16* __restrict__ k_cache,
                                                         const __nv_bfloat16* __restrict__ v_cache,
                                                         const int* __restrict__ positions,
                                                         const int* __restrict__ token_slot,
                                                         const int* __restrict__ page_table, int max_pages,
                                                         __nv_bfloat16* __restrict__ out, int nh, int nkv, int hd,
                                                         int page_size, float scale_log2) {
    const int m = blockIdx.x, hq = blockIdx.y, lane = threadIdx.x;
    const int g = hq / (nh / nkv), D = hd / 32;
    const int pos = positions[m], slot = token_slot[m];
    float qv[4], acc[4] = {0, 0, 0, 0};
    for (int d = 0; d < D; ++d) qv[d] = __bfloat162float(q[((size_t)m * nh + hq) * hd + lane * D + d]);
    float mx = -INFINITY, l = 0.0f;
    for (int j = 0; j <= pos; ++j) {
        const int page = page_table[(size_t)slot * max_pages + j / page_size];
        const size_t off = (((size_t)page * nkv + g) * page_size + j % page_size) * hd + lane * D;
        float s = 0.0f;
        for (int d = 0; d < D; ++d) s += qv[d] * __bfloat162float(k_cache[off + d]);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s *= scale_log2;
        const float mn = fmaxf(mx, s), corr = exp2f(mx - mn), p = exp2f(s - mn);
        l = l * corr + p;
        for (int d = 0; d < D; ++d) acc[d] = acc[d] * corr + p * __bfloat162float(v_cache[off + d]);
        mx = mn;
    }
    for (int d = 0; d < D; ++d) out[((size_t)m * nh + hq) * hd + lane * D + d] = __float2bfloat16(acc[d] / l);
}

// ------------------------------------------------------------------------------------------ tensor-core
constexpr int kKeyTile = 64;

struct AttnArgs {
    const __nv_bfloat16* q;        // [M, nh, hd]
    const __nv_bfloat16* k_cache;  // [pages, nkv, page, hd]
    const __nv_bfloat16* v_cache;
    const int* positions;          // [M]
    const int* cu_q;               // [nseq + 1] token ranges
    const int* seq_slot;           // [nseq]
    const int* page_table;
    int max_pages, nh, nkv, page_size, split_keys, nsplit_max, rg_count, kg_count;
    int page_shift;                // log2(page_size) when it is a power of two, else -1 (division fallback)
    float scale_log2;
    float* o_part;                 // [M, nh, nsplit_max, hd]
    float* ml_part;                // [M, nh, nsplit_max, 2]
    int* tickets;                  // [nseq_max * nkv], zero between launches
    __nv_bfloat16* out;            // [M, nh, hd]
    unsigned long long* trace;     // diagnostics: 16 globaltimer stamps per CTA (nullptr = off)
    // weights of the following GEMMs to pull into L2 while this kernel runs (see AttnLaunch::WeightPrefetch)
    int pf_ntiles[2], pf_ksplit[2], pf_kblocks[2], pf_kp[2];
};

__device__ __forceinline__ void attn_stamp(const AttnArgs& a, int slot) {
    if (a.trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        a.trace[(size_t)cta * 16 + slot] = t;
    }
}

// L2 prefetch of the weights the next GEMMs will stream: 16 KB boxes, k-block-major (the first k-block of every
// CTA of that GEMM first, so all of its CTAs find the same share of their slab in L2), dealt round-robin over
// this grid's CTAs and `nissue` threads per CTA
__device__ __forceinline__ void attn_prefetch_weights(const AttnArgs& a, int i, const CUtensorMap* tmap, int who,
                                                      int nissue) {
    const int ntiles = a.pf_ntiles[i], ksplit = a.pf_ksplit[i], kblocks = a.pf_kblocks[i], kp = a.pf_kp[i];
    if (kp <= 0) return;
    const int ncta_next = ntiles * ksplit, nbox = ncta_next * kp;
    const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const int nthr = gridDim.x * gridDim.y * gridDim.z * nissue;
    const int nbase = kblocks / ksplit, nrem = kblocks % ksplit;
    for (int b = cta * nissue + who; b < nbox; b += nthr) {
        const int j = b / ncta_next, c = b - j * ncta_next;
        const int nt = c % ntiles, ns = c / ntiles;
        const int nkb_c = nbase + (ns < nrem ? 1 : 0);
        if (j >= nkb_c) continue;
        const int kb = ns * nbase + (ns < nrem ? ns : nrem) + j;
        tma_prefetch_l2_2d(tmap, kb * 64, nt * 128);
    }
}

// HD = head_dim (64 or 128); KW = keys of each 64-key tile handled by one warp (64, 32 or 16).
// warp w = (row group w % rg_count, key group w / rg_count): 16 query rows x KW keys per tile.
// NS = depth of the cp.async K/V ring (4 when there is one CTA per SM, 2 otherwise).
template <int HD, int KW, int NS>
__global__ void __launch_bounds__(256) attn_mma_kernel(const AttnArgs a, const __grid_constant__ CUtensorMap tmap_pf0,
                                                       const __grid_constant__ CUtensorMap tmap_pf1) {
    constexpr int CH = HD / 8;          // 16-byte chunks per row
    constexpr int ROWB = HD * 2;        // bytes per row
    constexpr int KB = HD / 16;         // k-blocks over head_dim
    constexpr int NT = KW / 8;          // score n-tiles per warp
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ int s_last;
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rg = warp % a.rg_count, kg = warp / a.rg_count;
    uint8_t* sQ = smem_raw;                                   // [rg_count*16][HD]
    uint8_t* sK = sQ + (size_t)a.rg_count * 16 * ROWB;        // [NS][64][HD]
    uint8_t* sV = sK + NS * kKeyTile * ROWB;                  // [NS][64][HD]
    const uint32_t uQ = smem_u32(sQ), uK = smem_u32(sK), uV = smem_u32(sV);

    const int seq = blockIdx.x, g = blockIdx.y, sp = blockIdx.z;
    const int G = a.nh / a.nkv;
    attn_stamp(a, 0);
    // Everything up to the PDL wait reads only what earlier forwards (or the host) wrote: the batch
    // description, the page table and the K/V of keys older than this step's tokens.  The QKV GEMM that is
    // still running upstream only produces q and the K/V of the last qlen positions.
    const int q0 = a.cu_q[seq], qlen = a.cu_q[seq + 1] - q0;
    if (qlen <= 0) return;
    const int R = qlen * G;
    const int kv_len = a.positions[q0 + qlen - 1] + 1;
    const int kbeg = sp * a.split_keys;
    if (kbeg >= kv_len) return;
    const int kend = min(kv_len, kbeg + a.split_keys);
    const int nsplit_seq = (kv_len + a.split_keys - 1) / a.split_keys;
    const int old_keys = kv_len - qlen;
    // this CTA's slice of the sequence's page table, staged once (a lookup per key row would be a dependent
    // L2 round trip in front of every cp.async batch)
    int* s_pt = reinterpret_cast<int*>(sV + NS * kKeyTile * ROWB);
    const int page0 = kbeg / a.page_size;
    {
        const int* pt = a.page_table + (size_t)a.seq_slot[seq] * a.max_pages;
        const int npages = (kend - 1) / a.page_size - page0 + 1;
        for (int i = threadIdx.x; i < npages; i += blockDim.x) s_pt[i] = pt[page0 + i];
    }
    __syncthreads();
    attn_stamp(a, 1);

    // 4 threads per key row: CH/4 16-byte chunks of K and V each
    auto load_tile = [&](int tile, int buf) {
        const int j0 = kbeg + tile * kKeyTile;
        const int sub = threadIdx.x & 3;
        for (int r = threadIdx.x >> 2; r < kKeyTile; r += blockDim.x >> 2) {
            const int j = j0 + r;
            const bool ok = j < kend;
            const int jj = ok ? j : kbeg;
            const int pidx = a.page_shift >= 0 ? jj >> a.page_shift : jj / a.page_size;
            const int pin = a.page_shift >= 0 ? jj & (a.page_size - 1) : jj - pidx * a.page_size;
            const int page = s_pt[pidx - page0];
            const size_t off = (((size_t)page * a.nkv + g) * a.page_size + pin) * HD;
            const uint32_t drow = (uint32_t)(buf * kKeyTile * ROWB + r * ROWB);
#pragma unroll
            for (int cc = 0; cc < CH / 4; ++cc) {
                const int ch = sub * (CH / 4) + cc;
                const uint32_t d = drow + ((ch ^ (r & 7)) << 4);
                cp_async16(uK + d, a.k_cache + off + ch * 8, ok);
                cp_async16(uV + d, a.v_cache + off + ch * 8, ok);
            }
        }
    };
    const int ntiles = (kend - kbeg + kKeyTile - 1) / kKeyTile;
    auto tile_is_old = [&](int t) { return kbeg + (t + 1) * kKeyTile <= old_keys; };
#pragma unroll
    for (int t = 0; t < NS - 1; ++t)     // prologue, part 1: the old tiles among the first NS-1
        if (t < ntiles && tile_is_old(t)) load_tile(t, t);
    cp_async_commit();
    attn_stamp(a, 2);
    grid_dep_wait();   // q and the new K/V come from the QKV GEMM launched just before
    grid_dep_launch();
    if (threadIdx.x < 4) {   // HBM is otherwise idle from here on: pull the next GEMMs' weights into L2
        attn_prefetch_weights(a, 0, &tmap_pf0, threadIdx.x, 4);
        attn_prefetch_weights(a, 1, &tmap_pf1, threadIdx.x, 4);
    }

    // ---- stage Q (rows r = t*G + gq -> token q0+t, head g*G+gq), swizzled
    for (int c = threadIdx.x; c < a.rg_count * 16 * CH; c += blockDim.x) {
        const int r = c / CH, ch = c - r * CH;
        const bool ok = r < R;
        const int t = ok ? r / G : 0, gq = ok ? r - t * G : 0;
        const __nv_bfloat16* src = a.q + ((size_t)(q0 + t) * a.nh + g * G + gq) * HD + ch * 8;
        cp_async16(uQ + r * ROWB + ((ch ^ (r & 7)) << 4), src, ok);
    }
#pragma unroll
    for (int t = 0; t < NS - 1; ++t)     // prologue, part 2: tiles that hold this step's keys
        if (t < ntiles && !tile_is_old(t)) load_tile(t, t);
    cp_async_commit();

    attn_stamp(a, 3);
    const int r0 = rg * 16 + (lane >> 2), r1 = r0 + 8;
    const int qpos0 = kv_len - qlen + (r0 < R ? r0 / G : 0), qpos1 = kv_len - qlen + (r1 < R ? r1 / G : 0);
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
    float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    uint32_t qf[KB][4];

    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile % NS;
        if (tile == 0) cp_async_wait<0>();   // the whole prologue (NS-1 tiles and Q)
        if (tile + NS - 1 < ntiles) load_tile(tile + NS - 1, (tile + NS - 1) % NS);
        cp_async_commit();
        if (tile > 0) cp_async_wait<NS - 1>();
        __syncthreads();
        if (tile == 0) attn_stamp(a, 4);
        if (tile == 0) {
            const int mi = lane >> 3;
            const int row = rg * 16 + (lane & 7) + (mi & 1) * 8;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                const int ch = kb * 2 + (mi >> 1);
                ldsm_x4(uQ + row * ROWB + ((ch ^ (row & 7)) << 4), qf[kb]);
            }
        }
        const int key0 = kg * KW;  // first key of this warp inside the tile
        if (kbeg + tile * kKeyTile + key0 < kend) {   // warp-uniform: skip key groups past the end
            float s[NT][4];
#pragma unroll
            for (int i = 0; i < NT; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f;
            const uint32_t kb_base = uK + buf * kKeyTile * ROWB, vb_base = uV + buf * kKeyTile * ROWB;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                for (int np = 0; np < KW / 16; ++np) {
                    const int mi = lane >> 3;
                    const int row = key0 + np * 16 + (lane & 7) + (mi >> 1) * 8;
                    const int ch = kb * 2 + (mi & 1);
                    uint32_t b[4];
                    ldsm_x4(kb_base + row * ROWB + ((ch ^ (row & 7)) << 4), b);
                    mma_bf16(s[np * 2], qf[kb], b[0], b[1]);
                    mma_bf16(s[np * 2 + 1], qf[kb], b[2], b[3]);
                }
            }
            const int jbase = kbeg + tile * kKeyTile + key0 + (lane & 3) * 2;
            float tm0 = -INFINITY, tm1 = -INFINITY;
            // tiles that end at or before the first new position are visible to every query row: no causal / length
            // mask (only the rows past R stay masked)
            const bool interior = kbeg + (tile + 1) * kKeyTile <= old_keys + 1;
            if (interior) {
                const float sc0 = r0 < R ? a.scale_log2 : -INFINITY, sc1 = r1 < R ? a.scale_log2 : -INFINITY;
#pragma unroll
                for (int i = 0; i < NT; ++i) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        s[i][e] = r0 < R ? s[i][e] * sc0 : -INFINITY;
                        s[i][2 + e] = r1 < R ? s[i][2 + e] * sc1 : -INFINITY;
                        tm0 = fmaxf(tm0, s[i][e]);
                        tm1 = fmaxf(tm1, s[i][2 + e]);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < NT; ++i) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = jbase + i * 8 + e;
                        const bool in = j < kend;
                        s[i][e] = (in && r0 < R && j <= qpos0) ? s[i][e] * a.scale_log2 : -INFINITY;
                        s[i][2 + e] = (in && r1 < R && j <= qpos1) ? s[i][2 + e] * a.scale_log2 : -INFINITY;
                        tm0 = fmaxf(tm0, s[i][e]);
                        tm1 = fmaxf(tm1, s[i][2 + e]);
                    }
                }
            }
            tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 1));
            tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 2));
            tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 1));
            tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 2));
            const float mn0 = fmaxf(mx0, tm0), mn1 = fmaxf(mx1, tm1);
            const float ref0 = mn0 == -INFINITY ? 0.0f : mn0, ref1 = mn1 == -INFINITY ? 0.0f : mn1;
            const float c0 = exp2f(mx0 - ref0), c1 = exp2f(mx1 - ref1);
            float ps0 = 0.0f, ps1 = 0.0f;
            uint32_t pf[KW / 16][4];
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const float p00 = exp2f(s[i][0] - ref0), p01 = exp2f(s[i][1] - ref0);
                const float p10 = exp2f(s[i][2] - ref1), p11 = exp2f(s[i][3] - ref1);
                ps0 += p00 + p01;
                ps1 += p10 + p11;
                pf[i >> 1][(i & 1) * 2] = pack_bf16(p00, p01);
                pf[i >> 1][(i & 1) * 2 + 1] = pack_bf16(p10, p11);
            }
            l0 = l0 * c0 + ps0;
            l1 = l1 * c1 + ps1;
            mx0 = mn0;
            mx1 = mn1;
#pragma unroll
            for (int i = 0; i < HD / 8; ++i) {
                o[i][0] *= c0;
                o[i][1] *= c0;
                o[i][2] *= c1;
                o[i][3] *= c1;
            }
#pragma unroll
            for (int kk = 0; kk < KW / 16; ++kk) {
#pragma unroll
                for (int dn = 0; dn < HD / 16; ++dn) {
                    const int mi = lane >> 3;
                    const int row = key0 + kk * 16 + (lane & 7) + (mi & 1) * 8;
                    const int ch = dn * 2 + (mi >> 1);
                    uint32_t b[4];
                    ldsm_x4_t(vb_base + row * ROWB + ((ch ^ (row & 7)) << 4), b);
                    mma_bf16(o[dn * 2], pf[kk], b[0], b[1]);
                    mma_bf16(o[dn * 2 + 1], pf[kk], b[2], b[3]);
                }
            }
        }
        __syncthreads();  // everyone done with this buffer before it is refilled
    }
    cp_async_wait<0>();
    attn_stamp(a, 5);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // ---- merge the key groups of a row group through shared memory (K/V ring is idle now)
    float* xo = reinterpret_cast<float*>(sK);                 // [nwarps][16][HD]
    float* xml = xo + (size_t)nwarps * 16 * HD;               // [nwarps][16][2]
    if (a.kg_count > 1) {
        __syncthreads();
        float* wo = xo + (size_t)warp * 16 * HD;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
            const int d = i * 8 + (lane & 3) * 2;
            *reinterpret_cast<float2*>(wo + (lane >> 2) * HD + d) = make_float2(o[i][0], o[i][1]);
            *reinterpret_cast<float2*>(wo + ((lane >> 2) + 8) * HD + d) = make_float2(o[i][2], o[i][3]);
        }
        if ((lane & 3) == 0) {
            xml[(warp * 16 + (lane >> 2)) * 2] = mx0;
            xml[(warp * 16 + (lane >> 2)) * 2 + 1] = l0;
            xml[(warp * 16 + (lane >> 2) + 8) * 2] = mx1;
            xml[(warp * 16 + (lane >> 2) + 8) * 2 + 1] = l1;
        }
        __syncthreads();
        if (kg == 0) {
            float m0 = mx0, m1 = mx1;
            for (int k2 = 1; k2 < a.kg_count; ++k2) {
                const int w2 = k2 * a.rg_count + rg;
                m0 = fmaxf(m0, xml[(w2 * 16 + (lane >> 2)) * 2]);
                m1 = fmaxf(m1, xml[(w2 * 16 + (lane >> 2) + 8) * 2]);
            }
            const float f0 = m0 == -INFINITY ? 0.0f : m0, f1 = m1 == -INFINITY ? 0.0f : m1;
            float sc0 = exp2f(mx0 - f0), sc1 = exp2f(mx1 - f1);
            l0 *= sc0;
            l1 *= sc1;
#pragma unroll
            for (int i = 0; i < HD / 8; ++i) {
                o[i][0] *= sc0;
                o[i][1] *= sc0;
                o[i][2] *= sc1;
                o[i][3] *= sc1;
            }
            for (int k2 = 1; k2 < a.kg_count; ++k2) {
                const int w2 = k2 * a.rg_count + rg;
                const float* po = xo + (size_t)w2 * 16 * HD;
                sc0 = exp2f(xml[(w2 * 16 + (lane >> 2)) * 2] - f0);
                sc1 = exp2f(xml[(w2 * 16 + (lane >> 2) + 8) * 2] - f1);
                l0 += xml[(w2 * 16 + (lane >> 2)) * 2 + 1] * sc0;
                l1 += xml[(w2 * 16 + (lane >> 2) + 8) * 2 + 1] * sc1;
#pragma unroll
                for (int i = 0; i < HD / 8; ++i) {
                    const int d = i * 8 + (lane & 3) * 2;
                    const float2 a0 = *reinterpret_cast<const float2*>(po + (lane >> 2) * HD + d);
                    const float2 a1 = *reinterpret_cast<const float2*>(po + ((lane >> 2) + 8) * HD + d);
                    o[i][0] += a0.x * sc0;
                    o[i][1] += a0.y * sc0;
                    o[i][2] += a1.x * sc1;
                    o[i][3] += a1.y * sc1;
                }
            }
            mx0 = m0;
            mx1 = m1;
        }
    }
    attn_stamp(a, 6);
    // ---- results: final output when the sequence has a single split, else partials + last-CTA combine
    if (kg == 0) {
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int r = hrow ? r1 : r0;
            if (r >= R) continue;
            const int t = r / G, gq = r - t * G;
            const size_t th = (size_t)(q0 + t) * a.nh + g * G + gq;
            const float l = hrow ? l1 : l0;
            if (nsplit_seq == 1) {
                __nv_bfloat16* op = a.out + th * HD;
                const float inv = 1.0f / l;
#pragma unroll
                for (int i = 0; i < HD / 8; ++i) {
                    const int d = i * 8 + (lane & 3) * 2;
                    *reinterpret_cast<__nv_bfloat162*>(op + d) =
                        __floats2bfloat162_rn(o[i][hrow * 2] * inv, o[i][hrow * 2 + 1] * inv);
                }
            } else {
                const size_t idx = th * a.nsplit_max + sp;
                float* op = a.o_part + idx * HD;
#pragma unroll
                for (int i = 0; i < HD / 8; ++i) {
                    const int d = i * 8 + (lane & 3) * 2;
                    *reinterpret_cast<float2*>(op + d) = make_float2(o[i][hrow * 2], o[i][hrow * 2 + 1]);
                }
                if ((lane & 3) == 0) {
                    a.ml_part[idx * 2] = hrow ? mx1 : mx0;
                    a.ml_part[idx * 2 + 1] = l;
                }
            }
        }
    }
    attn_stamp(a, 7);
    if (nsplit_seq == 1) return;
    // the last CTA of this (sequence, kv head) to finish merges the splits (no separate combine launch)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int tk = atomicAdd(&a.tickets[seq * a.nkv + g], 1);
        s_last = (tk == nsplit_seq - 1);
        if (s_last) a.tickets[seq * a.nkv + g] = 0;
    }
    __syncthreads();
    attn_stamp(a, 8);
    if (!s_last) return;
    __threadfence();
    for (int idx = threadIdx.x; idx < R * (HD / 4); idx += blockDim.x) {
        const int r = idx / (HD / 4), d4 = idx - r * (HD / 4);
        const int t = r / G, gq = r - t * G;
        const size_t th = (size_t)(q0 + t) * a.nh + g * G + gq;
        const int qpos = kv_len - qlen + t;
        const int ns = min(nsplit_seq, qpos / a.split_keys + 1);   // splits that contain a visible key
        const size_t base = th * a.nsplit_max;
        float mx = -INFINITY;
        for (int s2 = 0; s2 < ns; ++s2) mx = fmaxf(mx, __ldcg(&a.ml_part[(base + s2) * 2]));
        float l = 0.0f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s2 = 0; s2 < ns; ++s2) {
            const float w = exp2f(__ldcg(&a.ml_part[(base + s2) * 2]) - mx);
            l += __ldcg(&a.ml_part[(base + s2) * 2 + 1]) * w;
            const float4 v = __ldcg(reinterpret_cast<const float4*>(a.o_part + (base + s2) * HD) + d4);
            acc.x += v.x * w;
            acc.y += v.y * w;
            acc.z += v.z * w;
            acc.w += v.w * w;
        }
        const float inv = 1.0f / l;
        __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(a.out + th * HD) + d4 * 2;
        op[0] = __floats2bfloat162_rn(acc.x * inv, acc.y * inv);
        op[1] = __floats2bfloat162_rn(acc.z * inv, acc.w * inv);
    }
    attn_stamp(a, 9);
}

unsigned long long* g_attn_trace = nullptr;
int g_attn_trace_max = 0, g_attn_trace_next = 0;

template <int HD, int KW, int NS>
static int launch_mma(const AttnArgs& a, const CUtensorMap& pf0, const CUtensorMap& pf1, dim3 grid, int threads,
                      size_t smem, cudaStream_t stream) {
    static PerDeviceOnce once;
    if (once.need()) {
        ASD_CUDA(cudaFuncSetAttribute(attn_mma_kernel<HD, KW, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      200 * 1024));
        prefer_max_smem(attn_mma_kernel<HD, KW, NS>);
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = tuning().glue_pdl ? 1 : 0;
    ASD_CUDA(cudaLaunchKernelEx(&cfg, attn_mma_kernel<HD, KW, NS>, a, pf0, pf1));
    return 0;
}

int launch_attention(const AttnLaunch& L, cudaStream_t stream) {
    if (L.M <= 0) return 0;
    const float scale_log2 = 1.4426950408889634f / sqrtf((float)L.hd);
    if (L.hd != 64 && L.hd != 128) return set_error("attention: head_dim must be 64 or 128");
    if (L.impl == 0) {
        attn_simple_kernel<<<dim3(L.M, L.nh), 32, 0, stream>>>(L.q, L.k_cache, L.v_cache, L.positions, L.token_slot,
                                                               L.page_table, L.max_pages, L.out, L.nh, L.nkv, L.hd,
                                                               L.page_size, scale_log2);
        ASD_CUDA(cudaGetLastError());
        count_launch(1);
        return 0;
    }
    const int G = L.nh / L.nkv;
    const int rows = L.max_qlen * G;
    if (L.impl == 2 && L.kv_map != nullptr && L.hd == 128 && rows <= 128 && L.split_keys % 128 == 0)
        return launch_attention_tc(L, stream);
    const int rg = (rows + 15) / 16;
    if (rg > 8) return set_error("attention: q_len * group = %d rows exceeds 128; chunk the query", rows);
    // key groups per row group: keep 4-8 warps per CTA so that one CTA per SM still hides latency
    const int kg = tuning().attn_wide ? (rg <= 2 ? 4 : (rg <= 4 ? 2 : 1)) : (rg == 1 ? 4 : (rg == 2 ? 2 : 1));
    const int warps = rg * kg;
    AttnArgs a;
    a.q = L.q;
    a.k_cache = L.k_cache;
    a.v_cache = L.v_cache;
    a.positions = L.positions;
    a.cu_q = L.cu_q;
    a.seq_slot = L.seq_slot;
    a.page_table = L.page_table;
    a.max_pages = L.max_pages;
    a.nh = L.nh;
    a.nkv = L.nkv;
    a.page_size = L.page_size;
    a.page_shift = -1;
    for (int b = 0; b < 16; ++b)
        if ((1 << b) == L.page_size) a.page_shift = b;
    a.split_keys = L.split_keys;
    a.nsplit_max = L.nsplit_max;
    a.rg_count = rg;
    a.kg_count = kg;
    a.scale_log2 = scale_log2;
    a.o_part = L.o_part;
    a.ml_part = L.ml_part;
    a.tickets = L.tickets;
    a.out = L.out;
    a.trace = nullptr;
    if (g_attn_trace && g_attn_trace_next < g_attn_trace_max)
        a.trace = g_attn_trace + (size_t)(g_attn_trace_next++) * 1024 * 16;
    const dim3 grid(L.nseq, L.nkv, L.nsplit_max);
    if (grid.x * grid.y * grid.z > 1024) a.trace = nullptr;
    const int ns = (int)(grid.x * grid.y * grid.z) <= 148 ? 4 : 2;   // one CTA per SM: deeper K/V ring
    size_t smem = (size_t)rg * 16 * L.hd * 2 + 2 * (size_t)ns * kKeyTile * L.hd * 2;
    const size_t pt_bytes = ((size_t)L.split_keys / L.page_size + 2) * sizeof(int);   // staged page-table slice
    smem += pt_bytes;
    const size_t merge = (size_t)rg * 16 * L.hd * 2 + (size_t)warps * 16 * (L.hd + 2) * 4;
    if (kg > 1 && merge > smem) smem = merge;
    const int kw = kKeyTile / kg, th = warps * 32;
    int rc;
    static const CUtensorMap no_map = {};
    const CUtensorMap* pfm[2] = {&no_map, &no_map};
    for (int i = 0; i < 2; ++i) {
        const bool on = L.pf[i].tmap != nullptr && L.pf[i].kp > 0;
        a.pf_ntiles[i] = L.pf[i].ntiles;
        a.pf_ksplit[i] = L.pf[i].ksplit > 0 ? L.pf[i].ksplit : 1;
        a.pf_kblocks[i] = L.pf[i].kblocks;
        a.pf_kp[i] = on ? L.pf[i].kp : 0;
        if (on) pfm[i] = L.pf[i].tmap;
    }
#define ASD_ATTN(HD_, KW_) (ns == 4 ? launch_mma<HD_, KW_, 4>(a, *pfm[0], *pfm[1], grid, th, smem, stream) \
                                    : launch_mma<HD_, KW_, 2>(a, *pfm[0], *pfm[1], grid, th, smem, stream))
    if (L.hd == 128)
        rc = kw == 64 ? ASD_ATTN(128, 64) : (kw == 32 ? ASD_ATTN(128, 32) : ASD_ATTN(128, 16));
    else
        rc = kw == 64 ? ASD_ATTN(64, 64) : (kw == 32 ? ASD_ATTN(64, 32) : ASD_ATTN(64, 16));
#undef ASD_ATTN
    if (rc) return rc;
    count_launch(1);
    return 0;
}

// see preload_layers()
int preload_attention() {
    cudaFuncAttributes fa;
    ASD_CUDA(cudaFuncGetAttributes(&fa, attn_simple_kernel));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<64, 64, 2>)));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<64, 64, 4>)));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<64, 32, 2>)));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<64, 32, 4>)));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<64, 16, 2>)));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<64, 16, 4>)));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<128, 64, 2>)));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<128, 64, 4>)));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<128, 32, 2>)));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<128, 32, 4>)));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<128, 16, 2>)));
    ASD_CUDA(cudaFuncGetAttributes(&fa, (attn_mma_kernel<128, 16, 4>)));
    return 0;
}

}  // namespace asd

extern "C" __attribute__((visibility("default"))) int asd_debug_attn_trace(unsigned long long* buf, int max_launches) {
    asd::g_attn_trace = buf;
    asd::g_attn_trace_max = buf ? max_launches : 0;
    asd::g_attn_trace_next = 0;
    return 1024 * 16;
}
