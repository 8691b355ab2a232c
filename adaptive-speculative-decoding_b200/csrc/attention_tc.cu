// Paged-KV verify attention on tcgen05 / TMEM (head_dim 128): the Blackwell-native replacement of the mma.sync
// kernel in attention.cu, which is bound by legacy-HMMA issue once the context is long (BASELINE configs[4]:
// 72B verify at a 4096-token prefix spent 28 of 78 ms there at 0.23 of HBM) and by per-tile latency when it is short.
//
// Work item = (sequence, kv head, kv split) as before: the query tile is all new positions x the GQA group
// (rows = q_len * G <= 128, padded to the 128 TMEM lanes), so each K/V byte is read once per sequence.
// Per 128-key tile:
//   S   = Q K^T     tcgen05.mma M=128 N=128 (8 k-steps over head_dim), both operands K-major SW128 in shared
//                   memory; K pages arrive by TMA straight from the paged pool (one 2-D box of one page x 64
//                   head-dim elements per page and half, gathered through the page table);
//   softmax         warps 0-3, thread = query row = TMEM lane: row maximum and exp2 over two tcgen05.ld passes
//                   (straight-line, select-based masking, four independent accumulators: the first version's
//                   per-pair branches serialised one MUFU latency after the other - 5.2 us per tile, measured with
//                   tools/trace_attn.py); the bf16 probabilities go straight back into TMEM (tcgen05.st over the
//                   first 64 columns of the S buffer they came from), never through shared memory;
//   O_t = P V       tcgen05.mma with A = P from TMEM and V as an MN-major B operand (the pool stores
//                   [position][head_dim], so keys are the contraction dimension: no transpose, no ldmatrix.trans);
//   o   = o * corr + O_t   warps 4-7 (thread = row), in registers; the rescale factor of every tile travels from
//                   the softmax warps through a 4-deep shared-memory ring, so softmax(j+1) and fold(j) overlap.
// S/P and O_t are double-buffered in TMEM (4 x 128 columns): QK^T of tile j+1 runs while the softmax of tile j is
// in flight, P V of tile j while the softmax of tile j+1 runs (P lives in the other S buffer: the second version
// kept P in ONE shared-memory buffer and its softmax warps spent a quarter of their time waiting for P V to release
// it, profiles/ncu_attn_tc_r02.txt).  Warp 8 streams K, warp 9 streams V (2-deep rings), one thread of warp 10
// issues all MMAs.  Warps whose 32 rows are all padding sit the loop out (the barrier counts are the number of
// active warps).  Long contexts are split over the KV length (flash decoding) with the same ticket merge as
// attention.cu.
//
// The reference has no attention code (vLLM does it: /root/reference/src/serving/real_model_pipeline.py:98-108);
// semantics are HF Qwen2 GQA attention with softmax scale 1/sqrt(head_dim), causal among the new positions.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>

#include "asd_internal.h"
#include "layers.h"
#include "ptx.cuh"

namespace asd {

extern unsigned long long* g_attn_trace;
extern int g_attn_trace_max, g_attn_trace_next;

constexpr int kAtThreads = 352;       // warps 0-3 softmax, 4-7 fold (thread = query row), 8 K loader, 9 V loader, 10 MMA
constexpr int kAtTile = 128;          // keys per tile
constexpr int kAtHalf = 128 * 128;    // bytes of one 64-element half of a [128 rows][128 dims] bf16 operand
constexpr int kAtOp = 2 * kAtHalf;    // 32 KB: Q, one K stage, one V stage

struct AttnTcArgs {
    const __nv_bfloat16* q;        // [M, nh, 128]
    const int* positions;          // [M]
    const int* cu_q;               // [nseq + 1]
    const int* seq_slot;           // [nseq]
    const int* page_table;
    int max_pages, nh, nkv, split_keys, nsplit_max;
    int page_shift;                // log2(positions per page): 4..7 (one TMA box = one page x 64 head-dim elements)
    int dbg;                       // diagnostics: 1 = skip the softmax arithmetic, 2 = also skip the fold
    long long k_row0, v_row0;      // first pool row (position-major rows of 128 dims) of this layer's K / V
    float scale_log2;
    float* o_part;                 // [M, nh, nsplit_max, 128]
    float* ml_part;                // [M, nh, nsplit_max, 2]
    int* tickets;
    __nv_bfloat16* out;            // [M, nh, 128]
    unsigned long long* trace;     // diagnostics: 16 globaltimer stamps per CTA (nullptr = off)
};
__device__ __forceinline__ void at_stamp(const AttnTcArgs& a, int slot) {
    if (a.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        a.trace[(size_t)cta * 16 + slot] = t;
    }
}

// K-major SW128 operand descriptor at byte address `addr` (1024-aligned atom rows): same as gemm.cu
__device__ __forceinline__ uint64_t at_desc_k(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// MN-major SW128 operand ([K rows of 128 B = 64 MN elements], 8-row atoms 1024 B apart along K = SBO, the next 64 MN
// elements one operand half further = LBO)
__device__ __forceinline__ uint64_t at_desc_mn(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(kAtHalf >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ float at_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void at_tma_box(void* dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A = 128 lanes x (K = 16 bf16 = 8 packed 32-bit columns per instruction)
__device__ __forceinline__ void at_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns <- 16 registers per thread (thread = lane)
__device__ __forceinline__ void at_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void at_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One 64-key half of the row maximum: element i of the half is visible iff i <= lim (lim >= 63: all of them).
template <bool MASKED>
__device__ __forceinline__ void at_max64(const uint32_t (&v0)[32], const uint32_t (&v1)[32], int lim, float (&m)[4]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float x0 = __uint_as_float(v0[i]), x1 = __uint_as_float(v1[i]);
        if (MASKED) {
            x0 = i <= lim ? x0 : -INFINITY;
            x1 = i + 32 <= lim ? x1 : -INFINITY;
        }
        m[i & 3] = fmaxf(m[i & 3], fmaxf(x0, x1));
    }
}
// exp2(s * scale - ref) of 32 scores -> 16 packed bf16 pairs; the row sum takes the unrounded values (as attention.cu)
template <bool MASKED>
__device__ __forceinline__ void at_exp32(const uint32_t (&v)[32], int lim, float scale, float nref, uint32_t (&pk)[16],
                                         float (&ps)[4]) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        float p0 = at_ex2(fmaf(__uint_as_float(v[i]), scale, nref));
        float p1 = at_ex2(fmaf(__uint_as_float(v[i + 1]), scale, nref));
        if (MASKED) {
            p0 = i <= lim ? p0 : 0.0f;
            p1 = i + 1 <= lim ? p1 : 0.0f;
        }
        ps[(i >> 1) & 3] += p0 + p1;
        const __nv_bfloat162 b = __floats2bfloat162_rn(p0, p1);
        pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&b);
    }
}

__global__ void __launch_bounds__(kAtThreads, 1) attn_tc_kernel(const __grid_constant__ CUtensorMap kv_map, const AttnTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kAtOp;              // [2 stages]
    uint8_t* sV = sK + 2 * kAtOp;          // [2 stages]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kAtOp);
    uint64_t *k_full = bars, *k_empty = bars + 2, *v_full = bars + 4, *v_empty = bars + 6, *s_full = bars + 8,
             *s_empty = bars + 10, *o_full = bars + 12, *o_empty = bars + 14, *p_full = bars + 16, *q_full = bars + 17,
             *c_full = bars + 18;   // c_full[4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
    int* s_last = reinterpret_cast<int*>(tmem_slot + 1);
    float* corr_s = reinterpret_cast<float*>(tmem_slot + 4);   // [4][128] rescale factor of tile j (ring j & 3)
    float* fin_s = corr_s + 4 * 128;                           // [2][128] final row maximum / row sum
    int* s_pt = reinterpret_cast<int*>(fin_s + 2 * 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seq = blockIdx.x, g = blockIdx.y, sp = blockIdx.z;
    const int G = a.nh / a.nkv;
    const int q0 = a.cu_q[seq], qlen = a.cu_q[seq + 1] - q0;
    if (qlen <= 0) return;
    const int R = qlen * G;
    const int nact = (R + 31) >> 5;          // softmax / fold warps that own at least one real row
    const int kv_len = a.positions[q0 + qlen - 1] + 1;
    const int kbeg = sp * a.split_keys;
    if (kbeg >= kv_len) return;
    const int kend = min(kv_len, kbeg + a.split_keys);
    const int nsplit_seq = (kv_len + a.split_keys - 1) / a.split_keys;
    const int old_keys = kv_len - qlen;
    const int nt = (kend - kbeg + kAtTile - 1) / kAtTile;
    const int psh = a.page_shift;
    const int page0 = kbeg >> psh, npages = ((kend - 1) >> psh) - page0 + 1;
    if (threadIdx.x == 0) at_stamp(a, 0);

    {   // this CTA's slice of the page table (static across forwards: safe before the PDL wait)
        const int* pt = a.page_table + (size_t)a.seq_slot[seq] * a.max_pages;
        for (int i = threadIdx.x; i < npages; i += kAtThreads) s_pt[i] = pt[page0 + i];
    }
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&kv_map);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 1);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], nact);
        }
        mbar_init(p_full, nact);
        mbar_init(q_full, 4);
        for (int i = 0; i < 4; ++i) mbar_init(&c_full[i], nact);
        fence_mbar_init();
    }
    if (warp == 10) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8 || warp == 9) {
        // ------------------------------------------------------------------ K loader (warp 8) / V loader (warp 9)
        if (lane == 0) {
            const int kv = warp - 8;
            uint64_t* full = kv ? v_full : k_full;
            uint64_t* empty = kv ? v_empty : k_empty;
            uint8_t* ring = kv ? sV : sK;
            const long long row0 = kv ? a.v_row0 : a.k_row0;
            bool waited = false;
            for (int j = 0; j < nt; ++j) {
                const int st = j & 1, use = j >> 1;
                const int key0 = kbeg + j * kAtTile;
                if (!waited && key0 + kAtTile > old_keys) {   // this tile holds keys the upstream QKV GEMM writes
                    grid_dep_wait();
                    waited = true;
                }
                uint8_t* dst = ring + st * kAtOp;
                if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
                mbar_expect_tx(&full[st], kAtOp);
#pragma unroll 1
                for (int p = 0; p < (kAtTile >> psh); ++p) {
                    int pi = (key0 >> psh) + p - page0;
                    if (pi >= npages) pi = npages - 1;       // past the end: any page of this sequence (masked)
                    const int row = (int)(row0 + (((long long)s_pt[pi] * a.nkv + g) << psh));
                    at_tma_box(dst + ((p * 128) << psh), &kv_map, 0, row, &full[st]);
                    at_tma_box(dst + kAtHalf + ((p * 128) << psh), &kv_map, 64, row, &full[st]);
                }
                if (j == 0) at_stamp(a, 11 + kv);
            }
            if (!waited) grid_dep_wait();
        }
        __syncwarp();
    } else if (warp == 10) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc_qk = umma_idesc_bf16(128, 128);
            const uint32_t idesc_pv = umma_idesc_bf16(128, 128) | (1u << 16);     // B (= V) is MN-major
            const uint32_t uQ = smem_u32(sQ), uK = smem_u32(sK), uV = smem_u32(sV);
            auto issue_qk = [&](int j) {
                const int st = j & 1, use = j >> 1;
                if (use > 0) mbar_wait(&s_empty[st], (use - 1) & 1);   // P V of tile j-2 has consumed P in this buffer
                mbar_wait(&k_full[st], use & 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(st * 128);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t off = (uint32_t)((k >> 2) * kAtHalf + (k & 3) * 32);
                    umma_f16(d, at_desc_k(uQ + off), at_desc_k(uK + st * kAtOp + off), idesc_qk, k != 0);
                }
                umma_commit(&s_full[st]);
                umma_commit(&k_empty[st]);
                if (j == 0) at_stamp(a, 13);
            };
            mbar_wait(q_full, 0);
            issue_qk(0);
            for (int j = 0; j < nt; ++j) {
                const int st = j & 1, use = j >> 1;
                if (j + 1 < nt) issue_qk(j + 1);
                mbar_wait(p_full, j & 1);
                mbar_wait(&v_full[st], use & 1);
                if (use > 0) mbar_wait(&o_empty[st], (use - 1) & 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(256 + st * 128);
                const uint32_t p = tmem_base + (uint32_t)(st * 128);     // bf16 P: 64 packed columns over S
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    at_umma_ts(d, p + (uint32_t)(k * 8), at_desc_mn(uV + st * kAtOp + k * 2048), idesc_pv, k != 0);
                umma_commit(&o_full[st]);
                umma_commit(&v_empty[st]);
                umma_commit(&s_empty[st]);
                if (j == 0) at_stamp(a, 14);
            }
        }
        __syncwarp();
    } else if (warp < 4) {
        // ------------------------------------------------------------------ softmax (thread = row)
        const int r = threadIdx.x;               // 0..127
        grid_dep_wait();                         // q comes from the QKV GEMM launched just before
        grid_dep_launch();
        {   // stage Q: row r = t * G + gq -> token q0 + t, head g * G + gq; zero rows past R
            const bool ok = r < R;
            const int t = ok ? r / G : 0, gq = ok ? r - t * G : 0;
            const uint4* src = reinterpret_cast<const uint4*>(a.q + ((size_t)(q0 + t) * a.nh + g * G + gq) * 128);
            uint4 v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = ok ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int c = 0; c < 16; ++c)
                *reinterpret_cast<uint4*>(sQ + (c >> 3) * kAtHalf + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = v[c];
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_full);
            if (r == 0) at_stamp(a, 1);
        }
        if (warp < nact) {
            const int qpos = r < R ? kv_len - qlen + r / G : -1;
            const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
            const float scale = a.scale_log2;
            float mx = -INFINITY, l = 0.0f;

            auto tile = [&](int j, auto masked_tag) {
                constexpr bool MASKED = decltype(masked_tag)::value;
                const int st = j & 1, use = j >> 1;
                const int lim = min(kend - 1, qpos) - (kbeg + j * kAtTile);   // element i of the tile visible iff i <= lim
                const uint32_t tS = lane_base + (uint32_t)(st * 128);
                mbar_wait(&s_full[st], use & 1);
                tc_fence_after();
                if (r == 0 && (j < 2 || j == nt - 1)) at_stamp(a, j == 0 ? 2 : (j == 1 ? 4 : 7));
                if (a.dbg >= 1) {
                    corr_s[(j & 3) * 128 + r] = 1.0f;
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(p_full);
                        mbar_arrive(&c_full[j & 3]);
                    }
                    return;
                }
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                // boundary tile: a 64-key half none of this warp's rows can see is skipped (its probabilities are zeros) -
                // with a short prefix the last tile holds only the step's own few keys
                const int wlim = MASKED ? __reduce_max_sync(0xffffffffu, lim) : kAtTile;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (MASKED && wlim < h * 64) continue;
                    uint32_t v0[32], v1[32];
                    tmem_ld_32x32(tS + (uint32_t)(h * 64), v0);
                    tmem_ld_32x32(tS + (uint32_t)(h * 64 + 32), v1);
                    tmem_ld_wait();
                    at_max64<MASKED>(v0, v1, lim - h * 64, m4);
                }
                const float tmax = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale;   // scale > 0
                const float mn = fmaxf(mx, tmax);
                const float ref = mn == -INFINITY ? 0.0f : mn;
                const float corr = at_ex2(mx - ref);      // mx = -inf -> 0
                corr_s[(j & 3) * 128 + r] = corr;
                if (r == 0 && j == 1) at_stamp(a, 15);
                float ps[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                // second pass: the packed probabilities of columns [32c, 32c + 32) overwrite columns [16c, 16c + 16) of
                // the same buffer, which this thread has already read (its own lane only)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v0[32], v1[32], pk[16];
                    if (MASKED && wlim < h * 64) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = 0u;
                        at_tmem_st16(tS + (uint32_t)(h * 32), pk);
                        at_tmem_st16(tS + (uint32_t)(h * 32 + 16), pk);
                        continue;
                    }
                    tmem_ld_32x32(tS + (uint32_t)(h * 64), v0);
                    tmem_ld_32x32(tS + (uint32_t)(h * 64 + 32), v1);
                    tmem_ld_wait();
                    at_exp32<MASKED>(v0, lim - h * 64, scale, -ref, pk, ps);
                    at_tmem_st16(tS + (uint32_t)(h * 32), pk);
                    at_exp32<MASKED>(v1, lim - h * 64 - 32, scale, -ref, pk, ps);
                    at_tmem_st16(tS + (uint32_t)(h * 32 + 16), pk);
                }
                at_tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(p_full);
                    mbar_arrive(&c_full[j & 3]);
                }
                if (r == 0 && j < 2) at_stamp(a, j == 0 ? 3 : 5);
                l = l * corr + ((ps[0] + ps[1]) + (ps[2] + ps[3]));
                mx = mn;
            };
            for (int j = 0; j < nt; ++j) {
                const int key0 = kbeg + j * kAtTile;
                // keys of a tile that ends at or before the first new position are visible to every real row (padding
                // rows compute on zero queries: finite values nobody reads)
                const bool interior = key0 + kAtTile <= old_keys + 1 && key0 + kAtTile <= kend;
                if (interior)
                    tile(j, std::false_type{});
                else
                    tile(j, std::true_type{});
            }
            fin_s[r] = mx;
            fin_s[128 + r] = l;
            __syncwarp();
            if (lane == 0) mbar_arrive(&c_full[nt & 3]);
            if (r == 0) at_stamp(a, 8);
        }
    } else {
        // ------------------------------------------------------------------ fold: o = o * corr + O_tile (thread = row)
        const int r = threadIdx.x - 128, w4 = warp - 4;
        grid_dep_wait();
        if (w4 < nact) {
            const uint32_t lane_base = tmem_base + ((uint32_t)(w4 * 32) << 16);
            float o[128];
#pragma unroll
            for (int i = 0; i < 128; ++i) o[i] = 0.0f;
            for (int j = 0; j < nt; ++j) {
                const int st = j & 1, use = j >> 1;
                mbar_wait(&c_full[j & 3], (j >> 2) & 1);
                const float corr = corr_s[(j & 3) * 128 + r];
                mbar_wait(&o_full[st], use & 1);
                tc_fence_after();
                if (a.dbg < 2) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32(lane_base + (uint32_t)(256 + st * 128 + c * 32), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], corr, __uint_as_float(v[i]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_empty[st]);
                if (r == 0 && j == 0) at_stamp(a, 6);
            }
            mbar_wait(&c_full[nt & 3], (nt >> 2) & 1);
            const float mx = fin_s[r], l = fin_s[128 + r];
            if (r == 0) at_stamp(a, 9);
            // ---- results: final output when the sequence has a single split, else partials + last-CTA combine
            if (r < R) {
                const int t = r / G, gq = r - t * G;
                const size_t th = (size_t)(q0 + t) * a.nh + g * G + gq;
                if (nsplit_seq == 1) {
                    const float inv = 1.0f / l;
                    uint4* op = reinterpret_cast<uint4*>(a.out + th * 128);
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        __nv_bfloat162 b[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) b[i] = __floats2bfloat162_rn(o[c * 8 + 2 * i] * inv, o[c * 8 + 2 * i + 1] * inv);
                        op[c] = *reinterpret_cast<const uint4*>(b);
                    }
                } else {
                    const size_t idx = th * a.nsplit_max + sp;
                    float4* op = reinterpret_cast<float4*>(a.o_part + idx * 128);
#pragma unroll
                    for (int c = 0; c < 32; ++c) op[c] = make_float4(o[c * 4], o[c * 4 + 1], o[c * 4 + 2], o[c * 4 + 3]);
                    a.ml_part[idx * 2] = mx;
                    a.ml_part[idx * 2 + 1] = l;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 10) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 0) at_stamp(a, 10);
    if (nsplit_seq == 1) return;
    // the last CTA of this (sequence, kv head) to finish merges the splits (no separate combine launch)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int tk = atomicAdd(&a.tickets[seq * a.nkv + g], 1);
        *s_last = (tk == nsplit_seq - 1);
        if (*s_last) a.tickets[seq * a.nkv + g] = 0;
    }
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    for (int idx = threadIdx.x; idx < R * 32; idx += kAtThreads) {
        const int r = idx >> 5, d4 = idx & 31;
        const int t = r / G, gq = r - t * G;
        const size_t th = (size_t)(q0 + t) * a.nh + g * G + gq;
        const int qpos = kv_len - qlen + t;
        const int ns = min(nsplit_seq, qpos / a.split_keys + 1);   // splits that contain a visible key
        const size_t base = th * a.nsplit_max;
        float mxs = -INFINITY;
        for (int s2 = 0; s2 < ns; ++s2) mxs = fmaxf(mxs, __ldcg(&a.ml_part[(base + s2) * 2]));
        float l = 0.0f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s2 = 0; s2 < ns; ++s2) {
            const float w = exp2f(__ldcg(&a.ml_part[(base + s2) * 2]) - mxs);
            l += __ldcg(&a.ml_part[(base + s2) * 2 + 1]) * w;
            const float4 v = __ldcg(reinterpret_cast<const float4*>(a.o_part + (base + s2) * 128) + d4);
            acc.x += v.x * w;
            acc.y += v.y * w;
            acc.z += v.z * w;
            acc.w += v.w * w;
        }
        const float inv = 1.0f / l;
        __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(a.out + th * 128) + d4 * 2;
        op[0] = __floats2bfloat162_rn(acc.x * inv, acc.y * inv);
        op[1] = __floats2bfloat162_rn(acc.z * inv, acc.w * inv);
    }
}

// ------------------------------------------------------------------------------------------- host
int launch_attention_tc(const AttnLaunch& L, cudaStream_t stream) {
    static PerDeviceOnce once;
    const size_t smem = 1024 + 5 * (size_t)kAtOp + 22 * 8 + 16 + 6 * 128 * 4 +
                        ((size_t)L.split_keys / L.page_size + 2) * 4 + 64;
    if (once.need()) {
        ASD_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        prefer_max_smem(attn_tc_kernel);
    }
    if (smem > 227 * 1024) return set_error("attention_tc: split of %d keys does not fit the page-table slice", L.split_keys);
    AttnTcArgs a;
    a.q = L.q;
    a.positions = L.positions;
    a.cu_q = L.cu_q;
    a.seq_slot = L.seq_slot;
    a.page_table = L.page_table;
    a.max_pages = L.max_pages;
    a.nh = L.nh;
    a.nkv = L.nkv;
    a.split_keys = L.split_keys;
    a.nsplit_max = L.nsplit_max;
    a.page_shift = 0;
    while ((1 << a.page_shift) < L.page_size) ++a.page_shift;
    a.dbg = tuning().attn_dbg;
    a.k_row0 = L.k_row0;
    a.v_row0 = L.v_row0;
    a.scale_log2 = 1.4426950408889634f / sqrtf((float)L.hd);
    a.o_part = L.o_part;
    a.ml_part = L.ml_part;
    a.tickets = L.tickets;
    a.out = L.out;
    a.trace = nullptr;
    if (g_attn_trace && g_attn_trace_next < g_attn_trace_max && L.nseq * L.nkv * L.nsplit_max <= 1024)
        a.trace = g_attn_trace + (size_t)(g_attn_trace_next++) * 1024 * 16;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(L.nseq, L.nkv, L.nsplit_max);
    cfg.blockDim = dim3(kAtThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = tuning().glue_pdl ? 1 : 0;
    ASD_CUDA(cudaLaunchKernelEx(&cfg, attn_tc_kernel, *L.kv_map, a));
    count_launch(1);
    return 0;
}

// Force the (lazily loaded) kernels of this file into the context now: a first launch that loads a kernel may need a
// context synchronisation, which deadlocks when another rank of the same process is spinning for this rank's launch.
int preload_attention_tc() {
    cudaFuncAttributes fa;
    ASD_CUDA(cudaFuncGetAttributes(&fa, attn_tc_kernel));
    return 0;
}

}  // namespace asd
