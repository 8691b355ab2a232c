// Paged-KV verify attention on tcgen05 / TMEM (head_dim 128): the Blackwell-native replacement of the mma.sync
// kernel in attention.cu, which is bound by legacy-HMMA throughput once the context is long (BASELINE configs[4]:
// 72B verify at a 4096-token prefix spent 28 of 80 ms there at 0.23 of HBM).
//
// Work item = (sequence, kv head, kv split) as before: the query tile is all new positions x the GQA group
// (rows = q_len * G <= 128, padded to the 128 TMEM lanes), so each K/V byte is read once per sequence.
// Per 128-key tile:
//   S   = Q K^T     tcgen05.mma M=128 N=128 (8 k-steps over head_dim), both operands K-major SW128 in shared
//                   memory; K pages arrive by TMA straight from the paged pool (one 2-D box of 16 positions x 64
//                   head-dim elements per page and half, gathered through the page table);
//   softmax         4 warps, thread = query row = TMEM lane: two tcgen05.ld passes over S (row maximum, then
//                   exp2 / row sum / bf16 P written into shared memory in the K-major SW128 layout);
//   O_t = P V       tcgen05.mma with V as an MN-major B operand (the pool stores [position][head_dim], so keys
//                   are the contraction dimension: no transpose, no ldmatrix.trans);
//   o   = o * corr + O_t   in registers (a fresh accumulator per tile: nothing in TMEM is ever rescaled).
// S and O_t are double-buffered in TMEM (4 x 128 columns), so QK^T of tile j+1 runs while the softmax of tile j
// is in flight and P V of tile j while tile j-1 is folded into the registers.  One warp streams K/V (2-deep
// rings, K and V tracked separately), one thread issues all MMAs.  Long contexts are split over the KV length
// (flash decoding) with the same ticket merge as attention.cu.
//
// The reference has no attention code (vLLM does it: /root/reference/src/serving/real_model_pipeline.py:98-108);
// semantics are HF Qwen2 GQA attention with softmax scale 1/sqrt(head_dim), causal among the new positions.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "asd_internal.h"
#include "layers.h"
#include "ptx.cuh"

namespace asd {

constexpr int kAtThreads = 192;       // warps 0..3 softmax (thread = query row), warp 4 K/V loader, warp 5 MMA
constexpr int kAtTile = 128;          // keys per tile
constexpr int kAtHalf = 128 * 128;    // bytes of one 64-element half of a [128 rows][128 dims] bf16 operand
constexpr int kAtOp = 2 * kAtHalf;    // 32 KB: Q, P, one K stage, one V stage

struct AttnTcArgs {
    const __nv_bfloat16* q;        // [M, nh, 128]
    const int* positions;          // [M]
    const int* cu_q;               // [nseq + 1]
    const int* seq_slot;           // [nseq]
    const int* page_table;
    int max_pages, nh, nkv, split_keys, nsplit_max;
    long long k_row0, v_row0;      // first pool row (position-major rows of 128 dims) of this layer's K / V
    float scale_log2;
    float* o_part;                 // [M, nh, nsplit_max, 128]
    float* ml_part;                // [M, nh, nsplit_max, 2]
    int* tickets;
    __nv_bfloat16* out;            // [M, nh, 128]
};

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
__device__ __forceinline__ void at_tma_box(void* dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__global__ void __launch_bounds__(kAtThreads, 1) attn_tc_kernel(const __grid_constant__ CUtensorMap kv_map, const AttnTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sP = sQ + kAtOp;
    uint8_t* sK = sP + kAtOp;              // [2 stages]
    uint8_t* sV = sK + 2 * kAtOp;          // [2 stages]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kAtOp);
    uint64_t *k_full = bars, *k_empty = bars + 2, *v_full = bars + 4, *v_empty = bars + 6, *s_full = bars + 8,
             *s_empty = bars + 10, *o_full = bars + 12, *o_empty = bars + 14, *p_full = bars + 16, *p_empty = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    int* s_last = reinterpret_cast<int*>(tmem_slot + 1);
    int* s_pt = reinterpret_cast<int*>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seq = blockIdx.x, g = blockIdx.y, sp = blockIdx.z;
    const int G = a.nh / a.nkv;
    const int q0 = a.cu_q[seq], qlen = a.cu_q[seq + 1] - q0;
    if (qlen <= 0) return;
    const int R = qlen * G;
    const int kv_len = a.positions[q0 + qlen - 1] + 1;
    const int kbeg = sp * a.split_keys;
    if (kbeg >= kv_len) return;
    const int kend = min(kv_len, kbeg + a.split_keys);
    const int nsplit_seq = (kv_len + a.split_keys - 1) / a.split_keys;
    const int old_keys = kv_len - qlen;
    const int nt = (kend - kbeg + kAtTile - 1) / kAtTile;
    const int page0 = kbeg >> 4, npages = ((kend - 1) >> 4) - page0 + 1;

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
            mbar_init(&s_empty[i], 4);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], 4);
        }
        mbar_init(p_full, 4);
        mbar_init(p_empty, 1);
        fence_mbar_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ------------------------------------------------------------------ K/V loader
        if (lane == 0) {
            bool waited = false;
            for (int j = 0; j < nt; ++j) {
                const int st = j & 1, use = j >> 1;
                const int key0 = kbeg + j * kAtTile;
                if (!waited && key0 + kAtTile > old_keys) {   // this tile holds keys the upstream QKV GEMM writes
                    grid_dep_wait();
                    waited = true;
                }
                for (int kv = 0; kv < 2; ++kv) {
                    uint64_t* full = kv ? &v_full[st] : &k_full[st];
                    uint64_t* empty = kv ? &v_empty[st] : &k_empty[st];
                    uint8_t* dst = (kv ? sV : sK) + st * kAtOp;
                    const long long row0 = kv ? a.v_row0 : a.k_row0;
                    if (use > 0) mbar_wait(empty, (use - 1) & 1);
                    mbar_expect_tx(full, kAtOp);
#pragma unroll 1
                    for (int p = 0; p < 8; ++p) {
                        int pi = (key0 >> 4) + p - page0;
                        if (pi >= npages) pi = npages - 1;       // past the end: any page of this sequence (masked)
                        const int row = (int)(row0 + ((long long)s_pt[pi] * a.nkv + g) * 16);
                        at_tma_box(dst + p * 2048, &kv_map, 0, row, full);
                        at_tma_box(dst + kAtHalf + p * 2048, &kv_map, 64, row, full);
                    }
                }
            }
            if (!waited) grid_dep_wait();
        }
        __syncwarp();
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc_qk = umma_idesc_bf16(128, 128);
            const uint32_t idesc_pv = umma_idesc_bf16(128, 128) | (1u << 16);     // B (= V) is MN-major
            const uint32_t uQ = smem_u32(sQ), uP = smem_u32(sP), uK = smem_u32(sK), uV = smem_u32(sV);
            auto issue_qk = [&](int j) {
                const int st = j & 1, use = j >> 1;
                if (use > 0) mbar_wait(&s_empty[st], (use - 1) & 1);
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
            };
            mbar_wait(p_full, 0);      // phase 0 of p_full doubles as "Q is staged" (see the softmax warps)
            issue_qk(0);
            for (int j = 0; j < nt; ++j) {
                const int st = j & 1, use = j >> 1;
                if (j + 1 < nt) issue_qk(j + 1);
                mbar_wait(p_full, (j + 1) & 1);
                mbar_wait(&v_full[st], use & 1);
                if (use > 0) mbar_wait(&o_empty[st], (use - 1) & 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(256 + st * 128);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t offp = (uint32_t)((k >> 2) * kAtHalf + (k & 3) * 32);
                    umma_f16(d, at_desc_k(uP + offp), at_desc_mn(uV + st * kAtOp + k * 2048), idesc_pv, k != 0);
                }
                umma_commit(&o_full[st]);
                umma_commit(&v_empty[st]);
                umma_commit(p_empty);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ softmax + accumulation (thread = row)
        const int r = threadIdx.x;               // 0..127
        grid_dep_wait();                         // q comes from the QKV GEMM launched just before
        grid_dep_launch();
        {   // stage Q: row r = t * G + gq -> token q0 + t, head g * G + gq; zero rows past R
            const bool ok = r < R;
            const int t = ok ? r / G : 0, gq = ok ? r - t * G : 0;
            const uint4* src = reinterpret_cast<const uint4*>(a.q + ((size_t)(q0 + t) * a.nh + g * G + gq) * 128);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const uint4 v = ok ? src[c] : make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(sQ + (c >> 3) * kAtHalf + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = v;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);      // phase 0: Q staged
        }
        const int qpos = r < R ? kv_len - qlen + r / G : -1;
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        float o[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) o[i] = 0.0f;
        float mx = -INFINITY, l = 0.0f, corr_pending = 1.0f;

        auto fold = [&](int j) {     // o = o * corr_j + O_tile(j)
            const int st = j & 1, use = j >> 1;
            mbar_wait(&o_full[st], use & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(lane_base + (uint32_t)(256 + st * 128 + c * 32), v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * corr_pending + __uint_as_float(v[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[st]);
        };

        for (int j = 0; j < nt; ++j) {
            const int st = j & 1, use = j >> 1;
            const int key0 = kbeg + j * kAtTile;
            // keys of a tile that ends at or before the first new position are visible to every row
            const bool interior = key0 + kAtTile <= old_keys + 1 && key0 + kAtTile <= kend;
            mbar_wait(&s_full[st], use & 1);
            tc_fence_after();
            float tmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(lane_base + (uint32_t)(st * 128 + c * 32), v);
                tmem_ld_wait();
                if (interior) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, __uint_as_float(v[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int key = key0 + c * 32 + i;
                        if (key < kend && key <= qpos) tmax = fmaxf(tmax, __uint_as_float(v[i]));
                    }
                }
            }
            if (r >= R) tmax = -INFINITY;
            tmax *= a.scale_log2;                    // scale_log2 > 0: max commutes with the scaling
            const float mn = fmaxf(mx, tmax);
            const float ref = mn == -INFINITY ? 0.0f : mn;
            const float corr = exp2f(mx - ref);      // mx = -inf -> 0
            if (j > 0) mbar_wait(p_empty, (j - 1) & 1);   // P V of the previous tile has consumed the P buffer
            float psum = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(lane_base + (uint32_t)(st * 128 + c * 32), v);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0, p1;
                    if (interior) {
                        p0 = exp2f(__uint_as_float(v[i]) * a.scale_log2 - ref);
                        p1 = exp2f(__uint_as_float(v[i + 1]) * a.scale_log2 - ref);
                    } else {
                        const int key = key0 + c * 32 + i;
                        p0 = (key < kend && key <= qpos) ? exp2f(__uint_as_float(v[i]) * a.scale_log2 - ref) : 0.0f;
                        p1 = (key + 1 < kend && key + 1 <= qpos) ? exp2f(__uint_as_float(v[i + 1]) * a.scale_log2 - ref) : 0.0f;
                    }
                    if (r >= R) p0 = p1 = 0.0f;
                    const __nv_bfloat162 b = __floats2bfloat162_rn(p0, p1);
                    // the row sum uses the rounded values that the tensor core will multiply with V
                    psum += __bfloat162float(b.x) + __bfloat162float(b.y);
                    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&b);
                }
                // 32 keys = 64 bytes = 4 chunks of 16 bytes: keys c*32 .. c*32+31 -> half c >> 1, chunks (c & 1) * 4 ..
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int ch = (c & 1) * 4 + q4;
                    *reinterpret_cast<uint4*>(sP + (c >> 1) * kAtHalf + r * 128 + ((ch ^ (r & 7)) << 4)) =
                        make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_empty[st]);
                mbar_arrive(p_full);
            }
            l = l * corr + psum;
            mx = mn;
            if (j > 0) fold(j - 1);      // with the correction of tile j-1 (computed one iteration ago)
            corr_pending = corr;
        }
        fold(nt - 1);

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
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
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
    const size_t smem = 1024 + 6 * (size_t)kAtOp + 18 * 8 + 16 + ((size_t)L.split_keys / 16 + 2) * 4 + 64;
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
    a.k_row0 = L.k_row0;
    a.v_row0 = L.v_row0;
    a.scale_log2 = 1.4426950408889634f / sqrtf((float)L.hd);
    a.o_part = L.o_part;
    a.ml_part = L.ml_part;
    a.tickets = L.tickets;
    a.out = L.out;
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

}  // namespace asd
