// Paged-KV attention for the verify step: every (k+1) new position of a sequence is scored against
// the paged prefix in ONE pass (causal among the new positions), split over the KV length
// (flash-decoding) so that B x n_kv x splits CTAs keep HBM busy; a combine kernel merges splits.
//
//   * work item = (sequence, kv head, kv split).  The query tile is all new tokens x the GQA group
//     of that kv head (rows = q_len * G, e.g. 6 x 5 = 30 for Qwen2.5-32B with k = 5), so each K/V
//     byte is read from HBM once per sequence, not once per query head;
//   * QK^T and PV run on tensor cores (mma.sync m16n8k16 bf16 -> fp32; 30..72 query rows cannot fill a
//     128-row tcgen05 tile and the kernel is bound by the KV stream, not by math), operands staged
//     with cp.async into XOR-swizzled shared memory and read with ldmatrix;
//   * K/V were appended in place by qkv_rope_kernel before this kernel runs; rejected speculative
//     positions are simply overwritten by the next step (rollback = not advancing the length).
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

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

struct AttnArgs {
    const __nv_bfloat16* q;        // [M, nh, hd]
    const __nv_bfloat16* k_cache;  // [pages, nkv, page, hd]
    const __nv_bfloat16* v_cache;
    const int* positions;          // [M]
    const int* cu_q;               // [nseq + 1] token ranges
    const int* seq_slot;           // [nseq]
    const int* page_table;
    int max_pages, nh, nkv, page_size, split_keys, nsplit_max;
    float scale_log2;
    float* o_part;                 // [M, nh, nsplit_max, hd]
    float* ml_part;                // [M, nh, nsplit_max, 2]
};

// HD = head_dim (64 or 128).  blockDim = 32 * warps, each warp owns 16 query rows.
template <int HD>
__global__ void __launch_bounds__(256) attn_mma_kernel(const AttnArgs a) {
    constexpr int CH = HD / 8;          // 16-byte chunks per row
    constexpr int ROWB = HD * 2;        // bytes per row
    constexpr int KB = HD / 16;         // k-blocks over head_dim
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sQ = smem_raw;                                   // [nwarps*16][HD]
    uint8_t* sK = sQ + (size_t)nwarps * 16 * ROWB;            // [2][64][HD]
    uint8_t* sV = sK + 2 * kKeyTile * ROWB;                   // [2][64][HD]
    const uint32_t uQ = smem_u32(sQ), uK = smem_u32(sK), uV = smem_u32(sV);

    const int seq = blockIdx.x, g = blockIdx.y, sp = blockIdx.z;
    const int G = a.nh / a.nkv;
    const int q0 = a.cu_q[seq], qlen = a.cu_q[seq + 1] - q0;
    if (qlen <= 0) return;
    const int R = qlen * G;
    const int kv_len = a.positions[q0 + qlen - 1] + 1;
    const int kbeg = sp * a.split_keys;
    if (kbeg >= kv_len) return;
    const int kend = min(kv_len, kbeg + a.split_keys);
    const int slot = a.seq_slot[seq];
    const int* pt = a.page_table + (size_t)slot * a.max_pages;

    // ---- stage Q (rows r = t*G + gq -> token q0+t, head g*G+gq), swizzled
    for (int c = threadIdx.x; c < nwarps * 16 * CH; c += blockDim.x) {
        const int r = c / CH, ch = c - r * CH;
        const bool ok = r < R;
        const int t = ok ? r / G : 0, gq = ok ? r - t * G : 0;
        const __nv_bfloat16* src = a.q + ((size_t)(q0 + t) * a.nh + g * G + gq) * HD + ch * 8;
        cp_async16(uQ + r * ROWB + ((ch ^ (r & 7)) << 4), src, ok);
    }
    auto load_tile = [&](int tile, int buf) {
        const int j0 = kbeg + tile * kKeyTile;
        for (int c = threadIdx.x; c < kKeyTile * CH; c += blockDim.x) {
            const int r = c / CH, ch = c - r * CH;
            const int j = j0 + r;
            const bool ok = j < kend;
            const int jj = ok ? j : kbeg;
            const int page = pt[jj / a.page_size];
            const size_t off = (((size_t)page * a.nkv + g) * a.page_size + jj % a.page_size) * HD + ch * 8;
            const uint32_t d = (uint32_t)(buf * kKeyTile * ROWB + r * ROWB + ((ch ^ (r & 7)) << 4));
            cp_async16(uK + d, a.k_cache + off, ok);
            cp_async16(uV + d, a.v_cache + off, ok);
        }
    };
    const int ntiles = (kend - kbeg + kKeyTile - 1) / kKeyTile;
    load_tile(0, 0);
    cp_async_commit();

    // per-thread rows
    const int r0 = warp * 16 + (lane >> 2), r1 = r0 + 8;
    const int qpos0 = kv_len - qlen + (r0 < R ? r0 / G : 0), qpos1 = kv_len - qlen + (r1 < R ? r1 / G : 0);
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
    float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    uint32_t qf[KB][4];
    bool q_loaded = false;

    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) load_tile(tile + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (!q_loaded) {
            const int mi = lane >> 3;
            const int row = warp * 16 + (lane & 7) + (mi & 1) * 8;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                const int ch = kb * 2 + (mi >> 1);
                ldsm_x4(uQ + row * ROWB + ((ch ^ (row & 7)) << 4), qf[kb]);
            }
            q_loaded = true;
        }
        // ---- S = Q K^T for 64 keys
        float s[kKeyTile / 8][4];
#pragma unroll
        for (int i = 0; i < kKeyTile / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f;
        const uint32_t kb_base = uK + buf * kKeyTile * ROWB, vb_base = uV + buf * kKeyTile * ROWB;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
            for (int np = 0; np < kKeyTile / 16; ++np) {
                const int mi = lane >> 3;
                const int row = np * 16 + (lane & 7) + (mi >> 1) * 8;
                const int ch = kb * 2 + (mi & 1);
                uint32_t b[4];
                ldsm_x4(kb_base + row * ROWB + ((ch ^ (row & 7)) << 4), b);
                mma_bf16(s[np * 2], qf[kb], b[0], b[1]);
                mma_bf16(s[np * 2 + 1], qf[kb], b[2], b[3]);
            }
        }
        // ---- mask, scale, online softmax
        const int jbase = kbeg + tile * kKeyTile + (lane & 3) * 2;
        float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < kKeyTile / 8; ++i) {
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
        tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 1));
        tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 2));
        tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 1));
        tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 2));
        const float mn0 = fmaxf(mx0, tm0), mn1 = fmaxf(mx1, tm1);
        // rows with nothing visible yet keep a finite reference so exp2f never sees (-inf) - (-inf)
        const float ref0 = mn0 == -INFINITY ? 0.0f : mn0, ref1 = mn1 == -INFINITY ? 0.0f : mn1;
        const float c0 = exp2f(mx0 - ref0), c1 = exp2f(mx1 - ref1);
        float ps0 = 0.0f, ps1 = 0.0f;
        uint32_t pf[kKeyTile / 16][4];
#pragma unroll
        for (int i = 0; i < kKeyTile / 8; ++i) {
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
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < kKeyTile / 16; ++kk) {
#pragma unroll
            for (int dn = 0; dn < HD / 16; ++dn) {
                const int mi = lane >> 3;
                const int row = kk * 16 + (lane & 7) + (mi & 1) * 8;
                const int ch = dn * 2 + (mi >> 1);
                uint32_t b[4];
                ldsm_x4_t(vb_base + row * ROWB + ((ch ^ (row & 7)) << 4), b);
                mma_bf16(o[dn * 2], pf[kk], b[0], b[1]);
                mma_bf16(o[dn * 2 + 1], pf[kk], b[2], b[3]);
            }
        }
        __syncthreads();  // everyone done with this buffer before it is refilled
    }
    cp_async_wait<0>();
    // ---- partial results (unnormalised O, row max, row sum)
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
        const int r = hrow ? r1 : r0;
        if (r >= R) continue;
        const int t = r / G, gq = r - t * G;
        const size_t idx = (((size_t)(q0 + t) * a.nh + g * G + gq) * a.nsplit_max + sp);
        float* op = a.o_part + idx * HD;
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
            const int d = i * 8 + (lane & 3) * 2;
            *reinterpret_cast<float2*>(op + d) = make_float2(o[i][hrow * 2], o[i][hrow * 2 + 1]);
        }
        if ((lane & 3) == 0) {
            a.ml_part[idx * 2] = hrow ? mx1 : mx0;
            a.ml_part[idx * 2 + 1] = hrow ? l1 : l0;
        }
    }
}

// one CTA per (token, head): merge the kv splits
__global__ void attn_combine_kernel(const float* __restrict__ o_part, const float* __restrict__ ml_part,
                                    const int* __restrict__ positions, int split_keys, int nsplit_max, int hd,
                                    __nv_bfloat16* __restrict__ out, int nh) {
    grid_dep_launch();
    const int m = blockIdx.x, hq = blockIdx.y;
    const int ns = min(nsplit_max, (positions[m] + split_keys) / split_keys);  // ceil((pos+1)/split)
    const size_t base = ((size_t)m * nh + hq) * nsplit_max;
    float mx = -INFINITY;
    for (int s = 0; s < ns; ++s) mx = fmaxf(mx, ml_part[(base + s) * 2]);
    float l = 0.0f;
    for (int s = 0; s < ns; ++s) l += ml_part[(base + s) * 2 + 1] * exp2f(ml_part[(base + s) * 2] - mx);
    for (int d = threadIdx.x; d < hd; d += blockDim.x) {
        float acc = 0.0f;
        for (int s = 0; s < ns; ++s) acc += o_part[(base + s) * hd + d] * exp2f(ml_part[(base + s) * 2] - mx);
        out[((size_t)m * nh + hq) * hd + d] = __float2bfloat16(acc / l);
    }
}

int attn_workspace_floats(int M, int nh, int hd, int nsplit_max, size_t* o_floats, size_t* ml_floats) {
    *o_floats = (size_t)M * nh * nsplit_max * hd;
    *ml_floats = (size_t)M * nh * nsplit_max * 2;
    return 0;
}

static int g_attn_attr = 0;

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
    const int warps = (rows + 15) / 16;
    if (warps > 8) return set_error("attention: q_len * group = %d rows exceeds 128; chunk the query", rows);
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
    a.split_keys = L.split_keys;
    a.nsplit_max = L.nsplit_max;
    a.scale_log2 = scale_log2;
    a.o_part = L.o_part;
    a.ml_part = L.ml_part;
    const size_t smem = (size_t)warps * 16 * L.hd * 2 + 4 * (size_t)kKeyTile * L.hd * 2;
    if (!g_attn_attr) {
        ASD_CUDA(cudaFuncSetAttribute(attn_mma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        ASD_CUDA(cudaFuncSetAttribute(attn_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        g_attn_attr = 1;
    }
    const dim3 grid(L.nseq, L.nkv, L.nsplit_max);
    if (L.hd == 128)
        attn_mma_kernel<128><<<grid, warps * 32, smem, stream>>>(a);
    else
        attn_mma_kernel<64><<<grid, warps * 32, smem, stream>>>(a);
    ASD_CUDA(cudaGetLastError());
    attn_combine_kernel<<<dim3(L.M, L.nh), 64, 0, stream>>>(L.o_part, L.ml_part, L.positions, L.split_keys,
                                                           L.nsplit_max, L.hd, L.out, L.nh);
    ASD_CUDA(cudaGetLastError());
    count_launch(2);
    return 0;
}

}  // namespace asd
