// Fused logits -> softmax -> rejection sampling -> residual resample -> stop-rule features, as a STREAMING pass:
// every logit of the verify step is read from HBM once, by an independent CTA, with no vocabulary-wide exchange
// in front of the exponentials.
//
//   sampler_stats_kernel   grid = rows x chunks (a chunk = 4096 consecutive logits of one row, the draft row's
//     chunk next to the target's), 256 threads; the chunk pair lands in shared memory through two 1-D bulk (TMA)
//     copies, so no register holds data in flight: 40 registers, 6 CTAs = 48 warps and 192 KB of loads per SM.
//     Per chunk: maxima (warp shuffles + one shared-memory hop), exponentials
//     relative to the CHUNK maximum (polynomial exp2 on the packed fp32x2 pipe), lane sums and the canonical
//     warp scan; one 32-byte record per (row, chunk).  The last CTA of a row to finish (atomic ticket) merges the
//     row's records (rescaling by 2^(m_c - M)), runs the min(1, p/q) accept test in binary64 and writes the
//     features; the last row of a sequence finds the first-reject prefix and emits the accepted tokens.
//   sampler_draw_kernel    grid = sequences x chunks, only for the ONE row per sequence that emits a new token (first
//     rejected position or bonus row): residual max(0, p - q) chunk totals, then the last CTA of the sequence walks
//     the chunk totals, re-reads the selected chunk (L2) and finishes the inverse-CDF draw inside it.  Launched
//     programmatically behind the first kernel and gated per sequence by a flag, so it overlaps the first kernel's
//     tail; not launched at all for greedy verification.
// DRAM traffic = (1 + ~1/(k+1)) x the algorithmic bytes.  The previous design (a cluster of 8 CTAs holding a row
// pair on chip through three dependent cluster-wide exchanges, round 1) read every byte exactly once but ran at
// 0.15-0.19 of HBM speed: one row pair occupied 8 SMs for ~15 us of barrier latency.
//
// ARITHMETIC CONTRACT v2 (bit-exact with oracle/sampler_oracle.c - see that file's header): every contract
// operation below is an explicit round-to-nearest intrinsic so that nvcc can neither contract nor reorder it.
// Inside a chunk, local element u lives in lane (u / 4) mod 256 = thread, visited in increasing u.
//
// Replaces (reference): per-token softmax + log(probs[token]) with a host sync per token,
// /root/reference/src/training/generate_training_data.py:128-134, and the logprob feature
// reductions of docs/guides/RESEARCH_PROTOCOL.md:379-398.  The accept/resample rule itself has no
// reference implementation.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "asd_internal.h"
#include "ptx.cuh"

namespace asd {

constexpr int kChunk = 4096;             // logits per chunk
constexpr int kST = 256;                 // threads per CTA = lanes per chunk
constexpr int kSW = kST / 32;            // warps
constexpr int kSlots = kChunk / 4 / kST; // float4 per thread and row (4)
constexpr int kMaxChunks = 52;           // V <= 212992
constexpr int kNFeat = ASD_NUM_FEATURES;

struct SamplerParams {
    const float* target;       // [B, k+1, V]
    const float* draft;        // [B, k, V] or null
    const int* draft_tokens;   // [B, k]
    const double* u_accept;    // [B, k]
    const double* u_resid;     // [B]
    int B, k, V, NC, greedy, flags;   // flags: 4 = the draw kernel is gated per sequence (no grid-wide wait), 8 = acq_rel ticket
    float c1;
    // outputs
    uint8_t* accept_mask;
    int* accepted_len;
    int* out_tokens;
    float* out_logprobs;
    float* features;
    // workspace (tickets zero between launches)
    int* row_ticket;    // [rows]
    int* seq_ticket;    // [B]
    int* seq_ticket2;   // [B]
    int* need_row;      // [B] position that emits the new token
    float4* rec_p;      // [rows][NC] {m, m2, Z, S} of the target chunk
    float4* rec_q;      // [rows][NC] {mq, Zq, e at the chunk arg-max, arg-max index (bits)}
    float* row_px;      // [rows] chunk-relative e_p of the draft token
    float* row_qx;      // [rows] chunk-relative e_q of the draft token
    float4* row_stat;   // [rows] {M, Zp, Mq, Zq}
    int* row_accept;    // [rows]
    float* row_lpx;     // [rows]
    int* row_amax;      // [rows]
    float* row_lpamax;  // [rows]
    float* seq_rc;      // [B][NC] residual chunk totals of the emitting row
};

__device__ __forceinline__ float exp2p(float t) {
    const float tc = fmaxf(t, -125.0f);
    const float r = __fadd_rn(tc, 12582912.0f);
    const float nf = __fadd_rn(r, -12582912.0f);
    const float f = __fadd_rn(tc, -nf);
    float p = 0x1.5c08e6p-10f;
    p = __fmaf_rn(p, f, 0x1.3d0c52p-7f);
    p = __fmaf_rn(p, f, 0x1.c6b6e4p-5f);
    p = __fmaf_rn(p, f, 0x1.ebf918p-3f);
    p = __fmaf_rn(p, f, 0x1.62e428p-1f);
    p = __fmaf_rn(p, f, 0x1.000002p+0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
// two lanes of the same polynomial on the packed fp32x2 pipe (FFMA2 / FADD2, sm_100): each component is an IEEE
// fma / add, so the results are bit-identical to the scalar contract
__device__ __forceinline__ float2 exp2p2(float2 t) {
    const float2 tc = make_float2(fmaxf(t.x, -125.0f), fmaxf(t.y, -125.0f));
    const float2 magic = make_float2(12582912.0f, 12582912.0f), nmagic = make_float2(-12582912.0f, -12582912.0f);
    const float2 r = __fadd2_rn(tc, magic);
    const float2 nf = __fadd2_rn(r, nmagic);
    const float2 f = __ffma2_rn(nf, make_float2(-1.0f, -1.0f), tc);   // tc - nf, exact
    float2 q = make_float2(0x1.5c08e6p-10f, 0x1.5c08e6p-10f);
    q = __ffma2_rn(q, f, make_float2(0x1.3d0c52p-7f, 0x1.3d0c52p-7f));
    q = __ffma2_rn(q, f, make_float2(0x1.c6b6e4p-5f, 0x1.c6b6e4p-5f));
    q = __ffma2_rn(q, f, make_float2(0x1.ebf918p-3f, 0x1.ebf918p-3f));
    q = __ffma2_rn(q, f, make_float2(0x1.62e428p-1f, 0x1.62e428p-1f));
    q = __ffma2_rn(q, f, make_float2(0x1.000002p+0f, 0x1.000002p+0f));
    return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(r.x) << 23)),
                       __int_as_float(__float_as_int(q.y) + (__float_as_int(r.y) << 23)));
}

__device__ __forceinline__ int atom_add_release(int* p, int v) {
    int old;
    asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

// Hillis-Steele inclusive scan inside the warp (the contract's order)
__device__ __forceinline__ float warp_hs(float x, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x = __fadd_rn(x, up);
    }
    return x;
}

// e = exp2p(fma(z, c1, -m)) for the four logits of a float4, in place
__device__ __forceinline__ void exp4(float4& z, float2 c1v, float2 nmv, float2& t0, float2& t1) {
    t0 = __ffma2_rn(make_float2(z.x, z.y), c1v, nmv);
    t1 = __ffma2_rn(make_float2(z.z, z.w), c1v, nmv);
    const float2 e0 = exp2p2(t0), e1 = exp2p2(t1);
    z = make_float4(e0.x, e0.y, e1.x, e1.y);
}

// ------------------------------------------------------------------------------------------------ kernel 1
// The chunk (16 KB target + 16 KB draft) lands in shared memory through two 1-D bulk (TMA) copies, so no register
// holds data in flight: 42 registers per thread, 6 CTAs = 48 warps per SM, up to 192 KB of loads in flight per SM.
__global__ void __launch_bounds__(kST, 6) sampler_stats_kernel(const SamplerParams p) {
    __shared__ __align__(128) float4 s_zt[kChunk / 4];
    __shared__ __align__(128) float4 s_zq[kChunk / 4];
    __shared__ __align__(16) float4 s_recp[kMaxChunks], s_recq[kMaxChunks];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ float s_red[kSW][4], s_red2[kSW][4];
    __shared__ unsigned long long s_key[kSW];
    __shared__ int s_last;
    grid_dep_launch();      // the draw kernel may be scheduled; it waits for this grid before reading anything
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k1 = p.k + 1, V = p.V, NC = p.NC;
    const int total = p.B * k1 * NC;
    const float c1 = p.c1;
    const float2 c1v = make_float2(c1, c1);
    // thread 0: bulk copies of one (row, chunk) item into the chunk buffers
    auto issue = [&](int item) {
        const int row = item / NC, c = item - row * NC;
        const int b = row / k1, i = row - b * k1;
        const bool hd = (i < p.k) && !p.greedy;
        const uint32_t bytes = (uint32_t)min(kChunk / 4, (V - c * kChunk) >> 2) * 16u;
        mbar_expect_tx(&s_bar, hd ? 2 * bytes : bytes);
        bulk_g2s(s_zt, p.target + (size_t)row * V + (size_t)c * kChunk, bytes, &s_bar);
        if (hd) bulk_g2s(s_zq, p.draft + ((size_t)b * p.k + i) * V + (size_t)c * kChunk, bytes, &s_bar);
    };
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        issue(blockIdx.x);
    }
    __syncthreads();            // the barrier is initialised
    uint32_t phase = 0;
    // Persistent CTAs, items dealt round-robin: the bulk copies of a CTA's NEXT item are issued as soon as every thread
    // has consumed the current chunk, so they fly while thread 0 finishes the chunk (sums, record, fence, ticket).
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int row = item / NC, c = item - row * NC;
    const int b = row / k1, i = row - b * k1;
    const bool has_draft = (i < p.k) && !p.greedy;
    const int n4 = min(kChunk / 4, (V - c * kChunk) >> 2);     // float4s of this chunk that exist
    mbar_wait(&s_bar, phase);
    phase ^= 1;

    // ---- chunk maxima on the raw logits: rn multiplication by c1 > 0 is monotone, so the two largest a = z * c1 of
    // the multiset are the products of the two largest z (exact, order independent)
    float m1 = -INFINITY, m2 = -INFINITY, mq = -INFINITY;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        if (s * kST + tid < n4) {
            const float4 z = s_zt[s * kST + tid];
            const float a[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                m2 = fmaxf(m2, fminf(m1, a[e]));
                m1 = fmaxf(m1, a[e]);
            }
            if (has_draft) {
                const float4 q = s_zq[s * kST + tid];
                mq = fmaxf(mq, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
            }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const float o1 = __shfl_xor_sync(0xffffffffu, m1, d), o2 = __shfl_xor_sync(0xffffffffu, m2, d);
        m2 = fmaxf(fminf(m1, o1), fmaxf(m2, o2));
        m1 = fmaxf(m1, o1);
        mq = fmaxf(mq, __shfl_xor_sync(0xffffffffu, mq, d));
    }
    if (lane == 0) {
        s_red[warp][0] = m1;
        s_red[warp][1] = m2;
        s_red[warp][2] = mq;
    }
    __syncthreads();
    m1 = m2 = mq = -INFINITY;
#pragma unroll
    for (int w = 0; w < kSW; ++w) {
        const float o1 = s_red[w][0], o2 = s_red[w][1];
        m2 = fmaxf(fminf(m1, o1), fmaxf(m2, o2));
        m1 = fmaxf(m1, o1);
        mq = fmaxf(mq, s_red[w][2]);
    }
    m1 = __fmul_rn(m1, c1);
    m2 = __fmul_rn(m2, c1);
    mq = __fmul_rn(mq, c1);

    // ---- exponentials relative to the chunk maximum; a lane keeps two running sums (elements with even / odd local
    // index, each left to right) on the packed fp32x2 pipe and adds them at the end
    float2 zs2 = make_float2(0.f, 0.f), ss2 = zs2, qs2 = zs2;
    unsigned long long key = ~0ull;                      // greedy: (lowest index attaining the maximum, its e)
    const float2 nm1v = make_float2(-m1, -m1), nmqv = make_float2(-mq, -mq);
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        if (s * kST + tid < n4) {
            float4 z = s_zt[s * kST + tid];
            const float4 zraw = z;
            float2 t0, t1;
            exp4(z, c1v, nm1v, t0, t1);
            zs2 = __fadd2_rn(__fadd2_rn(zs2, make_float2(z.x, z.y)), make_float2(z.z, z.w));
            ss2 = __fadd2_rn(__fadd2_rn(ss2, __fmul2_rn(make_float2(z.x, z.y), t0)), __fmul2_rn(make_float2(z.z, z.w), t1));
            if (p.greedy) {
                const float zz[4] = {zraw.x, zraw.y, zraw.z, zraw.w};
                const float ee[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (__fmul_rn(zz[e], c1) == m1) {
                        const unsigned long long kk =
                            ((unsigned long long)(unsigned)(c * kChunk + (s * kST + tid) * 4 + e) << 32) | __float_as_uint(ee[e]);
                        key = kk < key ? kk : key;
                    }
            }
            if (has_draft) {
                float4 q = s_zq[s * kST + tid];
                float2 u0, u1;
                exp4(q, c1v, nmqv, u0, u1);
                qs2 = __fadd2_rn(__fadd2_rn(qs2, make_float2(q.x, q.y)), make_float2(q.z, q.w));
            }
        }
    }
    const float zs = __fadd_rn(zs2.x, zs2.y), ss = __fadd_rn(ss2.x, ss2.y), qs = __fadd_rn(qs2.x, qs2.y);
    // ---- the draft token's exponentials (one thread recomputes them from the chunk in shared memory)
    const int x = (i < p.k) ? p.draft_tokens[b * p.k + i] : -1;
    const bool x_ok = (i < p.k) && x >= 0 && x < V;
    if (tid == 32 && x_ok && x / kChunk == c) {
        const int u = x - c * kChunk;
        p.row_px[row] = exp2p(__fmaf_rn(reinterpret_cast<const float*>(s_zt)[u], c1, -m1));
        if (has_draft) p.row_qx[row] = exp2p(__fmaf_rn(reinterpret_cast<const float*>(s_zq)[u], c1, -mq));
    }
    // ---- canonical chunk totals: Hillis-Steele inside the warp, sequential chain over the 8 warps
    const float hz = warp_hs(zs, lane), hs_ = warp_hs(ss, lane), hq = warp_hs(qs, lane);
    if (p.greedy) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, d);
            key = o < key ? o : key;
        }
    }
    if (lane == 31) {
        s_red2[warp][0] = hz;
        s_red2[warp][1] = hs_;
        s_red2[warp][2] = hq;
    }
    if (lane == 0) s_key[warp] = key;
    __syncthreads();            // every thread is done with the chunk buffers
    if (tid == 0) {
        fence_proxy_async_smem();   // generic reads of the buffers ordered before the async-proxy writes of the next copy
        if (item + (int)gridDim.x < total) issue(item + gridDim.x);
        float Z = 0.0f, S = 0.0f, Zq = 0.0f;
        unsigned long long kk = ~0ull;
        for (int w = 0; w < kSW; ++w) {
            Z = __fadd_rn(Z, s_red2[w][0]);
            S = __fadd_rn(S, s_red2[w][1]);
            Zq = __fadd_rn(Zq, s_red2[w][2]);
            kk = s_key[w] < kk ? s_key[w] : kk;
        }
        __stcg(&p.rec_p[(size_t)row * NC + c], make_float4(m1, m2, Z, S));
        __stcg(&p.rec_q[(size_t)row * NC + c],
               make_float4(mq, Zq, __uint_as_float((unsigned)(kk & 0xffffffffu)), __int_as_float((int)(kk >> 32))));
        if (p.flags & 8) {
            s_last = atom_add_release(&p.row_ticket[row], 1) == NC - 1;   // orders the record stores before the ticket
        } else {
            __threadfence();
            s_last = atomicAdd(&p.row_ticket[row], 1) == NC - 1;
        }
    }
    __syncthreads();
    if (!s_last) continue;

    // ================================================================ last CTA of the row: merge its chunk records
    __threadfence();
    if (tid < NC) {
        s_recp[tid] = __ldcg(&p.rec_p[(size_t)row * NC + tid]);
        s_recq[tid] = __ldcg(&p.rec_q[(size_t)row * NC + tid]);
    }
    __syncthreads();
    if (tid != 0) continue;
    float M = -INFINITY, M2 = -INFINITY, Mq = -INFINITY;
    for (int c2 = 0; c2 < NC; ++c2) {
        M2 = fmaxf(fminf(M, s_recp[c2].x), fmaxf(M2, s_recp[c2].y));
        M = fmaxf(M, s_recp[c2].x);
        Mq = fmaxf(Mq, s_recq[c2].x);
    }
    float Zp = 0.0f, St = 0.0f, Zq = 0.0f, s_x = 0.0f, sq_x = 0.0f, P_amax = 0.0f;
    int amax = -1;
    const int cx = x_ok ? x / kChunk : -1;
    for (int c2 = 0; c2 < NC; ++c2) {
        const float d = __fadd_rn(s_recp[c2].x, -M);
        const float sc = exp2p(d);
        Zp = __fadd_rn(Zp, __fmul_rn(s_recp[c2].z, sc));
        St = __fadd_rn(St, __fmul_rn(__fmaf_rn(s_recp[c2].z, d, s_recp[c2].w), sc));
        if (c2 == cx) s_x = sc;
        if (p.greedy && amax < 0 && s_recp[c2].x == M) {
            amax = __float_as_int(s_recq[c2].w);
            P_amax = __fmul_rn(s_recq[c2].z, sc);
        }
        if (has_draft) {
            const float sq = exp2p(__fadd_rn(s_recq[c2].x, -Mq));
            Zq = __fadd_rn(Zq, __fmul_rn(s_recq[c2].y, sq));
            if (c2 == cx) sq_x = sq;
        }
    }
    const float Px = x_ok ? __fmul_rn(__ldcg(&p.row_px[row]), s_x) : 0.0f;
    int acc = 0;
    if (p.greedy) {
        acc = x_ok && x == amax;
    } else if (has_draft && x_ok) {
        const float Qx = __fmul_rn(__ldcg(&p.row_qx[row]), sq_x);
        const double lhs = __dmul_rn(__dmul_rn(p.u_accept[b * p.k + i], (double)Qx), (double)Zp);
        const double rhs = __dmul_rn((double)Px, (double)Zq);
        acc = lhs <= rhs;
    }
    const float logZ = logf(Zp), log2Z = log2f(Zp);
    float* f = p.features + (size_t)row * kNFeat;
    f[0] = __fmul_rn(__fadd_rn(M, log2Z), 0x1.62e43p-1f);
    f[1] = 1.0f / Zp;
    f[2] = f[1] - exp2p(__fadd_rn(M2, -M)) / Zp;
    f[3] = (log2Z - St / Zp) * 0x1.62e43p-1f;
    f[4] = x_ok ? logf(Px) - logZ : -INFINITY;
    __stcg(&p.row_stat[row], make_float4(M, Zp, Mq, Zq));
    p.row_accept[row] = acc;
    p.row_lpx[row] = f[4];
    if (p.greedy) {
        p.row_amax[row] = amax;
        p.row_lpamax[row] = logf(P_amax) - logZ;
    }
    p.row_ticket[row] = 0;
    __threadfence();
    if (atomicAdd(&p.seq_ticket[b], 1) != p.k) continue;
    // ================================================================ last row of the sequence: first-reject prefix
    __threadfence();
    int n = 0;
    while (n < p.k && __ldcg(&p.row_accept[b * k1 + n])) ++n;
    for (int t = 0; t < p.k; ++t) p.accept_mask[b * p.k + t] = t < n;
    p.accepted_len[b] = n;
    for (int t = 0; t < k1; ++t) {
        if (t == n && !p.greedy) continue;          // written by the draw kernel
        int tok = -1;
        float lp = 0.0f;
        if (t < n) {
            tok = p.draft_tokens[b * p.k + t];
            lp = __ldcg(&p.row_lpx[b * k1 + t]);
        } else if (t == n) {
            tok = __ldcg(&p.row_amax[b * k1 + t]);
            lp = __ldcg(&p.row_lpamax[b * k1 + t]);
        }
        p.out_tokens[b * k1 + t] = tok;
        p.out_logprobs[b * k1 + t] = lp;
        p.features[(size_t)(b * k1 + t) * kNFeat + 5] = lp;
    }
    p.seq_ticket[b] = 0;
    if (!p.greedy) {
        // release: everything the draw kernel reads for this sequence (records, row statistics: written by other CTAs
        // and observed here through the tickets) is ordered before the flag
        __threadfence();
        *reinterpret_cast<volatile int*>(&p.need_row[b]) = n + 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------ kernel 2
// residual weights of this thread's elements of chunk c of the emitting row: r -> zq[], P -> zt[]; returns the lane sum
__device__ __forceinline__ float residual_chunk(const SamplerParams& p, int row, int b, int n, int c, bool has_draft,
                                                const float4& st, float4 (&zt)[kSlots], float4 (&zq)[kSlots],
                                                bool (&have)[kSlots]) {
    const int tid = threadIdx.x, V = p.V;
    const int n4 = min(kChunk / 4, (V - c * kChunk) >> 2);
    const float4* zt_g = reinterpret_cast<const float4*>(p.target + (size_t)row * V + (size_t)c * kChunk);
    const float4* zq_g =
        has_draft ? reinterpret_cast<const float4*>(p.draft + ((size_t)b * p.k + n) * V + (size_t)c * kChunk) : nullptr;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        have[s] = s * kST + tid < n4;
        zt[s] = have[s] ? __ldcg(zt_g + s * kST + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int s = 0; s < kSlots; ++s)
        zq[s] = (has_draft && have[s]) ? __ldcg(zq_g + s * kST + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 rp = __ldcg(&p.rec_p[(size_t)row * p.NC + c]);
    const float4 rq = __ldcg(&p.rec_q[(size_t)row * p.NC + c]);
    const float M = st.x, Zp = st.y, Mq = st.z, Zq = st.w;
    const float sp = exp2p(__fadd_rn(rp.x, -M));
    const float sq = has_draft ? exp2p(__fadd_rn(rq.x, -Mq)) : 0.0f;
    const float2 c1v = make_float2(p.c1, p.c1), nmv = make_float2(-rp.x, -rp.x), nqv = make_float2(-rq.x, -rq.x);
    const float2 spv = make_float2(sp, sp), sqv = make_float2(sq, sq), zqv = make_float2(Zq, Zq), nzpv = make_float2(-Zp, -Zp);
    float rs = 0.0f;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        if (have[s]) {
            float2 t0, t1;
            exp4(zt[s], c1v, nmv, t0, t1);
            const float2 P0 = __fmul2_rn(make_float2(zt[s].x, zt[s].y), spv);
            const float2 P1 = __fmul2_rn(make_float2(zt[s].z, zt[s].w), spv);
            zt[s] = make_float4(P0.x, P0.y, P1.x, P1.y);
            float4 r4;
            if (has_draft) {
                exp4(zq[s], c1v, nqv, t0, t1);
                const float2 Q0 = __fmul2_rn(make_float2(zq[s].x, zq[s].y), sqv);
                const float2 Q1 = __fmul2_rn(make_float2(zq[s].z, zq[s].w), sqv);
                const float2 w0 = __fmul2_rn(Q0, nzpv), w1 = __fmul2_rn(Q1, nzpv);      // -(Q * Zp), exact sign flip
                const float2 r0 = __ffma2_rn(P0, zqv, w0), r1 = __ffma2_rn(P1, zqv, w1);
                r4 = make_float4(fmaxf(r0.x, 0.0f), fmaxf(r0.y, 0.0f), fmaxf(r1.x, 0.0f), fmaxf(r1.y, 0.0f));
            } else {
                r4 = zt[s];
            }
            zq[s] = r4;
            rs = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(rs, r4.x), r4.y), r4.z), r4.w);
        }
    }
    return rs;
}

__global__ void __launch_bounds__(kST, 4) sampler_draw_kernel(const SamplerParams p) {
    __shared__ float s_red[kSW];
    __shared__ float s_rc[kMaxChunks];
    __shared__ int s_flag, s_cstar, s_tstar[kSW];
    __shared__ float s_tau, s_X;
    // Launched programmatically: these CTAs become resident while the statistics kernel drains its last wave and wait
    // for THEIR sequence only (flag = emitting position + 1, set by the sequence's finaliser with release semantics),
    // so the draws of the early sequences overlap the statistics of the late ones.  No deadlock: this grid is only
    // scheduled once every CTA of the statistics kernel has started, and those never wait for anything.
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k1 = p.k + 1, NC = p.NC;
    const int b = blockIdx.x / NC, c = blockIdx.x - b * NC;
    if (!(p.flags & 4)) grid_dep_wait();
    if (tid == 0) {
        int v;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p.need_row + b) : "memory");
        } while (v == 0);
        s_cstar = v - 1;
    }
    __syncthreads();
    const int n = s_cstar;
    __syncthreads();
    const int row = b * k1 + n;
    const bool has_draft = n < p.k;
    const float4 st = __ldcg(&p.row_stat[row]);
    float4 zt[kSlots], zq[kSlots];
    bool have[kSlots];
    float rs = residual_chunk(p, row, b, n, c, has_draft, st, zt, zq, have);
    float hs = warp_hs(rs, lane);
    if (lane == 31) s_red[warp] = hs;
    __syncthreads();
    if (tid == 0) {
        float R = 0.0f;
        for (int w = 0; w < kSW; ++w) R = __fadd_rn(R, s_red[w]);
        __stcg(&p.seq_rc[(size_t)b * NC + c], R);
        __threadfence();
        s_flag = atomicAdd(&p.seq_ticket2[b], 1) == NC - 1;
    }
    __syncthreads();
    if (!s_flag) return;

    // ================================================================ last CTA of the sequence: the draw
    __threadfence();
    if (tid < NC) s_rc[tid] = __ldcg(&p.seq_rc[(size_t)b * NC + tid]);
    __syncthreads();
    if (tid == 0) {
        float R = 0.0f;
        for (int c2 = 0; c2 < NC; ++c2) R = __fadd_rn(R, s_rc[c2]);
        int cstar = -1;
        float X = 0.0f, tau = 0.0f;
        if (R > 0.0f) {
            float ur = (float)p.u_resid[b];
            if (!(ur >= 0.0f)) ur = 0.0f;
            if (ur >= 1.0f) ur = 0x1.fffffep-1f;
            tau = __fmul_rn(ur, R);
            if (!(tau < R)) tau = 0.0f;
            float pre = 0.0f;
            for (int c2 = 0; c2 < NC; ++c2) {
                const float nx = __fadd_rn(pre, s_rc[c2]);
                if (nx > tau) {
                    cstar = c2;
                    X = pre;
                    break;
                }
                pre = nx;
            }
        }
        s_cstar = cstar;
        s_tau = tau;
        s_X = X;
        p.seq_ticket2[b] = 0;
        p.need_row[b] = 0;          // every CTA of this sequence has read the flag (they ticket after reading it)
    }
    __syncthreads();
    const int cstar = s_cstar;
    const float tau = s_tau, X = s_X;
    const float logZ = logf(st.y);
    if (cstar < 0) {
        // residual mass is zero (p == q on this row): the draft token itself is a valid draw
        if (tid == 0) {
            const int x = has_draft ? p.draft_tokens[b * p.k + n] : -1;
            const bool x_ok = has_draft && x >= 0 && x < p.V;
            const int y = x_ok ? x : 0;
            const float4 rp = __ldcg(&p.rec_p[(size_t)row * NC + y / kChunk]);
            const float sp = exp2p(__fadd_rn(rp.x, -st.x));
            const float ey = exp2p(__fmaf_rn(p.target[(size_t)row * p.V + y], p.c1, -rp.x));
            const float lp = logf(__fmul_rn(ey, sp)) - logZ;
            p.out_tokens[row] = y;
            p.out_logprobs[row] = lp;
            p.features[(size_t)row * kNFeat + 5] = lp;
        }
        return;
    }
    if (cstar != c) rs = residual_chunk(p, row, b, n, cstar, has_draft, st, zt, zq, have);
    hs = warp_hs(rs, lane);
    __syncthreads();       // s_red is rewritten
    if (lane == 31) s_red[warp] = hs;
    __syncthreads();
    float off = 0.0f;
    for (int w = 0; w < warp; ++w) off = __fadd_rn(off, s_red[w]);
    const float P = __fadd_rn(X, __fadd_rn(off, hs));
    const float Pprev = __shfl_up_sync(0xffffffffu, P, 1);
    const float Xl = lane == 0 ? __fadd_rn(X, off) : Pprev;
    const unsigned hit = __ballot_sync(0xffffffffu, P > tau);
    if (lane == 0) s_tstar[warp] = hit ? warp * 32 + __ffs(hit) - 1 : kST;
    __syncthreads();
    int tstar = kST;
#pragma unroll
    for (int w = 0; w < kSW; ++w) tstar = min(tstar, s_tstar[w]);
    if (tid != tstar) return;
    float cc = Xl, Py = 0.0f, P_last = 0.0f, P_first = 0.0f;
    int sel = -1, last_pos = -1, first = -1;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        if (have[s] && sel < 0) {
            const int v0 = cstar * kChunk + (s * kST + tid) * 4;
            const float rr[4] = {zq[s].x, zq[s].y, zq[s].z, zq[s].w};
            const float pp[4] = {zt[s].x, zt[s].y, zt[s].z, zt[s].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (sel < 0) {
                    if (first < 0) {
                        first = v0 + e;
                        P_first = pp[e];
                    }
                    if (rr[e] > 0.0f) {
                        last_pos = v0 + e;
                        P_last = pp[e];
                    }
                    cc = __fadd_rn(cc, rr[e]);
                    if (cc > tau) {
                        sel = v0 + e;
                        Py = pp[e];
                    }
                }
            }
        }
    }
    const int y = sel >= 0 ? sel : (last_pos >= 0 ? last_pos : (first >= 0 ? first : 0));
    const float Pv = sel >= 0 ? Py : (last_pos >= 0 ? P_last : P_first);
    const float lp = logf(Pv) - logZ;
    p.out_tokens[row] = y;
    p.out_logprobs[row] = lp;
    p.features[(size_t)row * kNFeat + 5] = lp;
}

// asd_reject_sample_set_impl bit flags: 2 = persistent statistics CTAs (experiment, slower), 4 = draw kernel gated per
// sequence instead of the grid-wide wait (default: +4-7 % at B*k >= 512), 8 = single acq_rel ticket atomic (no gain)
int g_sampler_impl = 5;

static inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

size_t reject_sample_workspace_bytes(int B, int k) {
    const size_t rows = (size_t)B * (k + 1);
    size_t n = 0;
    n += align256(sizeof(int) * rows);                    // row_ticket
    n += 3 * align256(sizeof(int) * (size_t)B);           // seq_ticket, seq_ticket2, need_row
    n += 2 * align256(sizeof(float4) * rows * kMaxChunks);  // rec_p, rec_q
    n += 2 * align256(sizeof(float) * rows);              // row_px, row_qx
    n += align256(sizeof(float4) * rows);                 // row_stat
    n += 4 * align256(sizeof(float) * rows);              // row_accept, row_lpx, row_amax, row_lpamax
    n += align256(sizeof(float) * (size_t)B * kMaxChunks);  // seq_rc
    return n + 256;
}

int launch_reject_sample(const float* target, const float* draft, const int* draft_tokens, const double* u_accept,
                         const double* u_resid, int B, int k, int V, float temperature, uint8_t* accept_mask,
                         int* accepted_len, int* out_tokens, float* out_logprobs, float* features, void* workspace,
                         cudaStream_t stream) {
    if (B <= 0) return 0;
    if (k < 0 || k > 64) return set_error("asd_reject_sample: k must be in [0, 64]");
    if (V < 4 || (V & 3)) return set_error("asd_reject_sample: V must be a positive multiple of 4");
    const int NC = (V + kChunk - 1) / kChunk;
    if (NC > kMaxChunks) return set_error("asd_reject_sample: V must be <= %d", kMaxChunks * kChunk);
    const bool greedy = !(temperature > 0.0f);
    if (!greedy && k > 0 && draft == nullptr) return set_error("asd_reject_sample: draft_logits required");
    if ((reinterpret_cast<uintptr_t>(target) & 15) || (draft && (reinterpret_cast<uintptr_t>(draft) & 15)))
        return set_error("asd_reject_sample: logits must be 16-byte aligned");
    const size_t rows = (size_t)B * (k + 1);
    if (rows * NC > 0x7fffffffull) return set_error("asd_reject_sample: too many rows");

    SamplerParams p;
    p.target = target;
    p.draft = draft;
    p.draft_tokens = draft_tokens;
    p.u_accept = u_accept;
    p.u_resid = u_resid;
    p.B = B;
    p.k = k;
    p.V = V;
    p.NC = NC;
    p.greedy = greedy;
    p.flags = g_sampler_impl;
    p.c1 = greedy ? 0x1.715476p+0f : (1.0f / temperature) * 0x1.715476p+0f;
    p.accept_mask = accept_mask;
    p.accepted_len = accepted_len;
    p.out_tokens = out_tokens;
    p.out_logprobs = out_logprobs;
    p.features = features;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        uint8_t* q = ws + off;
        off += align256(bytes);
        return q;
    };
    p.row_ticket = reinterpret_cast<int*>(take(sizeof(int) * rows));
    p.seq_ticket = reinterpret_cast<int*>(take(sizeof(int) * (size_t)B));
    p.seq_ticket2 = reinterpret_cast<int*>(take(sizeof(int) * (size_t)B));
    p.need_row = reinterpret_cast<int*>(take(sizeof(int) * (size_t)B));
    p.rec_p = reinterpret_cast<float4*>(take(sizeof(float4) * rows * kMaxChunks));
    p.rec_q = reinterpret_cast<float4*>(take(sizeof(float4) * rows * kMaxChunks));
    p.row_px = reinterpret_cast<float*>(take(sizeof(float) * rows));
    p.row_qx = reinterpret_cast<float*>(take(sizeof(float) * rows));
    p.row_stat = reinterpret_cast<float4*>(take(sizeof(float4) * rows));
    p.row_accept = reinterpret_cast<int*>(take(sizeof(float) * rows));
    p.row_lpx = reinterpret_cast<float*>(take(sizeof(float) * rows));
    p.row_amax = reinterpret_cast<int*>(take(sizeof(float) * rows));
    p.row_lpamax = reinterpret_cast<float*>(take(sizeof(float) * rows));
    p.seq_rc = reinterpret_cast<float*>(take(sizeof(float) * (size_t)B * kMaxChunks));

    static int sms[64];
    int dev = 0;
    ASD_CUDA(cudaGetDevice(&dev));
    if (sms[dev & 63] == 0) ASD_CUDA(cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    // One CTA per item by default.  impl 2 (asd_reject_sample_set_impl, experiments): persistent CTAs that prefetch
    // their next item - measured 2.5x SLOWER on B200 (18 us per item and CTA), so the hardware block scheduler keeps
    // the job of refilling SM slots.
    const size_t items = rows * NC, resident = (size_t)6 * sms[dev & 63];
    const size_t grid = (g_sampler_impl == 2 && items > resident) ? resident : items;
    sampler_stats_kernel<<<dim3((unsigned)grid), kST, 0, stream>>>(p);
    ASD_CUDA(cudaGetLastError());
    count_launch(1);
    if (!greedy) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.gridDim = dim3((unsigned)((size_t)B * NC));
        cfg.blockDim = dim3(kST);
        cfg.stream = stream;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        ASD_CUDA(cudaLaunchKernelEx(&cfg, sampler_draw_kernel, p));
        count_launch(1);
    }
    return 0;
}

}  // namespace asd
