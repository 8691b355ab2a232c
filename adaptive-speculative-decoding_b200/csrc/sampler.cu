// Fused logits -> softmax -> rejection sampling -> residual resample -> stop-rule features.
//
// One launch handles every (sequence, position) row of a verify step:
//   * a cluster of 8 CTAs x 512 threads owns one row pair (target + draft logits, fp32);
//     the pair is pulled from HBM exactly once by 1-D bulk (TMA) copies into the cluster's
//     shared memory (<= 2 x 80 KB per CTA for V = 152064) and all three passes (max, exp/sum,
//     residual) run out of shared memory;
//   * cross-CTA reductions go through distributed shared memory;
//   * the last cluster to finish a sequence (atomic ticket) folds the per-row decisions into
//     accept_mask / accepted_len / out_tokens, so no second launch and no logits round trip.
//
// ARITHMETIC CONTRACT (bit-exact with oracle/sampler_oracle.c - see that file's header): every
// contract operation below is an explicit round-to-nearest intrinsic so that nvcc can neither
// contract nor reorder it.  4096 abstract lanes = 8 CTAs x 512 threads; element v lives in lane
// (v/4) mod 4096, i.e. thread (v/4) mod 512 of CTA ((v/4) mod 4096) / 512.
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

constexpr int kCluster = 8;
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kSlabVec = kCluster * kThreads;  // float4s per slab (4096 lanes)
constexpr int kSlabBytesPerCta = kThreads * 16;
constexpr int kMaxSlabs = 13;
constexpr int kNFeat = ASD_NUM_FEATURES;

struct SamplerParams {
    const float* target;       // [B, k+1, V]
    const float* draft;        // [B, k, V] or null
    const int* draft_tokens;   // [B, k]
    const double* u_accept;    // [B, k]
    const double* u_resid;     // [B]
    int B, k, V, num_slabs, greedy;
    float c1;
    // outputs
    uint8_t* accept_mask;
    int* accepted_len;
    int* out_tokens;
    float* out_logprobs;
    float* features;
    // workspace
    int* seq_counter;  // [B] zero between launches
    int* row_accept;   // [B*(k+1)]
    int* row_cand;
    float* row_lpx;
    float* row_lpy;
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

struct SmemCtl {
    uint64_t bar;
    float warp_scratch[kWarps][4];
    float xchg[4][kCluster][4];  // [exchange index][source CTA][value]
    float scal[8];               // written remotely into CTA 0: ep_x, eq_x, ep_y, y(bits), has_y
};

// write n values into slot [ex][my_rank] of every CTA of the cluster
__device__ __forceinline__ void xchg_publish(SmemCtl* ctl, int ex, uint32_t my_rank, const float* v, int n) {
    const uint32_t base = smem_u32(&ctl->xchg[ex][my_rank][0]);
#pragma unroll
    for (int c = 0; c < kCluster; ++c) {
        const uint32_t ra = mapa(base, c);
        for (int i = 0; i < n; ++i) st_cluster_u32(ra + 4 * i, __float_as_uint(v[i]));
    }
}

// exact integer minimum over the whole cluster (order independent)
__device__ __forceinline__ int cluster_min_int(SmemCtl* ctl, int ex, uint32_t crank, int v) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, d));
    __syncthreads();
    if (lane == 0) ctl->warp_scratch[warp][3] = __int_as_float(v);
    __syncthreads();
    if (threadIdx.x == 0) {
        int m = 0x7fffffff;
        for (int w = 0; w < kWarps; ++w) m = min(m, __float_as_int(ctl->warp_scratch[w][3]));
        const float mf = __int_as_float(m);
        xchg_publish(ctl, ex, crank, &mf, 1);
    }
    cluster_sync();
    int m = 0x7fffffff;
#pragma unroll
    for (int c = 0; c < kCluster; ++c) m = min(m, __float_as_int(ctl->xchg[ex][c][0]));
    return m;
}

// Canonical scan of up to N per-thread values over the 4096 lanes of the cluster.
// Returns inclusive P, previous-lane X and total for each of the N values.
template <int N>
__device__ __forceinline__ void cluster_scan(SmemCtl* ctl, int ex, uint32_t crank, const float (&v)[N], float (&P)[N],
                                             float (&X)[N], float (&total)[N]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float hs[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float x = v[i];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x = __fadd_rn(x, up);
        }
        hs[i] = x;
    }
    __syncthreads();  // previous users of warp_scratch are done
    if (lane == 31) {
#pragma unroll
        for (int i = 0; i < N; ++i) ctl->warp_scratch[warp][i] = hs[i];
    }
    __syncthreads();
    float off[N], cta_total[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float o = 0.0f, mine = 0.0f;
        for (int w = 0; w < kWarps; ++w) {
            if (w == warp) mine = o;
            o = __fadd_rn(o, ctl->warp_scratch[w][i]);
        }
        off[i] = mine;
        cta_total[i] = o;
    }
    if (threadIdx.x == 0) xchg_publish(ctl, ex, crank, cta_total, N);
    cluster_sync();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float co = 0.0f, mine = 0.0f;
#pragma unroll
        for (int c = 0; c < kCluster; ++c) {
            if (c == (int)crank) mine = co;
            co = __fadd_rn(co, ctl->xchg[ex][c][i]);
        }
        total[i] = co;
        const float q = __fadd_rn(off[i], hs[i]);
        P[i] = __fadd_rn(mine, q);
        const float prev = __shfl_up_sync(0xffffffffu, P[i], 1);
        X[i] = lane == 0 ? __fadd_rn(mine, off[i]) : prev;
    }
}

// Row epilogue executed by ONE thread of the cluster (rank 0, thread 0): accept test, features, per-row
// scratch, and - for the last row of a sequence to finish - the first-reject prefix / emitted tokens.
__device__ __noinline__ void row_epilogue(const SamplerParams& p, const SmemCtl* ctl, int row, int b, int i, int x,
                                          bool x_ok, bool has_draft, int amax, float m1, float m2, float Zp, float Zq,
                                          float Ssum, const float* zt) {
    const float c1 = p.c1, nm1 = -m1;
    const int k1 = p.k + 1;

            const float ep_x = ctl->scal[0], eq_x = ctl->scal[1];
            const bool has_y = ctl->scal[4] != 0.0f;
            int y;
            float ep_y;
            int acc = 0;
            if (p.greedy) {
                y = amax;
                acc = x_ok && x == y;
                ep_y = exp2p(__fmaf_rn(zt[y], c1, nm1));
            } else {
                if (has_draft && x_ok) {
                    const double lhs = __dmul_rn(__dmul_rn(p.u_accept[b * p.k + i], (double)eq_x), (double)Zp);
                    const double rhs = __dmul_rn((double)ep_x, (double)Zq);
                    acc = lhs <= rhs;
                }
                if (has_y) {
                    y = __float_as_int(ctl->scal[3]);
                    ep_y = ctl->scal[2];
                } else {  // R == 0: p == q on this row
                    y = x_ok ? x : 0;
                    ep_y = x_ok ? ep_x : exp2p(__fmaf_rn(zt[0], c1, nm1));
                }
            }
            const float logZ = logf(Zp), log2Z = log2f(Zp);
            float* f = p.features + (size_t)row * kNFeat;
            f[0] = __fmul_rn(__fadd_rn(m1, log2Z), 0x1.62e43p-1f);
            f[1] = 1.0f / Zp;
            f[2] = f[1] - exp2p(__fadd_rn(m2, nm1)) / Zp;
            f[3] = (log2Z - Ssum / Zp) * 0x1.62e43p-1f;
            f[4] = x_ok ? logf(ep_x) - logZ : -INFINITY;
            f[5] = logf(ep_y) - logZ;
            p.row_accept[row] = acc;
            p.row_cand[row] = y;
            p.row_lpx[row] = f[4];
            p.row_lpy[row] = f[5];
            __threadfence();
            const int ticket = atomicAdd(&p.seq_counter[b], 1);
            if (ticket == p.k) {  // last row of this sequence: first-reject prefix + emitted tokens
                __threadfence();
                int n = 0;
                while (n < p.k && __ldcg(&p.row_accept[b * k1 + n])) ++n;
                for (int t = 0; t < p.k; ++t) p.accept_mask[b * p.k + t] = t < n;
                p.accepted_len[b] = n;
                for (int t = 0; t < k1; ++t) {
                    int tok = -1;
                    float lp = 0.0f;
                    if (t < n) {
                        tok = p.draft_tokens[b * p.k + t];
                        lp = __ldcg(&p.row_lpx[b * k1 + t]);
                    } else if (t == n) {
                        tok = __ldcg(&p.row_cand[b * k1 + t]);
                        lp = __ldcg(&p.row_lpy[b * k1 + t]);
                    }
                    p.out_tokens[b * k1 + t] = tok;
                    p.out_logprobs[b * k1 + t] = lp;
                }
                p.seq_counter[b] = 0;
            }
        }

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    reject_sample_kernel(const SamplerParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float4* bufP = reinterpret_cast<float4*>(smem_raw);
    float4* bufQ = bufP + p.num_slabs * kThreads;
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(bufQ + p.num_slabs * kThreads);

    const uint32_t crank = cluster_ctarank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cluster_id = blockIdx.x / kCluster, nclusters = gridDim.x / kCluster;
    const int k1 = p.k + 1, rows = p.B * k1, V = p.V, nvec = V >> 2;
    const float c1 = p.c1;

    if (tid == 0) {
        mbar_init(&ctl->bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    cluster_sync();

    uint32_t phase = 0;
    for (int row = cluster_id; row < rows; row += nclusters) {
        const int b = row / k1, i = row - b * k1;
        const bool has_draft = (i < p.k) && !p.greedy;
        const float* zt = p.target + (size_t)row * V;
        const float* zq = has_draft ? p.draft + ((size_t)b * p.k + i) * V : nullptr;

        // ------------------------------------------------------------ load: HBM -> smem, once
        if (tid == 0) {
            uint32_t bytes = 0;
            for (int s = 0; s < p.num_slabs; ++s) {
                const long long off = (long long)s * kSlabVec * 16 + (long long)crank * kSlabBytesPerCta;
                long long n = (long long)V * 4 - off;
                n = n > kSlabBytesPerCta ? kSlabBytesPerCta : n;
                if (n > 0) bytes += (uint32_t)n * (has_draft ? 2 : 1);
            }
            mbar_expect_tx(&ctl->bar, bytes);
            for (int s = 0; s < p.num_slabs; ++s) {
                const long long off = (long long)s * kSlabVec * 16 + (long long)crank * kSlabBytesPerCta;
                long long n = (long long)V * 4 - off;
                n = n > kSlabBytesPerCta ? kSlabBytesPerCta : n;
                if (n > 0) {
                    bulk_g2s(bufP + s * kThreads, reinterpret_cast<const uint8_t*>(zt) + off, (uint32_t)n, &ctl->bar);
                    if (has_draft)
                        bulk_g2s(bufQ + s * kThreads, reinterpret_cast<const uint8_t*>(zq) + off, (uint32_t)n,
                                 &ctl->bar);
                }
            }
            ctl->scal[4] = 0.0f;  // has_y flag (CTA 0's copy is the one that is read)
        }
        mbar_wait(&ctl->bar, phase);
        phase ^= 1;

        // ------------------------------------------------------------ pass 1: maxima
        float m1 = -INFINITY, m2 = -INFINITY, mq = -INFINITY;
        for (int s = 0; s < p.num_slabs; ++s) {
            const int j = s * kSlabVec + (int)crank * kThreads + tid;
            if (j < nvec) {
                const float4 z = bufP[s * kThreads + tid];
                const float a[4] = {__fmul_rn(z.x, c1), __fmul_rn(z.y, c1), __fmul_rn(z.z, c1), __fmul_rn(z.w, c1)};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    m2 = fmaxf(m2, fminf(m1, a[e]));
                    m1 = fmaxf(m1, a[e]);
                }
                if (has_draft) {
                    const float4 q = bufQ[s * kThreads + tid];
                    mq = fmaxf(mq, fmaxf(fmaxf(__fmul_rn(q.x, c1), __fmul_rn(q.y, c1)),
                                         fmaxf(__fmul_rn(q.z, c1), __fmul_rn(q.w, c1))));
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
        __syncthreads();
        if (lane == 0) {
            ctl->warp_scratch[warp][0] = m1;
            ctl->warp_scratch[warp][1] = m2;
            ctl->warp_scratch[warp][2] = mq;
        }
        __syncthreads();
        if (tid == 0) {
            float v[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int w = 0; w < kWarps; ++w) {
                const float o1 = ctl->warp_scratch[w][0], o2 = ctl->warp_scratch[w][1];
                v[1] = fmaxf(fminf(v[0], o1), fmaxf(v[1], o2));
                v[0] = fmaxf(v[0], o1);
                v[2] = fmaxf(v[2], ctl->warp_scratch[w][2]);
            }
            xchg_publish(ctl, 0, crank, v, 3);
        }
        cluster_sync();
        m1 = m2 = mq = -INFINITY;
#pragma unroll
        for (int c = 0; c < kCluster; ++c) {
            const float o1 = ctl->xchg[0][c][0], o2 = ctl->xchg[0][c][1];
            m2 = fmaxf(fminf(m1, o1), fmaxf(m2, o2));
            m1 = fmaxf(m1, o1);
            mq = fmaxf(mq, ctl->xchg[0][c][2]);
        }

        // ------------------------------------------------------------ pass 2: e = 2^(a - m), sums
        float sums[3] = {0.0f, 0.0f, 0.0f};  // Z_p, S, Z_q lane sums
        int amax = 0x7fffffff;
        const float nm1 = -m1, nmq = -mq;
        for (int s = 0; s < p.num_slabs; ++s) {
            const int j = s * kSlabVec + (int)crank * kThreads + tid;
            if (j < nvec) {
                float4 z = bufP[s * kThreads + tid];
                float* zz = reinterpret_cast<float*>(&z);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (p.greedy && __fmul_rn(zz[e], c1) == m1) amax = min(amax, j * 4 + e);
                    const float t = __fmaf_rn(zz[e], c1, nm1);
                    const float ex = exp2p(t);
                    sums[0] = __fadd_rn(sums[0], ex);
                    sums[1] = __fadd_rn(sums[1], __fmul_rn(ex, t));
                    zz[e] = ex;
                }
                bufP[s * kThreads + tid] = z;
                if (has_draft) {
                    float4 q = bufQ[s * kThreads + tid];
                    float* qq = reinterpret_cast<float*>(&q);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float ex = exp2p(__fmaf_rn(qq[e], c1, nmq));
                        sums[2] = __fadd_rn(sums[2], ex);
                        qq[e] = ex;
                    }
                    bufQ[s * kThreads + tid] = q;
                }
            }
        }
        float Pz[3], Xz[3], tot[3];
        cluster_scan<3>(ctl, 1, crank, sums, Pz, Xz, tot);
        const float Zp = tot[0], Ssum = tot[1], Zq = tot[2];

        // greedy: lowest index attaining the maximum (integer min is exact in any order)
        if (p.greedy) amax = cluster_min_int(ctl, 3, crank, amax);

        // ------------------------------------------------------------ pass 3: residual + inverse CDF
        const int x = (i < p.k) ? p.draft_tokens[b * p.k + i] : -1;
        const bool x_ok = (i < p.k) && x >= 0 && x < V;
        const uint32_t scal0 = mapa(smem_u32(&ctl->scal[0]), 0);
        if (x_ok) {  // the thread that owns element x reports e_p[x], e_q[x] to CTA 0
            const int j = x >> 2, s = j / kSlabVec, l = j - s * kSlabVec;
            if (l / kThreads == (int)crank && l % kThreads == tid) {
                st_cluster_f32(scal0 + 0, reinterpret_cast<const float*>(&bufP[s * kThreads + tid])[x & 3]);
                if (has_draft)
                    st_cluster_f32(scal0 + 4, reinterpret_cast<const float*>(&bufQ[s * kThreads + tid])[x & 3]);
            }
        }
        if (!p.greedy) {
            float rs[1] = {0.0f};
            for (int s = 0; s < p.num_slabs; ++s) {
                const int j = s * kSlabVec + (int)crank * kThreads + tid;
                if (j < nvec) {
                    const float4 ev = bufP[s * kThreads + tid];
                    const float* ee = reinterpret_cast<const float*>(&ev);
                    if (has_draft) {
                        const float4 qv = bufQ[s * kThreads + tid];
                        const float* qq = reinterpret_cast<const float*>(&qv);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float wq = __fmul_rn(qq[e], Zp);
                            rs[0] = __fadd_rn(rs[0], fmaxf(__fmaf_rn(ee[e], Zq, -wq), 0.0f));
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) rs[0] = __fadd_rn(rs[0], ee[e]);
                    }
                }
            }
            float Pr[1], Xr[1], Rt[1];
            cluster_scan<1>(ctl, 2, crank, rs, Pr, Xr, Rt);
            float ur = (float)p.u_resid[b];
            if (!(ur >= 0.0f)) ur = 0.0f;
            if (ur >= 1.0f) ur = 0x1.fffffep-1f;
            const float tau = __fmul_rn(ur, Rt[0]);
            if (Rt[0] > 0.0f && Pr[0] > tau && Xr[0] <= tau) {  // exactly one thread of the cluster
                float c = Xr[0];
                int sel = -1, last_pos = -1;
                for (int s = 0; s < p.num_slabs && sel < 0; ++s) {
                    const int j = s * kSlabVec + (int)crank * kThreads + tid;
                    if (j < nvec) {
                        const float* ee = reinterpret_cast<const float*>(&bufP[s * kThreads + tid]);
                        const float* qq = reinterpret_cast<const float*>(&bufQ[s * kThreads + tid]);
                        for (int e = 0; e < 4 && sel < 0; ++e) {
                            float r;
                            if (has_draft) {
                                const float wq = __fmul_rn(qq[e], Zp);
                                r = fmaxf(__fmaf_rn(ee[e], Zq, -wq), 0.0f);
                            } else {
                                r = ee[e];
                            }
                            if (r > 0.0f) last_pos = j * 4 + e;
                            c = __fadd_rn(c, r);
                            if (c > tau) sel = j * 4 + e;
                        }
                    }
                }
                const int y = sel >= 0 ? sel : last_pos;
                const int jy = y >> 2, sy = jy / kSlabVec;
                st_cluster_f32(scal0 + 8, reinterpret_cast<const float*>(&bufP[sy * kThreads + tid])[y & 3]);
                st_cluster_u32(scal0 + 12, (uint32_t)y);
                st_cluster_f32(scal0 + 16, 1.0f);
            }
        }
        fence_proxy_async_smem();  // generic-proxy smem traffic ordered before the next row's bulk copies
        cluster_sync();

        // ------------------------------------------------------------ row epilogue (one thread)
        if (crank == 0 && tid == 0)
            row_epilogue(p, ctl, row, b, i, x, x_ok, has_draft, amax, m1, m2, Zp, Zq, Ssum, zt);
        // CTA 0's leader must finish reading scal[] before any thread of the next row's pass 3
        // writes it: those writes happen after that row's cluster barriers, which the leader joins.
    }
    cluster_sync();
}

// ------------------------------------------------------------------------------------------------
// Register-resident variant (V <= 10 slabs = 163840): identical arithmetic contract, different data
// movement.  Shared memory is only a PREFETCH buffer: bulk TMA copies bring row r+1 in while row r is
// processed out of registers (each thread keeps its NS float4 of the target and of the draft row), so the
// HBM stream overlaps all three passes and no pass touches shared memory for data.  The polynomial
// exp2 and the element-wise products run on the packed fp32x2 pipe (FFMA2 / FADD2 / FMUL2, sm_100):
// each component is an IEEE fma/add/mul, so results are bit-identical to the scalar contract.
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

template <int NS>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
    reject_sample_reg_kernel(const SamplerParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float4* bufP = reinterpret_cast<float4*>(smem_raw);
    float4* bufQ = bufP + NS * kThreads;
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(bufQ + NS * kThreads);

    const uint32_t crank = cluster_ctarank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cluster_id = blockIdx.x / kCluster, nclusters = gridDim.x / kCluster;
    const int k1 = p.k + 1, rows = p.B * k1, V = p.V, nvec = V >> 2;
    const float c1 = p.c1;
    const float2 c1v = make_float2(c1, c1);

    auto issue_row = [&](int row) {   // thread 0: bulk copies of this CTA's share of a row (pair) into smem
        const int b = row / k1, i = row - b * k1;
        const bool hd = (i < p.k) && !p.greedy;
        const uint8_t* zt = reinterpret_cast<const uint8_t*>(p.target + (size_t)row * V);
        const uint8_t* zq = hd ? reinterpret_cast<const uint8_t*>(p.draft + ((size_t)b * p.k + i) * V) : nullptr;
        uint32_t bytes = 0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const long long off = (long long)s * kSlabVec * 16 + (long long)crank * kSlabBytesPerCta;
            long long n = (long long)V * 4 - off;
            n = n > kSlabBytesPerCta ? kSlabBytesPerCta : n;
            if (n > 0) bytes += (uint32_t)n * (hd ? 2 : 1);
        }
        mbar_expect_tx(&ctl->bar, bytes);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const long long off = (long long)s * kSlabVec * 16 + (long long)crank * kSlabBytesPerCta;
            long long n = (long long)V * 4 - off;
            n = n > kSlabBytesPerCta ? kSlabBytesPerCta : n;
            if (n > 0) {
                bulk_g2s(bufP + s * kThreads, zt + off, (uint32_t)n, &ctl->bar);
                if (hd) bulk_g2s(bufQ + s * kThreads, zq + off, (uint32_t)n, &ctl->bar);
            }
        }
    };

    if (tid == 0) {
        mbar_init(&ctl->bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    cluster_sync();
    if (tid == 0 && cluster_id < rows) issue_row(cluster_id);

    uint32_t phase = 0;
    for (int row = cluster_id; row < rows; row += nclusters) {
        const int b = row / k1, i = row - b * k1;
        const bool has_draft = (i < p.k) && !p.greedy;
        const float* zt_g = p.target + (size_t)row * V;

        // ------------------------------------------------------------ smem -> registers, then prefetch the next row
        mbar_wait(&ctl->bar, phase);
        phase ^= 1;
        float4 zt[NS], zq[NS];
        bool have[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int j = s * kSlabVec + (int)crank * kThreads + tid;
            have[s] = j < nvec;
            zt[s] = have[s] ? bufP[s * kThreads + tid] : make_float4(0.f, 0.f, 0.f, 0.f);
            zq[s] = (have[s] && has_draft) ? bufQ[s * kThreads + tid] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (tid == 0) ctl->scal[4] = 0.0f;
        fence_proxy_async_smem();
        __syncthreads();   // every thread has taken its data: the buffer is free for the next row
        if (tid == 0 && row + nclusters < rows) issue_row(row + nclusters);

        // ------------------------------------------------------------ pass 1: maxima
        float m1 = -INFINITY, m2 = -INFINITY, mq = -INFINITY;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (have[s]) {
                const float2 a0 = __fmul2_rn(make_float2(zt[s].x, zt[s].y), c1v);
                const float2 a1 = __fmul2_rn(make_float2(zt[s].z, zt[s].w), c1v);
                const float a[4] = {a0.x, a0.y, a1.x, a1.y};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    m2 = fmaxf(m2, fminf(m1, a[e]));
                    m1 = fmaxf(m1, a[e]);
                }
                if (has_draft) {
                    const float2 q0 = __fmul2_rn(make_float2(zq[s].x, zq[s].y), c1v);
                    const float2 q1 = __fmul2_rn(make_float2(zq[s].z, zq[s].w), c1v);
                    mq = fmaxf(mq, fmaxf(fmaxf(q0.x, q0.y), fmaxf(q1.x, q1.y)));
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
            ctl->warp_scratch[warp][0] = m1;
            ctl->warp_scratch[warp][1] = m2;
            ctl->warp_scratch[warp][2] = mq;
        }
        __syncthreads();
        if (tid == 0) {
            float v[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int w = 0; w < kWarps; ++w) {
                const float o1 = ctl->warp_scratch[w][0], o2 = ctl->warp_scratch[w][1];
                v[1] = fmaxf(fminf(v[0], o1), fmaxf(v[1], o2));
                v[0] = fmaxf(v[0], o1);
                v[2] = fmaxf(v[2], ctl->warp_scratch[w][2]);
            }
            xchg_publish(ctl, 0, crank, v, 3);
        }
        cluster_sync();
        m1 = m2 = mq = -INFINITY;
#pragma unroll
        for (int c = 0; c < kCluster; ++c) {
            const float o1 = ctl->xchg[0][c][0], o2 = ctl->xchg[0][c][1];
            m2 = fmaxf(fminf(m1, o1), fmaxf(m2, o2));
            m1 = fmaxf(m1, o1);
            mq = fmaxf(mq, ctl->xchg[0][c][2]);
        }

        // ------------------------------------------------------------ pass 2: e = 2^(a - m) in registers, sums
        float sums[3] = {0.0f, 0.0f, 0.0f};
        int amax = 0x7fffffff;
        const float nm1 = -m1, nmq = -mq;
        const float2 nm1v = make_float2(nm1, nm1), nmqv = make_float2(nmq, nmq);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (have[s]) {
                const int j = s * kSlabVec + (int)crank * kThreads + tid;
                if (p.greedy) {
                    const float zz[4] = {zt[s].x, zt[s].y, zt[s].z, zt[s].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (__fmul_rn(zz[e], c1) == m1) amax = min(amax, j * 4 + e);
                }
                const float2 t0 = __ffma2_rn(make_float2(zt[s].x, zt[s].y), c1v, nm1v);
                const float2 t1 = __ffma2_rn(make_float2(zt[s].z, zt[s].w), c1v, nm1v);
                const float2 e0 = exp2p2(t0), e1 = exp2p2(t1);
                const float2 w0 = __fmul2_rn(e0, t0), w1 = __fmul2_rn(e1, t1);
                sums[0] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sums[0], e0.x), e0.y), e1.x), e1.y);
                sums[1] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sums[1], w0.x), w0.y), w1.x), w1.y);
                zt[s] = make_float4(e0.x, e0.y, e1.x, e1.y);
                if (has_draft) {
                    const float2 q0 = exp2p2(__ffma2_rn(make_float2(zq[s].x, zq[s].y), c1v, nmqv));
                    const float2 q1 = exp2p2(__ffma2_rn(make_float2(zq[s].z, zq[s].w), c1v, nmqv));
                    sums[2] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sums[2], q0.x), q0.y), q1.x), q1.y);
                    zq[s] = make_float4(q0.x, q0.y, q1.x, q1.y);
                }
            }
        }
        float Pz[3], Xz[3], tot[3];
        cluster_scan<3>(ctl, 1, crank, sums, Pz, Xz, tot);
        const float Zp = tot[0], Ssum = tot[1], Zq = tot[2];
        if (p.greedy) amax = cluster_min_int(ctl, 3, crank, amax);

        // ------------------------------------------------------------ pass 3: residual + inverse CDF
        const int x = (i < p.k) ? p.draft_tokens[b * p.k + i] : -1;
        const bool x_ok = (i < p.k) && x >= 0 && x < V;
        const uint32_t scal0 = mapa(smem_u32(&ctl->scal[0]), 0);
        if (x_ok) {
            const int jx = x >> 2, sx = jx / kSlabVec, lx = jx - sx * kSlabVec;
            if (lx / kThreads == (int)crank && lx % kThreads == tid) {
                float epx = 0.f, eqx = 0.f;
#pragma unroll
                for (int s = 0; s < NS; ++s)
                    if (s == sx) {
                        const float pe[4] = {zt[s].x, zt[s].y, zt[s].z, zt[s].w};
                        const float qe[4] = {zq[s].x, zq[s].y, zq[s].z, zq[s].w};
                        epx = pe[x & 3];
                        eqx = qe[x & 3];
                    }
                st_cluster_f32(scal0 + 0, epx);
                if (has_draft) st_cluster_f32(scal0 + 4, eqx);
            }
        }
        if (!p.greedy) {
            // residual weights replace the draft registers: r = max(0, fma(e_p, Z_q, -(e_q * Z_p)))
            float rs[1] = {0.0f};
            const float2 zqv = make_float2(Zq, Zq), nzpv = make_float2(-Zp, -Zp);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (have[s]) {
                    float4 r4;
                    if (has_draft) {
                        const float2 w0 = __fmul2_rn(make_float2(zq[s].x, zq[s].y), nzpv);   // -(e_q * Z_p), exact sign flip
                        const float2 w1 = __fmul2_rn(make_float2(zq[s].z, zq[s].w), nzpv);
                        const float2 r0 = __ffma2_rn(make_float2(zt[s].x, zt[s].y), zqv, w0);
                        const float2 r1 = __ffma2_rn(make_float2(zt[s].z, zt[s].w), zqv, w1);
                        r4 = make_float4(fmaxf(r0.x, 0.0f), fmaxf(r0.y, 0.0f), fmaxf(r1.x, 0.0f), fmaxf(r1.y, 0.0f));
                    } else {
                        r4 = zt[s];
                    }
                    rs[0] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(rs[0], r4.x), r4.y), r4.z), r4.w);
                    zq[s] = r4;
                }
            }
            float Pr[1], Xr[1], Rt[1];
            cluster_scan<1>(ctl, 2, crank, rs, Pr, Xr, Rt);
            float ur = (float)p.u_resid[b];
            if (!(ur >= 0.0f)) ur = 0.0f;
            if (ur >= 1.0f) ur = 0x1.fffffep-1f;
            const float tau = __fmul_rn(ur, Rt[0]);
            if (Rt[0] > 0.0f && Pr[0] > tau && Xr[0] <= tau) {   // exactly one thread of the cluster
                float c = Xr[0], epy = 0.0f, ep_last = 0.0f;
                int sel = -1, last_pos = -1;
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    if (have[s] && sel < 0) {
                        const int j = s * kSlabVec + (int)crank * kThreads + tid;
                        const float rr[4] = {zq[s].x, zq[s].y, zq[s].z, zq[s].w};
                        const float pe[4] = {zt[s].x, zt[s].y, zt[s].z, zt[s].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (sel < 0) {
                                if (rr[e] > 0.0f) {
                                    last_pos = j * 4 + e;
                                    ep_last = pe[e];
                                }
                                c = __fadd_rn(c, rr[e]);
                                if (c > tau) {
                                    sel = j * 4 + e;
                                    epy = pe[e];
                                }
                            }
                        }
                    }
                }
                const int y = sel >= 0 ? sel : last_pos;
                st_cluster_f32(scal0 + 8, sel >= 0 ? epy : ep_last);
                st_cluster_u32(scal0 + 12, (uint32_t)y);
                st_cluster_f32(scal0 + 16, 1.0f);
            }
        }
        cluster_sync();
        if (crank == 0 && tid == 0)
            row_epilogue(p, ctl, row, b, i, x, x_ok, has_draft, amax, m1, m2, Zp, Zq, Ssum, zt_g);
    }
    cluster_sync();
}

static int g_max_clusters[2 * kMaxSlabs + 16];
int g_sampler_impl = 1;   // 1: register-resident kernel when V allows, 0: always the shared-memory kernel

size_t reject_sample_workspace_bytes(int B, int k) {
    const size_t rows = (size_t)B * (k + 1);
    return sizeof(int) * (size_t)B + rows * (2 * sizeof(int) + 2 * sizeof(float)) + 256;
}

int launch_reject_sample(const float* target, const float* draft, const int* draft_tokens, const double* u_accept,
                         const double* u_resid, int B, int k, int V, float temperature, uint8_t* accept_mask,
                         int* accepted_len, int* out_tokens, float* out_logprobs, float* features, void* workspace,
                         cudaStream_t stream) {
    if (B <= 0) return 0;
    if (k < 0 || k > 64) return set_error("asd_reject_sample: k must be in [0, 64]");
    if (V < 4 || (V & 3)) return set_error("asd_reject_sample: V must be a positive multiple of 4");
    const int num_slabs = ((V >> 2) + kSlabVec - 1) / kSlabVec;
    if (num_slabs > kMaxSlabs) return set_error("asd_reject_sample: V too large for shared-memory residency");
    const bool greedy = !(temperature > 0.0f);
    if (!greedy && k > 0 && draft == nullptr) return set_error("asd_reject_sample: draft_logits required");
    if ((reinterpret_cast<uintptr_t>(target) & 15) || (draft && (reinterpret_cast<uintptr_t>(draft) & 15)))
        return set_error("asd_reject_sample: logits must be 16-byte aligned");

    SamplerParams p;
    p.target = target;
    p.draft = draft;
    p.draft_tokens = draft_tokens;
    p.u_accept = u_accept;
    p.u_resid = u_resid;
    p.B = B;
    p.k = k;
    p.V = V;
    p.num_slabs = num_slabs;
    p.greedy = greedy;
    p.c1 = greedy ? 0x1.715476p+0f : (1.0f / temperature) * 0x1.715476p+0f;
    p.accept_mask = accept_mask;
    p.accepted_len = accepted_len;
    p.out_tokens = out_tokens;
    p.out_logprobs = out_logprobs;
    p.features = features;
    const size_t rows = (size_t)B * (k + 1);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    p.seq_counter = reinterpret_cast<int*>(ws);
    size_t off = ((sizeof(int) * (size_t)B + 255) / 256) * 256;
    p.row_accept = reinterpret_cast<int*>(ws + off);
    p.row_cand = p.row_accept + rows;
    p.row_lpx = reinterpret_cast<float*>(p.row_cand + rows);
    p.row_lpy = p.row_lpx + rows;

    // register-resident kernel for NS in {1, 2, 4, 7, 10}; the shared-memory-resident kernel above covers larger V
    static const int kRegNs[] = {1, 2, 4, 7, 10};
    int ns = 0;
    if (g_sampler_impl != 0)
        for (int c : kRegNs)
            if (c >= num_slabs) {
                ns = c;
                break;
            }
    const void* fn = (const void*)reject_sample_kernel;
    switch (ns) {
        case 1: fn = (const void*)reject_sample_reg_kernel<1>; break;
        case 2: fn = (const void*)reject_sample_reg_kernel<2>; break;
        case 4: fn = (const void*)reject_sample_reg_kernel<4>; break;
        case 7: fn = (const void*)reject_sample_reg_kernel<7>; break;
        case 10: fn = (const void*)reject_sample_reg_kernel<10>; break;
        default: break;
    }
    const int smem_slabs = ns ? ns : num_slabs;
    const size_t smem = (size_t)smem_slabs * kSlabBytesPerCta * 2 + sizeof(SmemCtl);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int slot = ns ? kMaxSlabs + 1 + ns : num_slabs;   // cache index of (kernel, footprint)
    static PerDeviceOnce attr_once[2 * kMaxSlabs + 16];   // the shared-memory opt-in is per (kernel, device)
    int dev = 0;
    if (attr_once[slot].need(&dev)) {
        int optin = 0;
        ASD_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        ASD_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        cfg.gridDim = dim3(kCluster * 148);
        int n = 0;
        ASD_CUDA(cudaOccupancyMaxActiveClusters(&n, fn, &cfg));
        if (n <= 0) return set_error("asd_reject_sample: no cluster of 8 CTAs fits on this device");
        g_max_clusters[slot] = n;
    }
    const int nclusters = (int)(rows < (size_t)g_max_clusters[slot] ? rows : g_max_clusters[slot]);
    cfg.gridDim = dim3(kCluster * nclusters);
    void* args[] = {(void*)&p};
    ASD_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    count_launch(1);
    return 0;
}

}  // namespace asd
