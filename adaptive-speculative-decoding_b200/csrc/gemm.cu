// Weight-streaming bf16 GEMM for the verify / draft forward:  Y[m, n] = sum_k X[m, k] * W[n, k].
//
// The verify step has few rows (M = B*(k+1) = 16..256 tokens) and huge weights, so it is bound by
// streaming W from HBM once.  B200-first design ("swap-AB"):
//   * the WEIGHT tile (128 rows of W, K-major) is the tcgen05 A operand (UMMA M = 128), the token
//     tile (MT = M rounded up to 16, <= 256 rows of X) is the B operand (UMMA N = MT): no tensor
//     work is wasted on padding rows and one TMEM accumulator [128 lanes x MT columns] holds the tile;
//   * both operands arrive by TMA (128-byte swizzle) into a multi-stage mbarrier ring; W with an
//     L2 evict-first policy (read exactly once), X with evict-last (re-read by every CTA);
//   * warp 0 = per-token rstd of the fused RMSNorm, warp 1 = TMEM allocator + single-thread tcgen05.mma
//     issuer, warps 2..5 = TMA producers during the main loop (lane 0 of each owns the ring slots s % 4) and
//     epilogue afterwards (tcgen05.ld -> registers);
//   * grid = (weight tiles, K splits, token tiles); the host picks the split so that all CTAs are
//     co-resident in ONE wave (2-3 CTAs/SM) - with HBM as the shared bottleneck every CTA then
//     progresses at the same rate and there is no tail;
//   * the K splits of a tile form a thread-block cluster (1, ksplit, 1): CTA r owns the token columns
//     [r*MT/ks, (r+1)*MT/ks), every CTA scatters those columns of its TMEM tile into the owner's idle tile
//     ring through DSMEM, the owner adds the partials in rank order (deterministic, no atomics, no HBM round
//     trip) and runs the epilogue for its tokens; without the cluster (option reduce = 0) the partials are fp32
//     slices that the glue kernels sum.
// Epilogues, all rolled 128-bit loops (a CTA runs its epilogue once: unrolled code is paid in instruction-cache
// misses): fp32 (+ residual accumulate, + bf16(resid * ln_w) and per-token sum of squares for the fused RMSNorm,
// + the tensor-parallel all-reduce over NVLink peer memory), QKV (rstd, bias, rotate-half RoPE, q store, paged
// K/V append), bf16, SwiGLU for the gate|up projection whose rows are pre-interleaved (64 gate rows, 64 up rows
// per tile) so silu(g)*u never round-trips to HBM, and fp32 logits * rstd for the lm_head.
// Programmatic dependent launch: weight stages are issued before griddepcontrol.wait, activations after it.
//
// Stands behind Stage.generate's model forward, which the reference delegates to vLLM
// (/root/reference/src/serving/real_model_pipeline.py:98-108,135).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "asd_internal.h"
#include "gemm.h"
#include "ptx.cuh"

namespace asd {

constexpr int kTileN = 128;       // weight rows per CTA (UMMA M)
constexpr int kBlockK = 64;       // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kGemmThreads = 192; // 6 warps
constexpr int kABytes = kTileN * kBlockK * 2;

struct GemmArgs {
    int M, N, K;        // logical problem
    int MT;             // token tile (multiple of 16, <= 256)
    int kblocks;        // ceil(K / 64)
    int ksplit;
    int stages;
    int mode;
    int ldo;            // leading dimension of the output (elements)
    int n_valid;        // rows of the output that exist (SwiGLU: ff; else N)
    void* out;
    uint32_t tmem_cols;
    int reduce;         // 1: the K splits of a tile form a cluster and reduce through DSMEM
    int accumulate;     // fp32 output is added to what is already there (fused residual add)
    int recv_dedicated; // small token tiles: the DSMEM receive buffer has its own smem, so no barrier before the scatter
    int early_trigger;  // issue griddepcontrol.launch_dependents at kernel start instead of after the main loop
    QkvEpilogue qkv;    // GEMM_OUT_QKV only
    NormFusion norm;
    // optional L2 prefetch of the NEXT GEMM's weights from the idle MMA warp once this cluster has streamed its own
    // (engine option next_prefetch_mb; measured neutral, off by default)
    int next_ntiles, next_ksplit, next_kblocks, next_kp;   // next_kp = k-blocks per next-kernel CTA to prefetch (0 = off)
    TpFusion tp;        // world > 1: all-reduce over peer memory inside the owner epilogue
    unsigned long long* trace;  // diagnostics: kTraceSlots globaltimer stamps per CTA (nullptr = off)
};

constexpr int kTraceSlots = 16;
__device__ __forceinline__ void trace_stamp(const GemmArgs& a, int slot) {
    if (a.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        a.trace[(size_t)cta * kTraceSlots + slot] = t;
    }
}

__device__ __forceinline__ float silu_mul(float g, float u) { return __fdividef(g, 1.0f + __expf(-g)) * u; }

// L2 prefetch of the next kernel's weights: boxes k-block-major (the first k-block of every next CTA first), dealt
// round-robin over this grid's CTAs and `nissue` issuing threads per CTA (this thread is number `who`)
__device__ __forceinline__ void prefetch_next_weights(const GemmArgs& a, const CUtensorMap* tmap_next, int who, int nissue) {
    const int ncta_next = a.next_ntiles * a.next_ksplit, nbox = ncta_next * a.next_kp;
    const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const int nthr = gridDim.x * gridDim.y * gridDim.z * nissue;
    const int nbase = a.next_kblocks / a.next_ksplit, nrem = a.next_kblocks % a.next_ksplit;
    for (int b = cta * nissue + who; b < nbox; b += nthr) {
        const int j = b / ncta_next, c = b - j * ncta_next;
        const int nt = c % a.next_ntiles, ns = c / a.next_ntiles;
        const int nkb_c = nbase + (ns < nrem ? 1 : 0);
        if (j >= nkb_c) continue;
        const int kb = ns * nbase + (ns < nrem ? ns : nrem) + j;
        tma_prefetch_l2_2d(tmap_next, kb * kBlockK, nt * kTileN);
    }
}

__global__ void __launch_bounds__(kGemmThreads, 3) gemm_ws_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                                const __grid_constant__ CUtensorMap tmap_x,
                                                                const __grid_constant__ CUtensorMap tmap_next,
                                                                const GemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic smem is only guaranteed 16-byte aligned: align the tile ring to 1024 by hand
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = kABytes + a.MT * 128;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + a.stages;
    uint64_t* tmem_full = empty_bar + a.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    uint64_t* rstd_bar = reinterpret_cast<uint64_t*>(tmem_slot + 4);
    float* rstd_s = reinterpret_cast<float*>(rstd_bar + 1);   // [MT] per-token rstd (fused RMSNorm consumer)
    float* xbuf = rstd_s + a.MT;  // SwiGLU exchange [2][64][33] floats / sum-of-squares exchange [256 + ks*MT]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_n = blockIdx.x, split = blockIdx.y, tile_m = blockIdx.z;
    if (threadIdx.x == 0) {
        trace_stamp(a, 0);
        if (a.trace) {
            unsigned sm;
            asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
            a.trace[(size_t)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) * kTraceSlots + 9] = sm;
        }
    }
    const int base = a.kblocks / a.ksplit, rem = a.kblocks % a.ksplit;
    const int kb0 = split * base + (split < rem ? split : rem);
    const int nkb = base + (split < rem ? 1 : 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(rstd_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, a.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) trace_stamp(a, 1);
    // early trigger: the dependent grid becomes schedulable once every CTA of this grid is resident, so its
    // weight stream fills the SMs that this grid's tail leaves idle (it still waits before reading our output)
    if (a.early_trigger) grid_dep_launch();
    // phase 0 of the cluster barrier only proves that every CTA of the cluster is running (DSMEM is legal from
    // then on); arriving here and waiting right before the scatter makes it free
    if (a.reduce && a.recv_dedicated) cluster_arrive();

    if (warp >= 2) {
        // ------------------------------------------------------------------ TMA producers
        // Lane 0 of each of the four epilogue warps (idle during the main loop) owns the ring slots
        // s with s % nprod == its index, so a producer is never more than one phase ahead of a slot's
        // barriers (mbarrier parity waits are only unambiguous then).
        if (lane == 0) {
            const int pidx = warp - 2;   // 0..3
            const int nprod = a.stages < 4 ? a.stages : 4;
            const uint64_t pol_w = policy_evict_first(), pol_x = policy_evict_last();
            bool waited = false;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % a.stages, round = kb / a.stages;
                if (s % nprod != pidx) continue;
                if (round > 0) mbar_wait(&empty_bar[s], (round & 1) ^ 1);
                mbar_expect_tx(&full_bar[s], stage_bytes);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                // weights do not depend on the upstream kernel: issue them before the PDL wait
                tma_load_2d_hint(st, &tmap_w, (kb0 + kb) * kBlockK, tile_n * kTileN, &full_bar[s], pol_w);
                if (!waited) {
                    grid_dep_wait();
                    waited = true;
                    if (pidx == 0) trace_stamp(a, 2);
                }
                tma_load_2d_hint(st + kABytes, &tmap_x, (kb0 + kb) * kBlockK, tile_m * a.MT, &full_bar[s], pol_x);
            }
        }
        __syncwarp();
    }
    if (warp == 0) {
        // ------------------------------------------------------------------ fused RMSNorm: per-token rstd
        if (a.norm.sumsq_in != nullptr) {
            grid_dep_wait();
            for (int c = lane; c < a.MT; c += 32) {
                const int m = tile_m * a.MT + c;
                float ssum = 0.0f;
                if (m < a.M)
                    for (int t = 0; t < a.norm.parts; ++t) ssum += a.norm.sumsq_in[(size_t)t * a.norm.ld + m];
                rstd_s[c] = rsqrtf(ssum / (float)a.norm.hidden + a.norm.eps);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(rstd_bar);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma_idesc_bf16(kTileN, a.MT);
        int s = 0;
        uint32_t ph = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (lane == 0) {
                if (kb == 0) trace_stamp(a, 3);
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + kABytes);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)  // 16 bf16 = 32 bytes = +2 in the >>4 address field
                    umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                umma_commit(&empty_bar[s]);
                if (kb == nkb - 1) umma_commit(tmem_full);
            }
            __syncwarp();
            if (++s == a.stages) {
                s = 0;
                ph ^= 1;
            }
        }
        if (nkb == 0 && lane == 0) mbar_arrive(tmem_full);
        if (!a.reduce && a.next_kp > 0 && lane < 4) {   // no cluster phase: prefetch once this CTA's MMAs are done
            mbar_wait(tmem_full, 0);
            prefetch_next_weights(a, &tmap_next, lane, 4);
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;             // accumulator row = weight row inside the tile
        const int m0 = tile_m * a.MT;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (threadIdx.x == 64) trace_stamp(a, 4);
        grid_dep_wait();     // (already satisfied) makes the upstream grid's writes to `out` visible here
        grid_dep_launch();
        const bool scale = a.norm.sumsq_in != nullptr;
        if (scale) mbar_wait(rstd_bar, 0);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (a.reduce) {
            // handled below by the whole cluster
        } else if (a.mode == GEMM_OUT_SWIGLU) {
            // rows 0..63 = gate, 64..127 = up of ff index tile_n*64 + (row & 63).  The accumulator tile is
            // transposed through the (now idle) tile ring as T[token][row]; then each thread combines 8
            // consecutive ff rows of one token and writes them with one 16-byte store.  Rolled loops: see the
            // note on instruction-cache misses in the cluster epilogue below.
            float* T = reinterpret_cast<float*>(smem);
            const int et = threadIdx.x - 64;      // 0..127 among epilogue threads
            __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a.out);
#pragma unroll 1
            for (int c0 = 0; c0 < a.MT; c0 += 32) {
                uint32_t r[32];
                if (nkb > 0) {
                    tmem_ld_32x32(taddr + c0, r);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = 0;
                }
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c0 + i < a.MT) T[(c0 + i) * kTileN + row] = __uint_as_float(r[i]);
            }
            if (threadIdx.x == 64) trace_stamp(a, 5);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const bool wide = (a.ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
#pragma unroll 1
            for (int g = et; g < a.MT * 8; g += 128) {
                const int col = g >> 3, f8 = (g & 7) * 8;
                const int m = m0 + col, j0 = tile_n * 64 + f8;
                if (m >= a.M || j0 >= a.n_valid) continue;
                const float rs = scale ? rstd_s[col] : 1.0f;
                const float4* tp = reinterpret_cast<const float4*>(T + col * kTileN + f8);
                const float4 g0 = tp[0], g1 = tp[1], u0 = tp[16], u1 = tp[17];
                float v[8];
                v[0] = silu_mul(g0.x * rs, u0.x * rs), v[1] = silu_mul(g0.y * rs, u0.y * rs);
                v[2] = silu_mul(g0.z * rs, u0.z * rs), v[3] = silu_mul(g0.w * rs, u0.w * rs);
                v[4] = silu_mul(g1.x * rs, u1.x * rs), v[5] = silu_mul(g1.y * rs, u1.y * rs);
                v[6] = silu_mul(g1.z * rs, u1.z * rs), v[7] = silu_mul(g1.w * rs, u1.w * rs);
                __nv_bfloat16* o = out + (size_t)m * a.ldo + j0;
                if (wide && j0 + 8 <= a.n_valid) {
                    __nv_bfloat162 p[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(p);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (j0 + i < a.n_valid) o[i] = __float2bfloat16(v[i]);
                }
            }
            if (threadIdx.x == 64) trace_stamp(a, 10);
        } else {
            const int n = tile_n * kTileN + row;
            for (int c0 = 0; c0 < a.MT; c0 += 32) {
                uint32_t r[32];
                if (nkb > 0) {
                    tmem_ld_32x32(taddr + c0, r);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = 0;
                }
                if (n < a.n_valid) {
                    if (a.mode == GEMM_OUT_BF16) {
                        __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a.out);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int m = m0 + c0 + i;
                            if (c0 + i < a.MT && m < a.M) out[(size_t)m * a.ldo + n] = __float2bfloat16(__uint_as_float(r[i]));
                        }
                    } else {  // fp32, one [M, ldo] slice per K split
                        float* out = static_cast<float*>(a.out) + (size_t)split * a.M * a.ldo;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int m = m0 + c0 + i;
                            if (c0 + i < a.MT && m < a.M) {
                                float* o = out + (size_t)m * a.ldo + n;
                                const float v = scale ? __uint_as_float(r[i]) * rstd_s[c0 + i] : __uint_as_float(r[i]);
                                *o = a.accumulate ? *o + v : v;
                            }
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    if (a.reduce) {
        // ---- split-K reduction inside the cluster (deterministic, no HBM round trip of partials):
        // CTA r of the cluster owns the token columns [r*cols_per, (r+1)*cols_per) of the tile; every CTA sends
        // those columns of its partial accumulator (all 128 rows) into the owner's shared memory through DSMEM,
        // the owner adds the ksplit contributions in rank order and runs the epilogue for its tokens.  Owning
        // whole tokens keeps every run contiguous: a warp's DSMEM store is 128 bytes, an owner reads float4s and
        // writes 512-byte rows, the per-token sum of squares and the rotary pairs of a head never leave the CTA.
        // The owner loops are kept rolled and 128-bit wide on purpose: every CTA runs them exactly once, so
        // straight-line unrolled code is paid for in instruction-cache misses, not saved.
        const int ks = a.ksplit, cols_per = (a.MT + ks - 1) / ks;
        const uint32_t my_rank = cluster_ctarank();
        const bool qkv = a.mode == GEMM_OUT_QKV;
        const int et = threadIdx.x - 64;                 // 0..127 among the epilogue threads (warps 2..5)
        const int m0 = tile_m * a.MT;
        const int c_first = (int)my_rank * cols_per;     // first tile column (token) this CTA owns
        const int c_cnt = max(0, min(cols_per, a.MT - c_first));
        float* const outf = static_cast<float*>(a.out);
        const bool vec = !qkv && ((a.ldo | a.n_valid) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
        const bool emit = a.norm.sumsq_out != nullptr;   // host guarantees vec in this mode
        const bool tpf = a.tp.world > 1;                 // host guarantees vec in this mode
        // fp32 owner: thread = (4 consecutive output columns n, token t = warp, warp + 4, ...)
        const int n4 = et & 31, vn = tile_n * kTileN + 4 * n4;
        const bool n_ok = vn < a.n_valid;
        float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
        float lw[4] = {0.f, 0.f, 0.f, 0.f};
        // QKV owner: thread = (4 consecutive rotary pairs of one head, token t = et / 16, + 8, ...)
        const int hd = qkv ? a.qkv.hd : kTileN, half = hd >> 1;
        const int heads_per_tile = kTileN / hd;
        const int gi = et & 15;
        const int qh2 = hd == 64 ? gi >> 3 : 0, qi = hd == 64 ? (gi & 7) * 4 : gi * 4;   // head in tile, first pair
        const int qhead = tile_n * heads_per_tile + qh2;                                // q heads, then k, then v
        int* const kvrow = reinterpret_cast<int*>(xbuf);                // [cols_per] cache row of each owned token
        float qb1[4] = {0.f, 0.f, 0.f, 0.f}, qb2[4] = {0.f, 0.f, 0.f, 0.f};
        if (warp >= 2) {
            // operands that do not depend on this GEMM are fetched before the cluster barriers
            if (vec) {
                const int t = et >> 5, m = m0 + c_first + t;
                if (n_ok) {
                    if (a.accumulate && t < c_cnt && m < a.M)
                        old = *reinterpret_cast<const float4*>(outf + (size_t)m * a.ldo + vn);
                    if (emit) {
                        const uint2 w4 = *reinterpret_cast<const uint2*>(a.norm.ln_w + vn);
                        const float2 w01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4.x));
                        const float2 w23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4.y));
                        lw[0] = w01.x, lw[1] = w01.y, lw[2] = w23.x, lw[3] = w23.y;
                    }
                }
            } else if (qkv) {
                const QkvEpilogue& e = a.qkv;
                if (qhead < e.nh + 2 * e.nkv) {
                    const uint2 b1 = *reinterpret_cast<const uint2*>(e.bias + qhead * hd + qi);
                    const uint2 b2 = *reinterpret_cast<const uint2*>(e.bias + qhead * hd + qi + half);
                    const float2 a01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&b1.x));
                    const float2 a23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&b1.y));
                    const float2 c01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&b2.x));
                    const float2 c23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&b2.y));
                    qb1[0] = a01.x, qb1[1] = a01.y, qb1[2] = a23.x, qb1[3] = a23.y;
                    qb2[0] = c01.x, qb2[1] = c01.y, qb2[2] = c23.x, qb2[3] = c23.y;
                }
                for (int t = et; t < c_cnt; t += 128) {
                    const int m = m0 + c_first + t;
                    int r = 0;
                    if (m < a.M) {
                        const int pos = e.positions[m];
                        const int page = e.page_table[(size_t)e.token_slot[m] * e.max_pages + pos / e.page_size];
                        r = page * e.nkv * e.page_size + pos % e.page_size;
                    }
                    kvrow[t] = r;
                }
            }
        }
        if (warp < 2) mbar_wait(tmem_full, 0);  // own MMAs done => nothing reads or fills the ring any more
        tc_fence_after();
        uint8_t* recv_base = smem;               // the receive buffer overlays the (idle) tile ring ...
        if (a.recv_dedicated) {                  // ... or, for small token tiles, has its own shared memory
            recv_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xbuf + (qkv ? a.MT : 0)) + 15) & ~uintptr_t(15));
            cluster_wait();
        } else {
            fence_proxy_async_smem();
            cluster_sync();                      // every CTA of the cluster is done with its ring
        }
        if (threadIdx.x == 64) trace_stamp(a, 5);
        // receive layout in the owner: recv[src][owned column][128 rows] - the 32 lanes of a warp (32 consecutive
        // accumulator rows) write one contiguous 128-byte run per column
        if (warp >= 2) {
            const int q = warp & 3, row = q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
            const uint32_t mine = smem_u32(recv_base) + (uint32_t)((int)my_rank * cols_per * kTileN + row) * 4u;
            int owner = 0, lc = 0;               // owner CTA and local column of tile column c0 + i
            for (int c0 = 0; c0 < a.MT; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c0, r);
                tmem_ld_wait();
                uint32_t dst = mapa(mine, (uint32_t)owner);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (c0 + i < a.MT) st_cluster_u32(dst + (uint32_t)(lc * kTileN) * 4u, r[i]);
                    if (++lc == cols_per) {
                        lc = 0;
                        ++owner;
                        dst = mapa(mine, (uint32_t)(owner < ks ? owner : 0));
                    }
                }
            }
            tc_fence_before();
        }
        if (threadIdx.x == 64) trace_stamp(a, 6);
        cluster_sync();
        if (threadIdx.x == 64) trace_stamp(a, 7);
        // every CTA of the cluster has streamed its weights: HBM idles from here to the next kernel's first tile
        if (warp == 1 && lane < 4 && a.next_kp > 0) prefetch_next_weights(a, &tmap_next, lane, 4);
        const float* recv = reinterpret_cast<const float*>(recv_base);
        const float4* recv4 = reinterpret_cast<const float4*>(recv_base);
        if (warp >= 2 && vec) {
            auto cluster_sum = [&](int t) {                  // the ks contributions of one float4, in rank order
                float4 acc = recv4[(t << 5) + n4];
#pragma unroll
                for (int src = 1; src < 8; ++src)
                    if (src < ks) {
                        const float4 v = recv4[((src * cols_per + t) << 5) + n4];
                        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
                    }
                return acc;
            };
            if (tpf) {
                // ---- tensor parallel, step 1: push this rank's partial rows into every peer's receive buffer.
                // Flags travel WITH the data (every 8-byte word = {value, epoch}, 16-byte stores), so there is no
                // system-scope fence and no separate flag round trip: a fence behind 2 MB of NVLink stores costs
                // 8-10 us (measured), the inline flag costs the one-way latency.
                const uint32_t ep = a.tp.epoch;
#pragma unroll 1
                for (int t = et >> 5; t < c_cnt; t += 4) {
                    const int m = m0 + c_first + t;
                    if (!(n_ok && m < a.M)) continue;
                    const float4 acc = cluster_sum(t);
                    const size_t off = 2 * ((size_t)a.tp.rank * a.tp.slot_stride + (size_t)m * a.ldo + vn);
                    for (int p = 0; p < a.tp.world; ++p) {
                        if (p == a.tp.rank) continue;
                        float* d = a.tp.recv[p] + off;
                        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(d), "r"(__float_as_uint(acc.x)),
                                     "r"(ep), "r"(__float_as_uint(acc.y))
                                     : "memory");
                        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(d + 4), "r"(__float_as_uint(acc.z)),
                                     "r"(ep), "r"(__float_as_uint(acc.w))
                                     : "memory");
                    }
                }
                if (threadIdx.x == 64) trace_stamp(a, 13);
            }
            const float* mine = tpf ? a.tp.recv[a.tp.rank] : nullptr;
#pragma unroll 1
            for (int t = et >> 5; t < c_cnt; t += 4) {       // t is warp-uniform
                const int m = m0 + c_first + t;
                float* o = outf + (size_t)m * a.ldo + vn;
                float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.accumulate && n_ok && t + 4 < c_cnt && m + 4 < a.M)
                    nxt = *reinterpret_cast<const float4*>(o + (size_t)4 * a.ldo);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!tpf) {
                    acc = cluster_sum(t);
                } else if (n_ok && m < a.M) {
                    // step 2: the world partials of this float4 in rank order (bit-identical on every rank): my own
                    // from the cluster, the peers' from my receive buffer as soon as their inline flags show up
                    const uint32_t ep = a.tp.epoch;
                    for (int r = 0; r < a.tp.world; ++r) {
                        float4 v;
                        if (r == a.tp.rank) {
                            v = cluster_sum(t);
                        } else {
                            const float* src = mine + 2 * ((size_t)r * a.tp.slot_stride + (size_t)m * a.ldo + vn);
                            uint4 w0, w1;
                            const long long t0 = clock64();
                            do {
                                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(w0.x), "=r"(w0.y), "=r"(w0.z), "=r"(w0.w)
                                             : "l"(src)
                                             : "memory");
                                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(w1.x), "=r"(w1.y), "=r"(w1.z), "=r"(w1.w)
                                             : "l"(src + 4)
                                             : "memory");
                                if (w0.y == ep && w0.w == ep && w1.y == ep && w1.w == ep) break;
                                if (clock64() - t0 > 4000000000LL) {   // ~2 s: a peer died; fail loudly, do not hang
                                    *a.tp.error = 1;
                                    break;
                                }
                            } while (true);
                            v = make_float4(__uint_as_float(w0.x), __uint_as_float(w0.z), __uint_as_float(w1.x),
                                            __uint_as_float(w1.z));
                        }
                        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
                    }
                }
                float sq = 0.0f;
                if (n_ok && m < a.M) {
                    if (a.accumulate) acc.x = old.x + acc.x, acc.y = old.y + acc.y, acc.z = old.z + acc.z, acc.w = old.w + acc.w;
                    *reinterpret_cast<float4*>(o) = acc;
                    if (emit) {
                        const __nv_bfloat162 p01 = __floats2bfloat162_rn(acc.x * lw[0], acc.y * lw[1]);
                        const __nv_bfloat162 p23 = __floats2bfloat162_rn(acc.z * lw[2], acc.w * lw[3]);
                        uint2 pk;
                        pk.x = *reinterpret_cast<const uint32_t*>(&p01);
                        pk.y = *reinterpret_cast<const uint32_t*>(&p23);
                        *reinterpret_cast<uint2*>(a.norm.resid_bf + (size_t)m * a.ldo + vn) = pk;
                        sq = (acc.x * acc.x + acc.y * acc.y) + (acc.z * acc.z + acc.w * acc.w);
                    }
                }
                if (emit) {   // the whole warp works on token m: its sum is this tile's share of the norm statistics
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
                    if (lane == 0 && m < a.M) a.norm.sumsq_out[(size_t)tile_n * a.norm.ld + m] = sq;
                }
                old = nxt;
            }
        } else if (warp >= 2 && !qkv) {
            // generic owner (unaligned output): one element per thread and step
#pragma unroll 1
            for (int idx = et; idx < c_cnt * kTileN; idx += 128) {
                const int t = idx >> 7, rw = idx & (kTileN - 1);
                float acc = recv[idx];
#pragma unroll 1
                for (int src = 1; src < ks; ++src) acc += recv[src * cols_per * kTileN + idx];
                const int n = tile_n * kTileN + rw, m = m0 + c_first + t;
                if (m < a.M && n < a.n_valid) {
                    float* o = outf + (size_t)m * a.ldo + n;
                    *o = a.accumulate ? *o + acc : acc;
                }
            }
        }
        if (threadIdx.x == 64) trace_stamp(a, 10);
        if ((emit || tpf) && !qkv && !vec) __trap();   // the host only asks for these on vectorisable shapes
        if (qkv && warp >= 2) {
            // bias + RoPE + q store / paged K,V append: 4 rotary pairs (dims i..i+3 and i+half..) per thread and step
            const QkvEpilogue& e = a.qkv;
            const bool head_ok = qhead < e.nh + 2 * e.nkv;
            const bool rot = qhead < e.nh + e.nkv;
            const int f1 = (qh2 * hd + qi) >> 2, f2 = (qh2 * hd + half + qi) >> 2;   // float4 index inside a column
            const bool scale = a.norm.sumsq_in != nullptr;
#pragma unroll 1
            for (int t = et >> 4; t < c_cnt; t += 8) {
                const int m = m0 + c_first + t;
                if (m >= a.M || !head_ok) continue;
                float4 c01 = make_float4(1.f, 0.f, 1.f, 0.f), c23 = c01;          // (cos, sin) pairs
                if (rot) {
                    const float4* cs4 = reinterpret_cast<const float4*>(e.cs + (size_t)m * half + qi);
                    c01 = cs4[0];
                    c23 = cs4[1];
                }
                float4 x1 = recv4[(t << 5) + f1], x2 = recv4[(t << 5) + f2];
#pragma unroll
                for (int src = 1; src < 8; ++src)
                    if (src < ks) {
                        const float4 t1 = recv4[((src * cols_per + t) << 5) + f1];
                        const float4 t2 = recv4[((src * cols_per + t) << 5) + f2];
                        x1.x += t1.x, x1.y += t1.y, x1.z += t1.z, x1.w += t1.w;
                        x2.x += t2.x, x2.y += t2.y, x2.z += t2.z, x2.w += t2.w;
                    }
                if (scale) {
                    const float rs = rstd_s[c_first + t];
                    x1.x *= rs, x1.y *= rs, x1.z *= rs, x1.w *= rs;
                    x2.x *= rs, x2.y *= rs, x2.z *= rs, x2.w *= rs;
                }
                x1.x += qb1[0], x1.y += qb1[1], x1.z += qb1[2], x1.w += qb1[3];
                x2.x += qb2[0], x2.y += qb2[1], x2.z += qb2[2], x2.w += qb2[3];
                // o1 = x1 cos - x2 sin, o2 = x2 cos + x1 sin (cos = 1, sin = 0 for the V heads)
                const __nv_bfloat162 lo01 = __floats2bfloat162_rn(x1.x * c01.x - x2.x * c01.y, x1.y * c01.z - x2.y * c01.w);
                const __nv_bfloat162 lo23 = __floats2bfloat162_rn(x1.z * c23.x - x2.z * c23.y, x1.w * c23.z - x2.w * c23.w);
                const __nv_bfloat162 hi01 = __floats2bfloat162_rn(x2.x * c01.x + x1.x * c01.y, x2.y * c01.z + x1.y * c01.w);
                const __nv_bfloat162 hi23 = __floats2bfloat162_rn(x2.z * c23.x + x1.z * c23.y, x2.w * c23.z + x1.w * c23.w);
                __nv_bfloat16* dstp;
                if (qhead < e.nh) {
                    dstp = e.q_out + ((size_t)m * e.nh + qhead) * hd;
                } else {
                    const int kvh = rot ? qhead - e.nh : qhead - e.nh - e.nkv;
                    __nv_bfloat16* cache = rot ? e.k_cache : e.v_cache;
                    dstp = cache + (size_t)(kvrow[t] + kvh * e.page_size) * hd;
                }
                uint2 lo, hi;
                lo.x = *reinterpret_cast<const uint32_t*>(&lo01);
                lo.y = *reinterpret_cast<const uint32_t*>(&lo23);
                hi.x = *reinterpret_cast<const uint32_t*>(&hi01);
                hi.y = *reinterpret_cast<const uint32_t*>(&hi23);
                *reinterpret_cast<uint2*>(dstp + qi) = lo;
                *reinterpret_cast<uint2*>(dstp + qi + half) = hi;
            }
            if (threadIdx.x == 64) trace_stamp(a, 11);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) trace_stamp(a, 8);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int load_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    ASD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) return set_error("cuTensorMapEncodeTiled unavailable");
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    return 0;
}

// bf16 row-major [rows, cols] matrix, box = {64 cols, box_rows}, 128-byte swizzle, zero OOB fill
int make_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                   uint32_t box_rows) {
    if (load_encode()) return -1;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (row_stride_elems * 2) % 16)
        return set_error("tensor map: base and row stride must be 16-byte aligned");
    if (box_rows == 0 || box_rows > 256) return set_error("tensor map: box rows must be in [1, 256]");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

thread_local NextPrefetch g_gemm_next;   // set by the caller right before gemm_launch, consumed by it (per host thread)
unsigned long long* g_gemm_trace = nullptr;   // diagnostics buffer [max launches][kTraceCtas][kTraceSlots]
int g_gemm_trace_max = 0, g_gemm_trace_next = 0;
constexpr int kTraceCtas = 1024;   // the caller marked it (the GEMM that follows the latency-bound attention kernel)
static int g_num_sms = 0;
static int g_smem_optin = 0;
static PerDeviceOnce g_gemm_attr;

static int device_props() {
    if (g_num_sms) return 0;
    int dev = 0;
    ASD_CUDA(cudaGetDevice(&dev));
    ASD_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    ASD_CUDA(cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return 0;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

int gemm_token_tile(int M) {
    if (M <= 256) return round_up(M < 16 ? 16 : M, 16);
    const int tiles = (M + 255) / 256;
    return round_up((M + tiles - 1) / tiles, 16);
}

// Choose the split so that all CTAs fit in one co-resident wave and there are enough bytes in flight.
int gemm_plan(GemmPlan* pl, int M, int N, int K, int mode, int force_ksplit, int force_stages, int reduce) {
    if (device_props()) return -1;
    pl->reduce = ((reduce && mode == GEMM_OUT_F32) || mode == GEMM_OUT_QKV) ? 1 : 0;
    if (M <= 0 || N <= 0 || K <= 0 || (K & 7)) return set_error("gemm: need M, N, K > 0 and K %% 8 == 0");
    pl->M = M;
    pl->N = N;
    pl->K = K;
    pl->mode = mode;
    pl->MT = gemm_token_tile(M);
    pl->m_tiles = (M + pl->MT - 1) / pl->MT;
    pl->n_tiles = (N + kTileN - 1) / kTileN;
    pl->kblocks = (K + kBlockK - 1) / kBlockK;
    const int stage_bytes = kABytes + pl->MT * 128;
    const int fixed = 1024 /*align*/ + 256 /*barriers*/ + pl->MT * 4 /*rstd*/ +
                      (mode == GEMM_OUT_QKV ? pl->MT * 4 : 0);   // QKV: cache row of each owned token
    const int max_ctas_per_sm = pl->MT <= 128 ? 3 : 2;   // TMEM: 128 / 256 columns per CTA
    const int tiles = pl->n_tiles * pl->m_tiles;
    // Split choice.  Measured on B200: one SM cannot ingest more than ~40 GB/s from HBM however deep its
    // pipeline is, so a weight stream reaches HBM speed only if all 148 SMs carry the same number of
    // bytes.  Pick the split that minimises the heaviest SM's load  ceil(CTAs / SMs) * kblocks / ksplit
    // while every CTA stays co-resident (one wave) and keeps >= 4 k-blocks of work.
    int ksplit = 1;
    if (mode != GEMM_OUT_SWIGLU && mode != GEMM_OUT_BF16) {
        // largest split that keeps every CTA co-resident in one wave with >= 4 k-blocks each; with the
        // in-cluster reduction it is a power of two <= 8 (cluster sizes that tile the GPCs evenly)
        const int slots = g_num_sms * max_ctas_per_sm;
        ksplit = slots / tiles;
        if (ksplit < 1) ksplit = 1;
        const int max_by_k = pl->kblocks / 4 > 0 ? pl->kblocks / 4 : 1;
        if (ksplit > max_by_k) ksplit = max_by_k;
        if (ksplit > (pl->reduce ? 8 : 16)) ksplit = pl->reduce ? 8 : 16;
        if (pl->reduce) {
            int p2 = 1;
            while (p2 * 2 <= ksplit) p2 *= 2;
            ksplit = p2;
        }
    }
    if (force_ksplit > 0) ksplit = force_ksplit;
    // pipeline depth from the expected residency: fewer CTAs per SM -> deeper ring per CTA
    int ctas_per_sm = (tiles * ksplit + g_num_sms - 1) / g_num_sms;
    if (ctas_per_sm > max_ctas_per_sm) ctas_per_sm = max_ctas_per_sm;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    // headroom: leave shared memory for that many CTAs of the NEXT kernel (programmatic launch + early trigger)
    // (measured: pays for the draft's 16-token tiles, whose rings stay >= 4 deep; costs the verify tiles a stage)
    const int headroom = pl->MT <= 32 ? tuning().gemm_headroom : 0;
    const int budget = (225 * 1024) / (ctas_per_sm + headroom) - fixed - 1024;
    int stages = budget / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
    if (force_stages > 0) stages = force_stages;
    pl->recv_dedicated = 0;
    int recv_extra = 0;
    if (pl->reduce && (ksplit > 1 || mode == GEMM_OUT_QKV)) {
        if (ksplit > 8) return set_error("gemm: cluster reduction supports ksplit <= 8");
        const int cols_per = (pl->MT + ksplit - 1) / ksplit;
        const int recv = ksplit * cols_per * kTileN * 4;
        if (pl->MT <= 32 && tuning().gemm_recv_dedicated) {
            pl->recv_dedicated = 1;                     // small tiles: own receive buffer, one cluster barrier less
            recv_extra = recv + 16;
            while (stages > 2 && fixed + recv_extra + stages * stage_bytes > budget + fixed) --stages;
        } else {
            while (stages * stage_bytes < recv) ++stages;   // the receive buffer overlays the tile ring
        }
    }
    if (mode == GEMM_OUT_SWIGLU)
        while (stages * stage_bytes < pl->MT * kTileN * 4) ++stages;   // the SwiGLU transpose overlays the tile ring
    if ((mode == GEMM_OUT_SWIGLU || mode == GEMM_OUT_BF16) && ksplit != 1)
        return set_error("gemm: bf16 / SwiGLU epilogues need ksplit == 1");
    if (ksplit > pl->kblocks) ksplit = pl->kblocks;
    pl->ksplit = ksplit;
    pl->stages = stages;
    if (ksplit == 1 && mode != GEMM_OUT_QKV) pl->reduce = 0;
    pl->smem_bytes = fixed + recv_extra + stages * stage_bytes;
    if (pl->smem_bytes > g_smem_optin) return set_error("gemm: tile does not fit in shared memory");
    uint32_t cols = 32;
    while ((int)cols < pl->MT) cols <<= 1;
    pl->tmem_cols = cols;
    return 0;
}

void gemm_set_next(const GemmPlan& next, const CUtensorMap* next_w) {
    g_gemm_next = NextPrefetch{};
    if (tuning().gemm_next_mb <= 0 || next_w == nullptr || next.m_tiles != 1) return;
    const long long ncta = (long long)next.n_tiles * next.ksplit;
    int kp = (int)(((long long)tuning().gemm_next_mb << 20) / (ncta * kABytes));
    const int per_cta = (next.kblocks + next.ksplit - 1) / next.ksplit;
    if (kp > per_cta) kp = per_cta;
    if (kp <= 0) return;
    g_gemm_next.tmap = next_w;
    g_gemm_next.ntiles = next.n_tiles;
    g_gemm_next.ksplit = next.ksplit;
    g_gemm_next.kblocks = next.kblocks;
    g_gemm_next.kp = kp;
}

int gemm_launch(const GemmPlan& pl, const CUtensorMap& tmap_w, const CUtensorMap& tmap_x, void* out, int ldo,
                int n_valid, bool pdl, cudaStream_t stream, bool accumulate, const QkvEpilogue* qkv,
                const NormFusion* norm, const TpFusion* tp) {
    if (g_gemm_attr.need()) {
        if (device_props()) return -1;
        ASD_CUDA(cudaFuncSetAttribute(gemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
        prefer_max_smem(gemm_ws_kernel);
    }
    GemmArgs a;
    a.M = pl.M;
    a.N = pl.N;
    a.K = pl.K;
    a.MT = pl.MT;
    a.kblocks = pl.kblocks;
    a.ksplit = pl.ksplit;
    a.stages = pl.stages;
    a.mode = pl.mode;
    a.ldo = ldo;
    a.n_valid = n_valid;
    a.out = out;
    a.tmem_cols = pl.tmem_cols;
    a.reduce = pl.reduce;
    a.recv_dedicated = pl.recv_dedicated;
    a.accumulate = accumulate ? 1 : 0;
    a.early_trigger = tuning().gemm_early_trigger;
    a.norm = norm ? *norm : NormFusion{};
    a.tp = TpFusion{};
    a.next_ntiles = a.next_ksplit = a.next_kblocks = a.next_kp = 0;
    const CUtensorMap* tnext = &tmap_w;
    if (g_gemm_next.kp > 0 && g_gemm_next.tmap != nullptr) {
        a.next_ntiles = g_gemm_next.ntiles;
        a.next_ksplit = g_gemm_next.ksplit;
        a.next_kblocks = g_gemm_next.kblocks;
        a.next_kp = g_gemm_next.kp;
        tnext = g_gemm_next.tmap;
    }
    g_gemm_next = NextPrefetch{};
    if (tp && tp->world > 1) {
        if (!pl.reduce || pl.mode != GEMM_OUT_F32 || ((ldo | n_valid) & 3) || (reinterpret_cast<uintptr_t>(out) & 15))
            return set_error("gemm: the fused all-reduce needs the cluster reduction on 16-byte aligned rows");
        if ((size_t)pl.M * ldo > tp->slot_stride) return set_error("gemm: receive slot smaller than the output");
        a.tp = *tp;
    }
    a.trace = nullptr;
    if (g_gemm_trace && g_gemm_trace_next < g_gemm_trace_max) {
        a.trace = g_gemm_trace + (size_t)g_gemm_trace_next * kTraceCtas * kTraceSlots;
        if (pl.n_tiles * pl.ksplit * pl.m_tiles > kTraceCtas) a.trace = nullptr;
        ++g_gemm_trace_next;
    }
    if (a.norm.sumsq_out != nullptr) {
        if (!pl.reduce || pl.mode != GEMM_OUT_F32)
            return set_error("gemm: the sum-of-squares epilogue needs the cluster reduction");
        if (!a.norm.resid_bf || !a.norm.ln_w) return set_error("gemm: fused norm producer needs resid_bf and ln_w");
        if (((ldo | n_valid) & 3) || (reinterpret_cast<uintptr_t>(out) & 15))
            return set_error("gemm: fused norm producer needs 16-byte aligned rows");
    }
    a.qkv = QkvEpilogue{};
    if (pl.mode == GEMM_OUT_QKV) {
        if (!qkv) return set_error("gemm: QKV epilogue needs its operands");
        a.qkv = *qkv;
        if (qkv->hd != 64 && qkv->hd != 128) return set_error("gemm: QKV epilogue shape");
    }
    if (accumulate && pl.mode != GEMM_OUT_F32) return set_error("gemm: accumulate needs the fp32 epilogue");
    if (accumulate && pl.ksplit > 1 && !pl.reduce) return set_error("gemm: accumulate needs the cluster reduction");
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    cfg.gridDim = dim3(pl.n_tiles, pl.ksplit, pl.m_tiles);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = stream;
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (pl.reduce) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1;
        attr[na].val.clusterDim.y = pl.ksplit;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    ASD_CUDA(cudaLaunchKernelEx(&cfg, gemm_ws_kernel, tmap_w, tmap_x, *tnext, a));
    count_launch(1);
    return 0;
}

// Force the (lazily loaded) kernels of this file into the context now: a first launch that loads a kernel may need a
// context synchronisation, which deadlocks when another rank of the same process is spinning for this rank's launch.
int preload_gemm() {
    cudaFuncAttributes fa;
    ASD_CUDA(cudaFuncGetAttributes(&fa, gemm_ws_kernel));
    return 0;
}

}  // namespace asd

// ------------------------------------------------------------------------------------------- C ABI
#include "../../include/asd_b200.h"
extern "C" int asd_linear_bf16(const void* x, const void* w, void* out, int M, int N, int K, int out_mode,
                               int ksplit, int stages, int* ksplit_used, void* stream) {
    using namespace asd;
    GemmPlan pl;
    if (out_mode < 0 || out_mode > 3) return set_error("asd_linear_bf16: bad out_mode");
    const int reduce = out_mode == 3;
    if (reduce) out_mode = GEMM_OUT_F32;
    if (gemm_plan(&pl, M, N, K, out_mode, ksplit, stages, reduce)) return -1;
    CUtensorMap tw, tx;
    if (make_tmap_bf16(&tw, w, N, K, K, 128)) return -1;
    if (make_tmap_bf16(&tx, x, M, K, K, pl.MT)) return -1;
    if (ksplit_used) *ksplit_used = pl.ksplit;
    const int ldo = out_mode == GEMM_OUT_SWIGLU ? N / 2 : N;
    return gemm_launch(pl, tw, tx, out, ldo, ldo, false, static_cast<cudaStream_t>(stream), false);
}
extern "C" int asd_linear_plan(int M, int N, int K, int out_mode, int* ksplit, int* stages, int* token_tile) {
    using namespace asd;
    GemmPlan pl;
    if (out_mode < 0 || out_mode > 3) return set_error("asd_linear_plan: bad out_mode");
    if (gemm_plan(&pl, M, N, K, out_mode == 3 ? GEMM_OUT_F32 : out_mode, 0, 0, out_mode == 3)) return -1;
    if (ksplit) *ksplit = pl.ksplit;
    if (stages) *stages = pl.stages;
    if (token_tile) *token_tile = pl.MT;
    return 0;
}
extern "C" ASD_API int asd_debug_gemm_trace(unsigned long long* buf, int max_launches) {
    asd::g_gemm_trace = buf;
    asd::g_gemm_trace_max = buf ? max_launches : 0;
    asd::g_gemm_trace_next = 0;
    return asd::kTraceCtas * asd::kTraceSlots;
}
