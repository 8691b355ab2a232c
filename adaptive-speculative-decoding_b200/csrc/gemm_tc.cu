// Tensor-bound bf16 GEMM for the large verify steps:  Y[m, n] = sum_k X[m, k] * W[n, k]  with M above the
// HBM/tensor ridge (~211 tokens on B200; BASELINE configs[4]: 72B verify of 64 x 9 = 576 tokens, and prefill).
//
// The weight-streaming kernel (gemm.cu) gives every CTA one 128-row weight tile and ALL its k-blocks once; with
// several token tiles it re-reads the weights per token tile and its 128 x MT tile moves 2 * (128 + MT) bytes per
// 128 * MT MACs through L2.  Here the tile is 256 weight rows x MT tokens on a CTA PAIR:
//   * tcgen05.mma.cta_group::2, UMMA M = 256 (128 weight rows per CTA = the A operand), N = MT <= 256 tokens (the B
//     operand, N-split: each CTA loads MT/2 token rows), K = 16; one elected thread of the leader CTA issues for
//     both SMs, accumulators [128 lanes x MT columns] fp32 live in each CTA's TMEM;
//   * every CTA therefore stages 16 KB + MT/2 * 128 B per 64-wide k-block for 128 * MT * 64 MACs - half the
//     shared-memory fill per MAC of a single-CTA tile;
//   * TMA (cta_group::2) of both CTAs completes on the LEADER's mbarrier; tcgen05.commit multicasts "slot free" /
//     "accumulator ready" to both CTAs;
//   * persistent: one pair per SM pair, static round-robin over (weight tile, K split, token tile) with the token
//     tile fastest, so neighbouring pairs stream the SAME weight tile at the same time and HBM sees it once
//     (the others hit L2); TMEM holds two accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * K splits write fp32 slices that the glue kernels (add_norm / reduce_slices) sum in order - deterministic.
// Epilogues: fp32 (+ per-token rstd of the fused RMSNorm, + accumulate) and SwiGLU -> bf16 for gate|up weights
// interleaved 64 gate / 64 up rows per 128-row tile (same layout as gemm.cu).
//
// Stands behind Stage.generate's model forward, which the reference delegates to vLLM
// (/root/reference/src/serving/real_model_pipeline.py:98-108,135).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "asd_internal.h"
#include "gemm.h"
#include "ptx.cuh"

namespace asd {

constexpr int kTcThreads = 256;          // warp 0 TMA, warp 1 MMA (leader), warp 2 TMEM, warps 4..7 epilogue
constexpr int kTcBlockK = 64;
constexpr int kTcABytes = 128 * kTcBlockK * 2;
constexpr int kTcAccCols = 256;          // TMEM column stride between the two accumulators
constexpr int kTcMaxM = 4096;            // per-token rstd table in shared memory

struct TcArgs {
    int M, N, K;
    int MT, m_tiles, w_tiles, kblocks, ksplit, stages, mode;
    int ldo, n_valid, accumulate;
    void* out;
    size_t slice_stride;   // elements between the K-split slices of the fp32 output
    NormFusion norm;       // consumer side only (sumsq_in / parts / ld / hidden / eps)
};

// ---- cta_group::2 flavours of the PTX wrappers
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are credited to an mbarrier given as a shared::cluster
// address (the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_cg2(void* dst_smem, const void* tmap, int c0, int c1, uint32_t bar_cluster_addr,
                                                uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst_smem)),
        "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in BOTH CTAs of the pair once all tcgen05.mma issued so
// far by this thread have completed
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ float tc_silu_mul(float g, float u) { return __fdividef(g, 1.0f + __expf(-g)) * u; }

struct TcUnit {
    int w_tile, ks, t_tile, kb0, nkb;
};
__device__ __forceinline__ TcUnit tc_unit(const TcArgs& a, int u) {
    TcUnit r;
    r.t_tile = u % a.m_tiles;
    const int v = u / a.m_tiles;
    r.ks = v % a.ksplit;
    r.w_tile = v / a.ksplit;
    const int base = a.kblocks / a.ksplit, rem = a.kblocks % a.ksplit;
    r.kb0 = r.ks * base + (r.ks < rem ? r.ks : rem);
    r.nkb = base + (r.ks < rem ? 1 : 0);
    return r;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = (a.MT / 2) * 128;
    const int stage_bytes = kTcABytes + b_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + a.stages;
    uint64_t* tfull_bar = empty_bar + a.stages;   // [2] accumulator ready (per CTA)
    uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained (leader's copy is the one waited on)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* rstd_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));   // [M]
    float* stage_t = rstd_s + ((a.norm.sumsq_in != nullptr ? a.M : 0) + 3) / 4 * 4;   // SwiGLU exchange [2][32][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int units = a.w_tiles * a.ksplit * a.m_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 8);    // 4 epilogue warps x 2 CTAs
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_cg2(tmem_slot, 512);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync();      // both CTAs' barriers are initialised before anything arrives remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            // weights are shared by the m_tiles pairs that stream the same tile: keep them in L2 for those
            const uint64_t pol_w = a.m_tiles > 1 ? policy_evict_normal() : policy_evict_first();
            const uint64_t pol_x = policy_evict_last();
            const uint32_t full_leader = mapa(smem_u32(full_bar), 0);
            bool waited = false;
            int it = 0;
            for (int u = pair; u < units; u += npairs) {
                const TcUnit t = tc_unit(a, u);
                for (int kb = 0; kb < t.nkb; ++kb, ++it) {
                    const int s = it % a.stages, round = it / a.stages;
                    if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
                    if (leader) mbar_expect_tx(&full_bar[s], 2 * stage_bytes);
                    uint8_t* st = smem + (size_t)s * stage_bytes;
                    const uint32_t bar = full_leader + (uint32_t)s * 8u;
                    tma_load_2d_cg2(st, &tmap_w, (t.kb0 + kb) * kTcBlockK, t.w_tile * 256 + (int)rank * 128, bar, pol_w);
                    if (!waited) {   // weights do not depend on the upstream kernel, activations do
                        grid_dep_wait();
                        waited = true;
                    }
                    tma_load_2d_cg2(st + kTcABytes, &tmap_x, (t.kb0 + kb) * kTcBlockK,
                                    t.t_tile * a.MT + (int)rank * (a.MT / 2), bar, pol_x);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader && lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(256, a.MT);
            int it = 0, tile = 0;
            for (int u = pair; u < units; u += npairs, ++tile) {
                const TcUnit t = tc_unit(a, u);
                const int acc = tile & 1, use = tile >> 1;
                if (use > 0) mbar_wait(&tempty_bar[acc], (use - 1) & 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * kTcAccCols);
                for (int kb = 0; kb < t.nkb; ++kb, ++it) {
                    const int s = it % a.stages, round = it / a.stages;
                    mbar_wait(&full_bar[s], round & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + kTcABytes);
#pragma unroll
                    for (int k = 0; k < kTcBlockK / 16; ++k) umma_f16_cg2(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_cg2(&empty_bar[s]);
                }
                umma_commit_cg2(&tfull_bar[acc]);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (both CTAs)
        const int q = warp & 3, row = q * 32 + lane, et = threadIdx.x - 128;
        const bool scale = a.norm.sumsq_in != nullptr;
        grid_dep_wait();
        if (scale) {
            for (int m = et; m < a.M; m += 128) {
                float ssum = 0.0f;
                for (int t = 0; t < a.norm.parts; ++t) ssum += a.norm.sumsq_in[(size_t)t * a.norm.ld + m];
                rstd_s[m] = rsqrtf(ssum / (float)a.norm.hidden + a.norm.eps);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        const uint32_t tempty_leader = mapa(smem_u32(tempty_bar), 0);
        int tile = 0;
        for (int u = pair; u < units; u += npairs, ++tile) {
            const TcUnit t = tc_unit(a, u);
            const int acc = tile & 1, use = tile >> 1;
            mbar_wait(&tfull_bar[acc], use & 1);
            tc_fence_after();
            if (tile == 0) grid_dep_launch();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kTcAccCols);
            const int m0 = t.t_tile * a.MT;
            const int tile128 = t.w_tile * 2 + (int)rank;          // 128-row weight tile this CTA holds
            if (a.mode == GEMM_OUT_SWIGLU) {
                __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a.out);
                const bool wide = (a.ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
#pragma unroll 1
                for (int c0 = 0; c0 < a.MT; c0 += 32) {
                    float* T = stage_t + ((c0 >> 5) & 1) * (32 * 128);
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) T[i * 128 + row] = __uint_as_float(r[i]);
                    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
                    for (int g = et; g < 32 * 8; g += 128) {
                        const int col = g >> 3, f8 = (g & 7) * 8, c = c0 + col;
                        const int m = m0 + c, j0 = tile128 * 64 + f8;
                        if (c >= a.MT || m >= a.M || j0 >= a.n_valid) continue;
                        const float rs = scale ? rstd_s[m] : 1.0f;
                        const float4* tp = reinterpret_cast<const float4*>(T + col * 128 + f8);
                        const float4 g0 = tp[0], g1 = tp[1], u0 = tp[16], u1 = tp[17];
                        float v[8];
                        v[0] = tc_silu_mul(g0.x * rs, u0.x * rs), v[1] = tc_silu_mul(g0.y * rs, u0.y * rs);
                        v[2] = tc_silu_mul(g0.z * rs, u0.z * rs), v[3] = tc_silu_mul(g0.w * rs, u0.w * rs);
                        v[4] = tc_silu_mul(g1.x * rs, u1.x * rs), v[5] = tc_silu_mul(g1.y * rs, u1.y * rs);
                        v[6] = tc_silu_mul(g1.z * rs, u1.z * rs), v[7] = tc_silu_mul(g1.w * rs, u1.w * rs);
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
                }
            } else {
                float* out = static_cast<float*>(a.out) + (size_t)t.ks * a.slice_stride;
                const int n = tile128 * 128 + row;
#pragma unroll 1
                for (int c0 = 0; c0 < a.MT; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + c0, r);
                    tmem_ld_wait();
                    if (n < a.n_valid) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int m = m0 + c0 + i;
                            if (c0 + i < a.MT && m < a.M) {
                                float* o = out + (size_t)m * a.ldo + n;
                                const float v = scale ? __uint_as_float(r[i]) * rstd_s[m] : __uint_as_float(r[i]);
                                *o = a.accumulate ? *o + v : v;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader + (uint32_t)acc * 8u);
        }
    }
    tc_fence_before();
    cluster_sync();      // the leader's MMAs read the peer's shared memory and both CTAs' TMEM until here
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------- host
static PerDeviceOnce g_tc_attr;
static int g_tc_sms = 0, g_tc_smem_optin = 0;

int gemm_tc_plan(TcPlan* pl, int M, int N, int K, int mode, int force_ksplit, int force_stages) {
    if (!g_tc_sms) {
        int dev = 0;
        ASD_CUDA(cudaGetDevice(&dev));
        ASD_CUDA(cudaDeviceGetAttribute(&g_tc_sms, cudaDevAttrMultiProcessorCount, dev));
        ASD_CUDA(cudaDeviceGetAttribute(&g_tc_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    }
    if (M <= 0 || N <= 0 || K <= 0 || (K & 7)) return set_error("gemm_tc: need M, N, K > 0 and K %% 8 == 0");
    if (mode != GEMM_OUT_F32 && mode != GEMM_OUT_SWIGLU) return set_error("gemm_tc: fp32 and SwiGLU epilogues only");
    if (M > kTcMaxM) return set_error("gemm_tc: at most %d tokens per call", kTcMaxM);
    pl->M = M;
    pl->N = N;
    pl->K = K;
    pl->mode = mode;
    pl->w_tiles = (N + 255) / 256;
    pl->kblocks = (K + kTcBlockK - 1) / kTcBlockK;
    const int npairs = g_tc_sms / 2;
    static const int env_mt = []() {
        const char* v = getenv("ASD_TC_MT");     // experiments: force the token tile
        return v ? atoi(v) : 0;
    }();
    // Tile / split choice.  Measured on B200 (tools/bench_gemm_tc.py): an SM takes in ~58 bytes per clock from L2,
    // so a k-block of a 256 x MT pair tile costs max(2 * MT clocks of tensor time, (16 KB + 64 * MT) / 58 clocks of
    // staging) - MT = 192 runs at 0.81 of the cuBLAS rate, MT = 96 at 0.48 - and what remains is wave quantisation
    // (e.g. 96 tiles on 74 pairs = 65 %).  Enumerate token tilings (1..8 equal tiles, MT <= 256) and K splits
    // (fp32 epilogue only, >= 8 k-blocks per unit, each extra slice charged its write + read) and keep the cheapest
    // waves * k-blocks * clocks-per-k-block.
    double best_cost = 1e300;
    int best_mt = 0, best_ks = 1;
    for (int nt = 1; nt <= 16; ++nt) {
        int mt = ((M + nt - 1) / nt + 15) / 16 * 16;
        if (mt < 32) mt = 32;
        if (mt > 256) continue;
        if (env_mt >= 32 && nt != (M + env_mt - 1) / env_mt) continue;
        const int m_tiles = (M + mt - 1) / mt;
        const double clk = std::max(2.0 * mt, (16384.0 + 64.0 * mt) / 58.0);
        for (int ks = 1; ks <= (mode == GEMM_OUT_F32 ? 8 : 1); ++ks) {
            if (ks > 1 && pl->kblocks / ks < 8) break;
            if (force_ksplit > 0 && ks != std::min(force_ksplit, pl->kblocks)) continue;
            const int units = pl->w_tiles * m_tiles * ks;
            const int waves = (units + npairs - 1) / npairs;
            const double kb_unit = (double)((pl->kblocks + ks - 1) / ks);
            // per unit: prologue/epilogue hand-over ~1200 clocks; a slice of 256 x mt fp32 costs ~ mt * 18 clocks
            const double cost = waves * (kb_unit * clk + 1200.0 + (ks > 1 ? mt * 18.0 : 0.0));
            if (cost < best_cost) {
                best_cost = cost;
                best_mt = mt;
                best_ks = ks;
            }
        }
        if (mt == 32) break;
    }
    if (best_mt == 0) return set_error("gemm_tc: no tiling for M = %d", M);
    const int mt = best_mt;
    pl->MT = mt;
    pl->m_tiles = (M + mt - 1) / mt;
    int best = best_ks;
    if (force_ksplit > 0) best = force_ksplit;
    if (best > pl->kblocks) best = pl->kblocks;
    if (mode == GEMM_OUT_SWIGLU) best = 1;
    pl->ksplit = best;
    const int stage_bytes = kTcABytes + (mt / 2) * 128;
    const int fixed = 1024 + 256 + 4 * kTcMaxM + (mode == GEMM_OUT_SWIGLU ? 2 * 32 * 128 * 4 : 0) + 64;
    int stages = (g_tc_smem_optin - fixed) / stage_bytes;
    if (stages > 8) stages = 8;
    if (force_stages > 0 && force_stages < stages) stages = force_stages;
    if (stages < 2) return set_error("gemm_tc: tile does not fit in shared memory");
    pl->stages = stages;
    pl->smem_bytes = fixed + stages * stage_bytes;
    int units = pl->w_tiles * pl->m_tiles * pl->ksplit;
    pl->pairs = units < npairs ? units : npairs;
    return 0;
}

int gemm_tc_launch(const TcPlan& pl, const CUtensorMap& tmap_w, const CUtensorMap& tmap_x, void* out, int ldo, int n_valid,
                   size_t slice_stride, bool pdl, cudaStream_t stream, bool accumulate, const NormFusion* norm) {
    if (g_tc_attr.need()) {
        ASD_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_tc_smem_optin));
        prefer_max_smem(gemm_tc_kernel);
    }
    if (accumulate && (pl.mode != GEMM_OUT_F32 || pl.ksplit != 1))
        return set_error("gemm_tc: accumulate needs the fp32 epilogue without a K split");
    TcArgs a;
    a.M = pl.M;
    a.N = pl.N;
    a.K = pl.K;
    a.MT = pl.MT;
    a.m_tiles = pl.m_tiles;
    a.w_tiles = pl.w_tiles;
    a.kblocks = pl.kblocks;
    a.ksplit = pl.ksplit;
    a.stages = pl.stages;
    a.mode = pl.mode;
    a.ldo = ldo;
    a.n_valid = n_valid;
    a.accumulate = accumulate ? 1 : 0;
    a.out = out;
    a.slice_stride = slice_stride;
    a.norm = norm ? *norm : NormFusion{};
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(2 * pl.pairs);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = stream;
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    ASD_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel, tmap_w, tmap_x, a));
    count_launch(1);
    return 0;
}

// Force the (lazily loaded) kernels of this file into the context now: a first launch that loads a kernel may need a
// context synchronisation, which deadlocks when another rank of the same process is spinning for this rank's launch.
int preload_gemm_tc() {
    cudaFuncAttributes fa;
    ASD_CUDA(cudaFuncGetAttributes(&fa, gemm_tc_kernel));
    return 0;
}

}  // namespace asd

#include "../../include/asd_b200.h"
extern "C" int asd_linear_bf16_tc(const void* x, const void* w, void* out, int M, int N, int K, int out_mode, int ksplit,
                                  int stages, int* ksplit_used, void* stream) {
    using namespace asd;
    if (out_mode != 0 && out_mode != 2) return set_error("asd_linear_bf16_tc: out_mode 0 (fp32 slices) or 2 (SwiGLU)");
    TcPlan pl;
    if (gemm_tc_plan(&pl, M, N, K, out_mode == 2 ? GEMM_OUT_SWIGLU : GEMM_OUT_F32, ksplit, stages)) return -1;
    CUtensorMap tw, tx;
    if (make_tmap_bf16(&tw, w, N, K, K, 128)) return -1;
    if (make_tmap_bf16(&tx, x, M, K, K, pl.MT / 2)) return -1;
    if (ksplit_used) *ksplit_used = pl.ksplit;
    const int ldo = out_mode == 2 ? N / 2 : N;
    return gemm_tc_launch(pl, tw, tx, out, ldo, ldo, (size_t)M * ldo, false, static_cast<cudaStream_t>(stream), false,
                          nullptr);
}
