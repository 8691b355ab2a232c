// Persistent forward kernel: the whole draft / verify forward of a Qwen2 model in ONE cooperative launch.
//
// Why: with one launch per GEMM (gemm.cu) the weight stream stops at every kernel boundary - cluster barrier, DSMEM
// scatter, owner epilogue, exit, the next grid's entry and first activation tile: 8-12 us of idle HBM per boundary,
// five boundaries per layer (25-45 % of a layer at 16-96 tokens).  Here 148 CTAs (one per SM) stay
// resident for the whole forward and the weight stream never stops for a boundary:
//   * warp 0 (one thread) is the WEIGHT producer: it walks the static schedule of every GEMM of the forward
//     (QKV, O, gate|up, down of each layer, lm_head) and keeps TMA loads of 128 x 64 weight blocks in flight into a
//     shared-memory ring.  Weights depend on nothing, so it runs ahead across phase boundaries as far as the ring
//     allows: while the epilogue / dependency chain of phase p resolves, the blocks of phase p+1 are already landing;
//   * warp 1 (one thread) is the ACTIVATION producer: per phase it waits for the producing phase's completion counter
//     (ld.acquire.gpu spin, bounded) and then TMA-loads the X blocks that pair with the weight blocks (same ring stage,
//     same mbarrier, two expect_tx arrivals);
//   * warp 2 (one thread) issues tcgen05.mma (swap-AB as in gemm.cu: weight block = A, M = 128; token tile = B,
//     N = tokens rounded to 16) into one of two TMEM accumulators, so the epilogue of a segment overlaps the MMAs of
//     the next;
//   * warps 4-7 drain accumulators and run everything that is not a GEMM main loop: the fused epilogues (RMSNorm
//     scale, bias + RoPE + q store + paged K/V append, residual + norm statistics, SwiGLU, logits), the embedding
//     gather, the paged-KV attention (mma.sync, same arithmetic as attention.cu) and the logit-row gather.
// Work split ("stream-K"): the (weight tile, k-block) space of a GEMM is cut into 148 equal contiguous ranges, so
// every SM streams the same number of bytes whatever the shape (no wave quantisation, no split heuristics).  A range
// covers a tail piece of one tile, some whole tiles and a head piece of another; a CTA that does not hold a tile's
// first k-block writes its fp32 partial to an L2-resident slot and raises a flag, the CTA that holds the first
// k-block adds the partials in k order (deterministic) and runs the tile's epilogue.  Tail pieces come first in a
// CTA's range, so their partials are long done when the finisher needs them.
// Phases are ordered by per-phase completion counters in global memory (every CTA adds one when its share of the
// phase is done); there is no grid-wide barrier instruction and a CTA never waits for a CTA that waits for it.
// The launch is cooperative (co-residency is checked by the driver: it fails loudly instead of deadlocking); every
// global spin is bounded and reports through PersistLaunch::error.
//
// Stands behind Stage.generate's model forward, which the reference delegates to vLLM
// (/root/reference/src/serving/real_model_pipeline.py:98-108,135).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "asd_internal.h"
#include "gemm.h"
#include "mma_sync.cuh"
#include "persist.h"
#include "ptx.cuh"

namespace asd {

constexpr int kPThreads = 256;
constexpr int kPTile = 128;            // weight rows per block (UMMA M)
constexpr int kPBlockK = 64;           // bf16 per k-block = one 128-byte swizzle row
constexpr int kPWBytes = kPTile * kPBlockK * 2;
constexpr int kPKeyTile = 64;
constexpr long long kPSpinLimit = 6000000000LL;   // ~3 s of clock64 ticks: a wait this long means a lost CTA

enum PMode { PM_QKV = 0, PM_RESID = 1, PM_SWIGLU = 2, PM_LOGITS = 3 };

struct PArgs {
    PersistLaunch L;
    int MT, M_lm, MT_lm, MTmax;
    int stages, stage_bytes, aux_bytes, ahead;
    int rg_count, kg_count, page_shift;
    float scale_log2;
    int n_phases;
    unsigned* counters;     // [n_phases]
    unsigned* part_flag;    // [gridDim.x]
};

// ---------------------------------------------------------------------------------------------- small device helpers
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// bounded spin until *p >= target (phase counters) - sets *err and gives up after ~3 s so that a lost CTA turns into
// a reported error instead of a hung GPU
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned target, int* err) {
    if (ld_acquire_u32(p) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_u32(p) < target) {
        __nanosleep(20);
        if (clock64() - t0 > kPSpinLimit) {
            *err = 1;
            break;
        }
    }
}
__device__ __forceinline__ void wait_flag_eq(const unsigned* p, unsigned value, int* err) {
    if (ld_acquire_u32(p) == value) return;
    const long long t0 = clock64();
    while (ld_acquire_u32(p) != value) {
        __nanosleep(20);
        if (clock64() - t0 > kPSpinLimit) {
            *err = 1;
            break;
        }
    }
}

__device__ __forceinline__ float p_silu_mul(float g, float u) { return __fdividef(g, 1.0f + __expf(-g)) * u; }

// contiguous range of the (tile, k-block) space owned by CTA c, walked tile piece by tile piece
struct SegIter {
    long long b, b1;
    int KB;
    __device__ SegIter(int tiles, int kblocks, int c, int P) : KB(kblocks) {
        const long long NB = (long long)tiles * kblocks;
        b = NB * c / P;
        b1 = NB * (c + 1) / P;
    }
    __device__ bool next(int& tile, int& ka, int& kb) {
        if (b >= b1) return false;
        tile = (int)(b / KB);
        ka = (int)(b - (long long)tile * KB);
        const long long rem = b1 - b;
        kb = (int)((long long)ka + rem < (long long)KB ? (long long)ka + rem : (long long)KB);
        b += kb - ka;
        return true;
    }
};

struct GemmPhase {
    const CUtensorMap* tw;
    const CUtensorMap* tx;
    int tiles, kblocks, mt, mode, phase;
};

// ---------------------------------------------------------------------------------------------- attention (one unit)
// Same arithmetic as attn_mma_kernel<HD, KW, 2> (attention.cu) for 4 warps = the 128 epilogue threads.
template <int HD, int KW>
__device__ __forceinline__ void attn_unit(const PArgs& a, const __nv_bfloat16* __restrict__ k_cache,
                                          const __nv_bfloat16* __restrict__ v_cache, uint8_t* aux, int* s_last,
                                          int seq, int g, int sp, int et, const unsigned* dep, unsigned dep_target) {
    constexpr int NS = 2;
    constexpr int CH = HD / 8, ROWB = HD * 2, KB = HD / 16, NT = KW / 8;
    const PersistLaunch& L = a.L;
    const int warp = et >> 5, lane = et & 31;
    const int rg = warp % a.rg_count, kg = warp / a.rg_count;
    uint8_t* sQ = aux;
    uint8_t* sK = sQ + (size_t)a.rg_count * 16 * ROWB;
    uint8_t* sV = sK + NS * kPKeyTile * ROWB;
    int* s_pt = reinterpret_cast<int*>(sV + NS * kPKeyTile * ROWB);
    const uint32_t uQ = smem_u32(sQ), uK = smem_u32(sK), uV = smem_u32(sV);
    const int G = L.nh / L.nkv;
    const int q0 = L.cu_q[seq], qlen = L.cu_q[seq + 1] - q0;
    const int R = qlen * G;
    const int kv_len = qlen > 0 ? L.positions[q0 + qlen - 1] + 1 : 0;
    const int kbeg = sp * L.split_keys;
    const bool active = qlen > 0 && kbeg < kv_len;       // uniform over the 128 threads
    const int kend = min(kv_len, kbeg + L.split_keys);
    const int nsplit_seq = active ? (kv_len + L.split_keys - 1) / L.split_keys : 1;
    const int old_keys = kv_len - qlen;
    const int page0 = kbeg / L.page_size;
    if (active) {
        const int* pt = L.page_table + (size_t)L.seq_slot[seq] * L.max_pages;
        const int npages = (kend - 1) / L.page_size - page0 + 1;
        for (int i = et; i < npages; i += 128) s_pt[i] = pt[page0 + i];
    }
    epi_sync();
    auto load_tile = [&](int tile, int buf) {
        const int j0 = kbeg + tile * kPKeyTile;
        const int sub = et & 3;
        for (int r = et >> 2; r < kPKeyTile; r += 32) {
            const int j = j0 + r;
            const bool ok = j < kend;
            const int jj = ok ? j : kbeg;
            const int pidx = a.page_shift >= 0 ? jj >> a.page_shift : jj / L.page_size;
            const int pin = a.page_shift >= 0 ? jj & (L.page_size - 1) : jj - pidx * L.page_size;
            const int page = s_pt[pidx - page0];
            const size_t off = (((size_t)page * L.nkv + g) * L.page_size + pin) * HD;
            const uint32_t drow = (uint32_t)(buf * kPKeyTile * ROWB + r * ROWB);
#pragma unroll
            for (int cc = 0; cc < CH / 4; ++cc) {
                const int ch = sub * (CH / 4) + cc;
                const uint32_t d = drow + ((ch ^ (r & 7)) << 4);
                cp_async16(uK + d, k_cache + off + ch * 8, ok);
                cp_async16(uV + d, v_cache + off + ch * 8, ok);
            }
        }
    };
    const int ntiles = active ? (kend - kbeg + kPKeyTile - 1) / kPKeyTile : 0;
    auto tile_is_old = [&](int t) { return kbeg + (t + 1) * kPKeyTile <= old_keys; };
    // keys older than this step's tokens do not depend on the QKV phase still finishing elsewhere
    if (0 < ntiles && tile_is_old(0)) load_tile(0, 0);
    cp_async_commit();
    if (et == 0) wait_counter(dep, dep_target, L.error);
    epi_sync();
    if (!active) return;

    for (int c = et; c < a.rg_count * 16 * CH; c += 128) {
        const int r = c / CH, ch = c - r * CH;
        const bool ok = r < R;
        const int t = ok ? r / G : 0, gq = ok ? r - t * G : 0;
        const __nv_bfloat16* src = L.q + ((size_t)(q0 + t) * L.nh + g * G + gq) * HD + ch * 8;
        cp_async16(uQ + r * ROWB + ((ch ^ (r & 7)) << 4), src, ok);
    }
    if (0 < ntiles && !tile_is_old(0)) load_tile(0, 0);
    cp_async_commit();

    const int r0 = rg * 16 + (lane >> 2), r1 = r0 + 8;
    const int qpos0 = kv_len - qlen + (r0 < R ? r0 / G : 0), qpos1 = kv_len - qlen + (r1 < R ? r1 / G : 0);
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
    float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    uint32_t qf[KB][4];

    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile % NS;
        if (tile == 0) cp_async_wait<0>();
        if (tile + NS - 1 < ntiles) load_tile(tile + NS - 1, (tile + NS - 1) % NS);
        cp_async_commit();
        if (tile > 0) cp_async_wait<NS - 1>();
        epi_sync();
        if (tile == 0) {
            const int mi = lane >> 3;
            const int row = rg * 16 + (lane & 7) + (mi & 1) * 8;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                const int ch = kb * 2 + (mi >> 1);
                ldsm_x4(uQ + row * ROWB + ((ch ^ (row & 7)) << 4), qf[kb]);
            }
        }
        const int key0 = kg * KW;
        if (kbeg + tile * kPKeyTile + key0 < kend) {
            float s[NT][4];
#pragma unroll
            for (int i = 0; i < NT; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f;
            const uint32_t kb_base = uK + buf * kPKeyTile * ROWB, vb_base = uV + buf * kPKeyTile * ROWB;
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
            const int jbase = kbeg + tile * kPKeyTile + key0 + (lane & 3) * 2;
            float tm0 = -INFINITY, tm1 = -INFINITY;
            const bool interior = kbeg + (tile + 1) * kPKeyTile <= old_keys + 1;
            if (interior) {
#pragma unroll
                for (int i = 0; i < NT; ++i) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        s[i][e] = r0 < R ? s[i][e] * a.scale_log2 : -INFINITY;
                        s[i][2 + e] = r1 < R ? s[i][2 + e] * a.scale_log2 : -INFINITY;
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
        epi_sync();
    }
    cp_async_wait<0>();
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // merge the key groups of a row group through shared memory (the K/V ring is idle now)
    float* xo = reinterpret_cast<float*>(sK);
    float* xml = xo + (size_t)4 * 16 * HD;
    if (a.kg_count > 1) {
        epi_sync();
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
        epi_sync();
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
    if (kg == 0) {
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int r = hrow ? r1 : r0;
            if (r >= R) continue;
            const int t = r / G, gq = r - t * G;
            const size_t th = (size_t)(q0 + t) * L.nh + g * G + gq;
            const float l = hrow ? l1 : l0;
            if (nsplit_seq == 1) {
                __nv_bfloat16* op = L.attn + th * HD;
                const float inv = 1.0f / l;
#pragma unroll
                for (int i = 0; i < HD / 8; ++i) {
                    const int d = i * 8 + (lane & 3) * 2;
                    *reinterpret_cast<__nv_bfloat162*>(op + d) =
                        __floats2bfloat162_rn(o[i][hrow * 2] * inv, o[i][hrow * 2 + 1] * inv);
                }
            } else {
                const size_t idx = th * L.nsplit_max + sp;
                float* op = L.o_part + idx * HD;
#pragma unroll
                for (int i = 0; i < HD / 8; ++i) {
                    const int d = i * 8 + (lane & 3) * 2;
                    *reinterpret_cast<float2*>(op + d) = make_float2(o[i][hrow * 2], o[i][hrow * 2 + 1]);
                }
                if ((lane & 3) == 0) {
                    L.ml_part[idx * 2] = hrow ? mx1 : mx0;
                    L.ml_part[idx * 2 + 1] = l;
                }
            }
        }
    }
    if (nsplit_seq == 1) {
        epi_sync();     // the K/V ring (and the merge buffers over it) may be refilled by the next unit
        return;
    }
    // the last unit of this (sequence, kv head) to finish merges the splits
    __threadfence();
    epi_sync();
    if (et == 0) {
        const int tk = atomicAdd(&L.tickets[seq * L.nkv + g], 1);
        *s_last = (tk == nsplit_seq - 1);
        if (*s_last) L.tickets[seq * L.nkv + g] = 0;
    }
    epi_sync();
    if (!*s_last) return;
    __threadfence();
    for (int idx = et; idx < R * (HD / 4); idx += 128) {
        const int r = idx / (HD / 4), d4 = idx - r * (HD / 4);
        const int t = r / G, gq = r - t * G;
        const size_t th = (size_t)(q0 + t) * L.nh + g * G + gq;
        const int qpos = kv_len - qlen + t;
        const int ns = min(nsplit_seq, qpos / L.split_keys + 1);
        const size_t base = th * L.nsplit_max;
        float mx = -INFINITY;
        for (int s2 = 0; s2 < ns; ++s2) mx = fmaxf(mx, __ldcg(&L.ml_part[(base + s2) * 2]));
        float l = 0.0f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s2 = 0; s2 < ns; ++s2) {
            const float w = exp2f(__ldcg(&L.ml_part[(base + s2) * 2]) - mx);
            l += __ldcg(&L.ml_part[(base + s2) * 2 + 1]) * w;
            const float4 v = __ldcg(reinterpret_cast<const float4*>(L.o_part + (base + s2) * HD) + d4);
            acc.x += v.x * w;
            acc.y += v.y * w;
            acc.z += v.z * w;
            acc.w += v.w * w;
        }
        const float inv = 1.0f / l;
        __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(L.attn + th * HD) + d4 * 2;
        op[0] = __floats2bfloat162_rn(acc.x * inv, acc.y * inv);
        op[1] = __floats2bfloat162_rn(acc.z * inv, acc.w * inv);
    }
    epi_sync();
}

// ---------------------------------------------------------------------------------------------- the kernel
template <int HD, int KW>
__global__ void __launch_bounds__(kPThreads, 1)
    fwd_persist_kernel(const __grid_constant__ CUtensorMap tm_xn, const __grid_constant__ CUtensorMap tm_attn,
                       const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_xsel,
                       const __grid_constant__ CUtensorMap tm_lm, const PArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint8_t* aux = ring + (size_t)a.stages * a.stage_bytes;            // T tile / attention buffers
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux + a.aux_bytes);
    uint64_t* empty_bar = full_bar + a.stages;
    uint64_t* acc_full = empty_bar + a.stages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    int* s_last = reinterpret_cast<int*>(tmem_slot + 1);
    float* scratch = reinterpret_cast<float*>(tmem_slot + 4);          // [8]
    float* rstd_s = scratch + 8;                                       // [128]
    int* kvrow_s = reinterpret_cast<int*>(rstd_s + 128);               // [128]

    const PersistLaunch& L = a.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x, P = gridDim.x;
    const int nL = L.n_layers;
    const bool want_logits = a.M_lm > 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&full_bar[s], 2);      // weight producer + activation producer
            mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);     // lane 0 of each epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int t_qkv = (L.nqkv + kPTile - 1) / kPTile, t_h = (L.h + kPTile - 1) / kPTile, t_gu = 2 * L.ffp / kPTile,
              t_v = (L.vocab + kPTile - 1) / kPTile;
    const int kb_h = (L.h + kPBlockK - 1) / kPBlockK, kb_q = (L.qdim + kPBlockK - 1) / kPBlockK,
              kb_f = (L.ffn + kPBlockK - 1) / kPBlockK;
    // GEMM number g of the forward: 4 per layer (QKV, O, gate|up, down), then the lm_head
    auto gemm_phase = [&](int gi) -> GemmPhase {
        GemmPhase ph;
        if (gi == 4 * nL) {
            ph.tw = &tm_lm, ph.tx = &tm_xsel, ph.tiles = t_v, ph.kblocks = kb_h, ph.mt = a.MT_lm, ph.mode = PM_LOGITS;
            ph.phase = 2 + 5 * nL;
            return ph;
        }
        const int l = gi >> 2, w = gi & 3;
        const PLayer* pl = L.layers + l;
        ph.mt = a.MT;
        ph.phase = 1 + 5 * l + (w == 0 ? 0 : w + 1);
        if (w == 0) ph.tw = &pl->t_qkv, ph.tx = &tm_xn, ph.tiles = t_qkv, ph.kblocks = kb_h, ph.mode = PM_QKV;
        else if (w == 1) ph.tw = &pl->t_o, ph.tx = &tm_attn, ph.tiles = t_h, ph.kblocks = kb_q, ph.mode = PM_RESID;
        else if (w == 2) ph.tw = &pl->t_gu, ph.tx = &tm_xn, ph.tiles = t_gu, ph.kblocks = kb_h, ph.mode = PM_SWIGLU;
        else ph.tw = &pl->t_down, ph.tx = &tm_act, ph.tiles = t_h, ph.kblocks = kb_f, ph.mode = PM_RESID;
        return ph;
    };
    const int n_gemms = 4 * nL + (want_logits ? 1 : 0);

    if (warp == 0) {
        // ------------------------------------------------------------------ weight producer (runs ahead of everything)
        // Two cursors walk the same static block schedule: the LOAD cursor fills the shared-memory ring (bounded by the
        // ring depth), the PREFETCH cursor runs `ahead` blocks further and only pulls blocks into L2
        // (cp.async.bulk.prefetch.tensor).  While the dependency chain of a phase boundary resolves the ring is full
        // and the load cursor is blocked, but HBM keeps streaming the next blocks into L2; the ring then refills from
        // L2 faster than HBM could deliver, so the boundary costs (almost) no HBM time.
        if (lane == 0) {
            const uint64_t pol_w = policy_evict_first();
            auto dims = [&](int gi, int& tiles, int& kbl) {
                if (gi == 4 * nL) { tiles = t_v, kbl = kb_h; return; }
                const int w = gi & 3;
                tiles = w == 0 ? t_qkv : (w == 2 ? t_gu : t_h);
                kbl = w == 1 ? kb_q : (w == 3 ? kb_f : kb_h);
            };
            auto wmap = [&](int gi) -> const CUtensorMap* {
                return gi == 4 * nL ? &tm_lm : &L.layers[gi >> 2].t_qkv + (gi & 3);
            };
            auto advance = [&](int& gi, long long& bb, long long& bb1, int& kbl) -> bool {
                while (bb >= bb1) {
                    if (++gi >= n_gemms) return false;
                    int tiles;
                    dims(gi, tiles, kbl);
                    const long long NB = (long long)tiles * kbl;
                    bb = NB * c / P;
                    bb1 = NB * (c + 1) / P;
                }
                return true;
            };
            int gi_l = -1, kb_l = 1, gi_p = -1, kb_p = 1;
            long long b_l = 0, e_l = 0, b_p = 0, e_p = 0;
            int it = 0, pt = 0;
            bool pf_live = a.ahead > 0;
            while (advance(gi_l, b_l, e_l, kb_l)) {
                while (pf_live && pt < it + a.stages + a.ahead) {
                    if (!advance(gi_p, b_p, e_p, kb_p)) {
                        pf_live = false;
                        break;
                    }
                    if (pt >= it + a.stages) {
                        const int tile = (int)(b_p / kb_p), k = (int)(b_p - (long long)tile * kb_p);
                        tma_prefetch_l2_2d(wmap(gi_p), k * kPBlockK, tile * kPTile);
                    }
                    ++b_p;
                    ++pt;
                }
                const int tile = (int)(b_l / kb_l), k = (int)(b_l - (long long)tile * kb_l);
                const int s = it % a.stages, round = it / a.stages;
                if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
                mbar_expect_tx(&full_bar[s], kPWBytes);
                tma_load_2d_hint(ring + (size_t)s * a.stage_bytes, wmap(gi_l), k * kPBlockK, tile * kPTile, &full_bar[s],
                                 pol_w);
                ++b_l;
                ++it;
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ activation producer
        if (lane == 0) {
            const uint64_t pol_x = policy_evict_last();
            int it = 0;
            for (int gi = 0; gi < n_gemms; ++gi) {
                const GemmPhase ph = gemm_phase(gi);
                SegIter si(ph.tiles, ph.kblocks, c, P);
                int tile, ka, kb;
                bool waited = false;
                while (si.next(tile, ka, kb)) {
                    if (!waited) {
                        wait_counter(a.counters + ph.phase - 1, (unsigned)P, L.error);
                        fence_proxy_async_all();     // peers' generic-proxy stores -> my async-proxy (TMA) loads
                        waited = true;
                    }
                    for (int k = ka; k < kb; ++k, ++it) {
                        const int s = it % a.stages, round = it / a.stages;
                        if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
                        mbar_expect_tx(&full_bar[s], (uint32_t)ph.mt * 128u);
                        tma_load_2d_hint(ring + (size_t)s * a.stage_bytes + kPWBytes, ph.tx, k * kPBlockK, 0, &full_bar[s],
                                         pol_x);
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int it = 0, nseg = 0;
            for (int gi = 0; gi < n_gemms; ++gi) {
                const GemmPhase ph = gemm_phase(gi);
                const uint32_t idesc = umma_idesc_bf16(kPTile, ph.mt);
                SegIter si(ph.tiles, ph.kblocks, c, P);
                int tile, ka, kb;
                while (si.next(tile, ka, kb)) {
                    const int acc = nseg & 1, use = nseg >> 1;
                    if (use > 0) mbar_wait(&acc_empty[acc], (use - 1) & 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(acc * 128);
                    for (int k = ka; k < kb; ++k, ++it) {
                        const int s = it % a.stages, round = it / a.stages;
                        mbar_wait(&full_bar[s], round & 1);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(ring + (size_t)s * a.stage_bytes);
                        const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + kPWBytes);
#pragma unroll
                        for (int kk = 0; kk < kPBlockK / 16; ++kk)
                            umma_f16(d, da + 2 * kk, db + 2 * kk, idesc, (uint32_t)((k > ka) | kk));
                        umma_commit(&empty_bar[s]);
                    }
                    umma_commit(&acc_full[acc]);
                    ++nseg;
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogues, glue, attention (128 threads)
        const int et = threadIdx.x - 128, ew = et >> 5;
        const int qd = warp & 3;                          // TMEM lane quarter this warp may read
        const int row = qd * 32 + lane;                   // accumulator row == et
        float* T = reinterpret_cast<float*>(aux);         // [token][128] fp32
        const float4* T4 = reinterpret_cast<const float4*>(aux);
        auto signal_phase = [&](int phase) {
            fence_proxy_async_all();
            epi_sync();
            if (et == 0) {
                red_release_add(a.counters + phase, 1u);
                if (L.trace) {
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    L.trace[(size_t)c * kPersistMaxPhases + phase + 1] = t;
                }
            }
        };
        // trace planes (each [CTAs][kPersistMaxPhases]): 0 = share of phase p done (slot p + 1; slot 0 = kernel entry),
        // 1 = first accumulator of the phase ready, 2 = last accumulator ready, 3 = partial flags of my tiles all seen
        auto tstamp = [&](int plane, int phase) {
            if (L.trace && et == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                L.trace[((size_t)plane * P + c) * kPersistMaxPhases + phase + 1] = t;
            }
        };
        if (L.trace && et == 0) {      // slot 0: kernel entry; slot p + 1: this CTA's share of phase p done
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            L.trace[(size_t)c * kPersistMaxPhases] = t;
        }
        // ---- phase 0: embedding gather + first-norm statistics, RoPE table, K/V cache row of every token
        {
            const int nvec = L.h >> 2;
            for (int m = c; m < L.M; m += P) {
                const int tok = L.tokens[m];
                const __nv_bfloat162* e = reinterpret_cast<const __nv_bfloat162*>(L.embed + (size_t)tok * L.h);
                const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(L.layers[0].ln1);
                float ss = 0.0f;
                for (int v = et; v < nvec; v += 128) {
                    const float2 x0 = __bfloat1622float2(e[2 * v]), x1 = __bfloat1622float2(e[2 * v + 1]);
                    const float2 w0 = __bfloat1622float2(wp[2 * v]), w1 = __bfloat1622float2(wp[2 * v + 1]);
                    reinterpret_cast<float4*>(L.resid + (size_t)m * L.h)[v] = make_float4(x0.x, x0.y, x1.x, x1.y);
                    __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(L.resid_bf + (size_t)m * L.h) + 2 * v;
                    op[0] = __floats2bfloat162_rn(x0.x * w0.x, x0.y * w0.y);
                    op[1] = __floats2bfloat162_rn(x1.x * w1.x, x1.y * w1.y);
                    ss += x0.x * x0.x + x0.y * x0.y + x1.x * x1.x + x1.y * x1.y;
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
                epi_sync();
                if (lane == 0) scratch[ew] = ss;
                epi_sync();
                if (et == 0) L.sumsq[m] = (scratch[0] + scratch[1]) + (scratch[2] + scratch[3]);
                const float pos = (float)L.positions[m];
                const int half = HD / 2;
                for (int i = et; i < half; i += 128) {
                    float sn, cs;
                    sincosf(pos * L.inv_freq[i], &sn, &cs);
                    L.rope_cs[(size_t)m * half + i] = make_float2(cs, sn);
                }
            }
            for (int t = et; t < a.MT; t += 128) {
                int r = 0;
                if (t < L.M) {
                    const int pos = L.positions[t];
                    const int page = L.page_table[(size_t)L.token_slot[t] * L.max_pages + pos / L.page_size];
                    r = page * L.nkv * L.page_size + pos % L.page_size;
                }
                kvrow_s[t] = r;
            }
            signal_phase(0);
        }
        int nseg = 0;
        int parts = 1;                                    // rows of sumsq that describe the current residual
        for (int gi = 0; gi < n_gemms; ++gi) {
            const GemmPhase ph = gemm_phase(gi);
            const int l = gi >> 2;
            const bool is_lm = ph.mode == PM_LOGITS;
            const PLayer* pl = L.layers + (is_lm ? 0 : l);
            // ---- non-GEMM work that precedes this GEMM
            if (ph.mode == PM_RESID && (gi & 3) == 1) {
                // attention of layer l (phase 2 + 5 l), after the QKV phase
                const int n_units = L.nseq * L.nkv * L.nsplit_max;
                const unsigned* dep = a.counters + (1 + 5 * l);
                bool any = false;
                for (int u = c; u < n_units; u += P) {
                    const int seq = u % L.nseq, g = (u / L.nseq) % L.nkv, sp = u / (L.nseq * L.nkv);
                    attn_unit<HD, KW>(a, pl->k_cache, pl->v_cache, aux, s_last, seq, g, sp, et, dep, (unsigned)P);
                    any = true;
                }
                (void)any;
                signal_phase(2 + 5 * l);
            }
            if (is_lm) {
                // logit-row gather (phase 1 + 5 nL), after the last down projection
                if (et == 0) wait_counter(a.counters + 5 * nL, (unsigned)P, L.error);
                epi_sync();
                if (L.logit_rows != nullptr) {
                    for (int r = c; r < a.M_lm; r += P) {
                        const int sr = L.logit_rows[r];
                        for (int t = et; t < parts; t += 128)
                            L.sumsq_sel[(size_t)t * L.Mx + r] = __ldcg(&L.sumsq[(size_t)t * L.Mx + sr]);
                        const uint4* s4 = reinterpret_cast<const uint4*>(L.resid_bf + (size_t)sr * L.h);
                        uint4* d4 = reinterpret_cast<uint4*>(L.xsel + (size_t)r * L.h);
                        for (int i = et; i < L.h / 8; i += 128) d4[i] = __ldcg(s4 + i);
                    }
                }
                signal_phase(1 + 5 * nL);
            }
            // ---- this GEMM's share
            SegIter si(ph.tiles, ph.kblocks, c, P);
            const long long NB = (long long)ph.tiles * ph.kblocks;
            const int n_tok = is_lm ? a.M_lm : L.M;                  // token columns that exist
            int tile, ka, kb;
            bool prepared = false;
            int n_def = 0, def_tile[2] = {0, 0};                     // partial pieces whose tiles are finished at phase end
            int piece = 0;
            // the epilogue of tokens [t_lo, t_hi) of `tile`, reading the reduced fp32 tile T[token][row]
            auto tile_epilogue = [&](int tile, int t_lo, int t_hi) {
                const int n4 = et & 31, vn = tile * kPTile + 4 * n4;
                if (ph.mode == PM_RESID) {
                    const bool n_ok = vn < L.h;
                    const __nv_bfloat16* ln_w = (gi & 3) == 1 ? pl->ln2 : (l + 1 < nL ? L.layers[l + 1].ln1 : L.final_norm);
                    float lw[4] = {0.f, 0.f, 0.f, 0.f};
                    if (n_ok) {
                        const uint2 w4 = *reinterpret_cast<const uint2*>(ln_w + vn);
                        const float2 w01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4.x));
                        const float2 w23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4.y));
                        lw[0] = w01.x, lw[1] = w01.y, lw[2] = w23.x, lw[3] = w23.y;
                    }
#pragma unroll 1
                    for (int t = t_lo + ew; t < t_hi; t += 4) {      // t is warp-uniform
                        float sq = 0.0f;
                        if (n_ok) {
                            float* o = L.resid + (size_t)t * L.h + vn;
                            const float4 old = __ldcg(reinterpret_cast<const float4*>(o));
                            float4 v = T4[(t << 5) + n4];
                            v.x = old.x + v.x, v.y = old.y + v.y, v.z = old.z + v.z, v.w = old.w + v.w;
                            __stcg(reinterpret_cast<float4*>(o), v);
                            const __nv_bfloat162 p01 = __floats2bfloat162_rn(v.x * lw[0], v.y * lw[1]);
                            const __nv_bfloat162 p23 = __floats2bfloat162_rn(v.z * lw[2], v.w * lw[3]);
                            uint2 pk;
                            pk.x = *reinterpret_cast<const uint32_t*>(&p01);
                            pk.y = *reinterpret_cast<const uint32_t*>(&p23);
                            *reinterpret_cast<uint2*>(L.resid_bf + (size_t)t * L.h + vn) = pk;
                            sq = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                        }
#pragma unroll
                        for (int d = 16; d >= 1; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
                        if (lane == 0) L.sumsq[(size_t)tile * L.Mx + t] = sq;
                    }
                } else if (ph.mode == PM_LOGITS) {
                    const bool vec = (L.logits_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(L.logits) & 15) == 0 &&
                                     vn + 4 <= L.vocab;
#pragma unroll 1
                    for (int t = t_lo + ew; t < t_hi; t += 4) {
                        if (vn >= L.vocab) continue;
                        const float rs = rstd_s[t];
                        float4 v = T4[(t << 5) + n4];
                        v.x *= rs, v.y *= rs, v.z *= rs, v.w *= rs;
                        float* o = L.logits + (size_t)t * L.logits_ld + vn;
                        if (vec) {
                            __stcs(reinterpret_cast<float4*>(o), v);
                        } else {
                            const float vv[4] = {v.x, v.y, v.z, v.w};
                            for (int i = 0; i < 4; ++i)
                                if (vn + i < L.vocab) o[i] = vv[i];
                        }
                    }
                } else if (ph.mode == PM_SWIGLU) {
                    // rows 0..63 = gate, 64..127 = up of ff index tile*64 + (row & 63): 8 consecutive ff outputs of one
                    // token per thread and step, one 16-byte store
                    const bool wide = (L.ffn & 7) == 0;
#pragma unroll 1
                    for (int g = et; g < (t_hi - t_lo) * 8; g += 128) {
                        const int col = t_lo + (g >> 3), f8 = (g & 7) * 8;
                        const int j0 = tile * 64 + f8;
                        if (j0 >= L.ffn) continue;
                        const float rs = rstd_s[col];
                        const float4* tp = reinterpret_cast<const float4*>(T + col * kPTile + f8);
                        const float4 g0 = tp[0], g1 = tp[1], u0 = tp[16], u1 = tp[17];
                        float v[8];
                        v[0] = p_silu_mul(g0.x * rs, u0.x * rs), v[1] = p_silu_mul(g0.y * rs, u0.y * rs);
                        v[2] = p_silu_mul(g0.z * rs, u0.z * rs), v[3] = p_silu_mul(g0.w * rs, u0.w * rs);
                        v[4] = p_silu_mul(g1.x * rs, u1.x * rs), v[5] = p_silu_mul(g1.y * rs, u1.y * rs);
                        v[6] = p_silu_mul(g1.z * rs, u1.z * rs), v[7] = p_silu_mul(g1.w * rs, u1.w * rs);
                        __nv_bfloat16* o = L.act + (size_t)col * L.ffn + j0;
                        if (wide && j0 + 8 <= L.ffn) {
                            __nv_bfloat162 p[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                            *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(p);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (j0 + i < L.ffn) o[i] = __float2bfloat16(v[i]);
                        }
                    }
                } else {
                    // QKV: rstd, bias, rotate-half RoPE, q store / paged K,V append; 4 rotary pairs per thread and step
                    constexpr int half = HD / 2, heads_per_tile = kPTile / HD;
                    const int gi16 = et & 15;
                    const int qh2 = HD == 64 ? gi16 >> 3 : 0, qi = HD == 64 ? (gi16 & 7) * 4 : gi16 * 4;
                    const int qhead = tile * heads_per_tile + qh2;
                    const bool head_ok = qhead < L.nh + 2 * L.nkv;
                    const bool rot = qhead < L.nh + L.nkv;
                    float qb1[4] = {0.f, 0.f, 0.f, 0.f}, qb2[4] = {0.f, 0.f, 0.f, 0.f};
                    if (head_ok) {
                        const uint2 b1 = *reinterpret_cast<const uint2*>(pl->bqkv + qhead * HD + qi);
                        const uint2 b2 = *reinterpret_cast<const uint2*>(pl->bqkv + qhead * HD + qi + half);
                        const float2 a01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&b1.x));
                        const float2 a23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&b1.y));
                        const float2 c01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&b2.x));
                        const float2 c23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&b2.y));
                        qb1[0] = a01.x, qb1[1] = a01.y, qb1[2] = a23.x, qb1[3] = a23.y;
                        qb2[0] = c01.x, qb2[1] = c01.y, qb2[2] = c23.x, qb2[3] = c23.y;
                    }
                    const int f1 = (qh2 * HD + qi) >> 2, f2 = (qh2 * HD + half + qi) >> 2;
#pragma unroll 1
                    for (int t = t_lo + (et >> 4); t < t_hi; t += 8) {
                        if (!head_ok) continue;
                        float4 c01 = make_float4(1.f, 0.f, 1.f, 0.f), c23 = c01;
                        if (rot) {
                            const float4* cs4 = reinterpret_cast<const float4*>(L.rope_cs + (size_t)t * half + qi);
                            c01 = __ldcg(cs4);
                            c23 = __ldcg(cs4 + 1);
                        }
                        float4 x1 = T4[(t << 5) + f1], x2 = T4[(t << 5) + f2];
                        const float rs = rstd_s[t];
                        x1.x *= rs, x1.y *= rs, x1.z *= rs, x1.w *= rs;
                        x2.x *= rs, x2.y *= rs, x2.z *= rs, x2.w *= rs;
                        x1.x += qb1[0], x1.y += qb1[1], x1.z += qb1[2], x1.w += qb1[3];
                        x2.x += qb2[0], x2.y += qb2[1], x2.z += qb2[2], x2.w += qb2[3];
                        const __nv_bfloat162 lo01 = __floats2bfloat162_rn(x1.x * c01.x - x2.x * c01.y, x1.y * c01.z - x2.y * c01.w);
                        const __nv_bfloat162 lo23 = __floats2bfloat162_rn(x1.z * c23.x - x2.z * c23.y, x1.w * c23.z - x2.w * c23.w);
                        const __nv_bfloat162 hi01 = __floats2bfloat162_rn(x2.x * c01.x + x1.x * c01.y, x2.y * c01.z + x1.y * c01.w);
                        const __nv_bfloat162 hi23 = __floats2bfloat162_rn(x2.z * c23.x + x1.z * c23.y, x2.w * c23.z + x1.w * c23.w);
                        __nv_bfloat16* dstp;
                        if (qhead < L.nh) {
                            dstp = L.q + ((size_t)t * L.nh + qhead) * HD;
                        } else {
                            const int kvh = rot ? qhead - L.nh : qhead - L.nh - L.nkv;
                            __nv_bfloat16* cache = rot ? pl->k_cache : pl->v_cache;
                            dstp = cache + (size_t)(kvrow_s[t] + kvh * L.page_size) * HD;
                        }
                        uint2 lo, hi;
                        lo.x = *reinterpret_cast<const uint32_t*>(&lo01);
                        lo.y = *reinterpret_cast<const uint32_t*>(&lo23);
                        hi.x = *reinterpret_cast<const uint32_t*>(&hi01);
                        hi.y = *reinterpret_cast<const uint32_t*>(&hi23);
                        *reinterpret_cast<uint2*>(dstp + qi) = lo;
                        *reinterpret_cast<uint2*>(dstp + qi + half) = hi;
                    }
                }
            };
            // One loop, one epilogue call site (code size: every instruction of the tail is fetched once per phase, so
            // the tail is kept small): first this CTA's pieces in range order, then the tiles it holds partial pieces of.
            int d_next = 0;
            for (;;) {
                int e_tile, e_lo, e_hi;
                if (si.next(tile, ka, kb)) {
                    if (!prepared) {
                        prepared = true;
                        if (ph.mode != PM_RESID) {
                            // consumer half of the fused RMSNorm: per-token rstd from the producer's per-tile sums
                            if (et == 0) wait_counter(a.counters + ph.phase - 1, (unsigned)P, L.error);
                            epi_sync();
                            const float* ss = (is_lm && L.logit_rows != nullptr) ? L.sumsq_sel : L.sumsq;
                            for (int t = et; t < ph.mt; t += 128) {
                                float sum = 0.0f;
                                if (t < n_tok)
                                    for (int p2 = 0; p2 < parts; ++p2) sum += __ldcg(&ss[(size_t)p2 * L.Mx + t]);
                                rstd_s[t] = rsqrtf(sum / (float)L.h + L.eps);
                            }
                            epi_sync();
                        }
                    }
                    const int acc = nseg & 1, use = nseg >> 1;
                    ++nseg;
                    mbar_wait(&acc_full[acc], use & 1);
                    tc_fence_after();
                    if (piece == 0) tstamp(1, ph.phase);
                    tstamp(2, ph.phase);
                    const uint32_t taddr = tmem_base + (uint32_t)(acc * 128) + ((uint32_t)(qd * 32) << 16);
                    const bool whole = ka == 0 && kb == ph.kblocks;
                    // whole tile: accumulator -> T[token][row] in shared memory; partial piece: -> my L2-resident slot
                    // (0: first piece of my range, 1: a later one) as [token][row], flag = phase + 1
                    const int slot = piece == 0 ? 0 : 1;
                    float* dst = whole ? T : L.part_ws + ((size_t)c * 2 + slot) * (kPTile * 128);
                    ++piece;
#pragma unroll 1
                    for (int c0 = 0; c0 < ph.mt; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld_32x32(taddr + c0, r);
                        tmem_ld_wait();
                        if (whole) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) dst[(c0 + i) * kPTile + row] = __uint_as_float(r[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (c0 + i < n_tok) __stcg(&dst[(c0 + i) * kPTile + row], __uint_as_float(r[i]));
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[acc]);
                    epi_sync();
                    if (!whole) {
                        // st.release.gpu orders the partial (the other threads' stores included: they happened before
                        // the barrier) before the flag; no separate fence
                        if (et == 0) st_release_u32(a.part_flag + c * 2 + slot, (unsigned)(ph.phase + 1));
                        tstamp(4, ph.phase);
                        def_tile[n_def++] = tile;
                        continue;
                    }
                    e_tile = tile, e_lo = 0, e_hi = n_tok;
                } else if (d_next < n_def) {
                    // ---- a tile this CTA holds a piece of: its S pieces' CTAs each reduce and finish 1/S of the tokens
                    // (partials summed in k order: deterministic), so the phase's tail is S-way parallel
                    const int dt = def_tile[d_next++];
                    const long long tb = (long long)dt * ph.kblocks, te = tb + ph.kblocks;
                    const bool dense = NB >= P;       // every CTA holds at least one block (else: exactly one or none)
                    const int cf = (int)(((tb + 1) * P + NB - 1) / NB) - 1;      // CTA holding the tile's first k-block
                    const int cl = (int)((te * P + NB - 1) / NB) - 1;            // ... and its last
                    const int S = dense ? cl - cf + 1 : ph.kblocks;
                    auto member_word = [&](int j) -> int {                       // j-th piece in k order: 2 * cta + slot
                        const int c2 = dense ? cf + j : (int)(((tb + j + 1) * P + NB - 1) / NB) - 1;
                        return c2 * 2 + ((NB * c2 / P >= tb) ? 0 : 1);
                    };
                    int rank = 0;
                    for (int j = 0; j < S; ++j)
                        if ((member_word(j) >> 1) == c) rank = j;
                    const int per = (n_tok + S - 1) / S;
                    e_tile = dt, e_lo = rank * per, e_hi = min(n_tok, e_lo + per);
                    if (ew == 0) {      // one flag per lane: all of them polled in one L2 round trip
                        for (int j = lane; j < S; j += 32)
                            wait_flag_eq(a.part_flag + member_word(j), (unsigned)(ph.phase + 1), L.error);
                    }
                    epi_sync();
                    tstamp(3, ph.phase);
                    if (e_lo < e_hi) {
                        float4* Tw = reinterpret_cast<float4*>(aux);
                        const int i_hi = e_hi * 32;
#pragma unroll 1
                        for (int i0 = e_lo * 32 + et; i0 < i_hi; i0 += 256) {
                            const bool two = i0 + 128 < i_hi;
                            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll 1
                            for (int j0 = 0; j0 < S; j0 += 4) {      // 8 independent 16-byte loads in flight per thread
                                float4 v0[4], v1[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    v0[j] = v1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                                    if (j0 + j < S) {
                                        const float4* ws4 = reinterpret_cast<const float4*>(
                                            L.part_ws + (size_t)member_word(j0 + j) * (kPTile * 128));
                                        v0[j] = __ldcg(ws4 + i0);
                                        if (two) v1[j] = __ldcg(ws4 + i0 + 128);
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    s0.x += v0[j].x, s0.y += v0[j].y, s0.z += v0[j].z, s0.w += v0[j].w;
                                    s1.x += v1[j].x, s1.y += v1[j].y, s1.z += v1[j].z, s1.w += v1[j].w;
                                }
                            }
                            Tw[i0] = s0;
                            if (two) Tw[i0 + 128] = s1;
                        }
                    }
                    epi_sync();
                    tstamp(5, ph.phase);
                } else {
                    break;
                }
                if (e_lo < e_hi) tile_epilogue(e_tile, e_lo, e_hi);
                epi_sync();      // T is rewritten by the next piece
                tstamp(6, ph.phase);
            }
            signal_phase(ph.phase);
            if (ph.mode == PM_RESID) parts = t_h;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------------- host
static int g_p_sms = 0, g_p_smem = 0;
static int p_props() {
    if (g_p_sms) return 0;
    int dev = 0;
    ASD_CUDA(cudaGetDevice(&dev));
    ASD_CUDA(cudaDeviceGetAttribute(&g_p_sms, cudaDevAttrMultiProcessorCount, dev));
    ASD_CUDA(cudaDeviceGetAttribute(&g_p_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return 0;
}

int persist_num_ctas() {
    if (p_props()) return 0;
    return g_p_sms < 256 ? g_p_sms : 256;
}
size_t persist_part_ws_bytes() { return (size_t)persist_num_ctas() * 2 * kPTile * 128 * sizeof(float); }

bool persist_supported(int M, int n_logit_rows, int max_qlen, int nh, int nkv, int hd, int page_size, int h, int n_layers) {
    if (M <= 0 || M > 128 || n_logit_rows > 128 || n_layers > 128) return false;
    if (hd != 64 && hd != 128) return false;
    if (nkv <= 0 || nh % nkv) return false;
    if (max_qlen * (nh / nkv) > 64) return false;
    if (h % 8 || page_size <= 0) return false;
    return true;
}

template <int HD, int KW>
static int p_launch_t(const CUtensorMap maps[5], const PArgs& a, int grid, size_t smem, cudaStream_t stream) {
    static PerDeviceOnce once;
    if (once.need()) {
        ASD_CUDA(cudaFuncSetAttribute(fwd_persist_kernel<HD, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_p_smem));
        prefer_max_smem(fwd_persist_kernel<HD, KW>);
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kPThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ASD_CUDA(cudaLaunchKernelEx(&cfg, fwd_persist_kernel<HD, KW>, maps[0], maps[1], maps[2], maps[3], maps[4], a));
    return 0;
}

int persist_launch(const PersistLaunch& L, cudaStream_t stream) {
    if (p_props()) return -1;
    if (!persist_supported(L.M, L.n_logit_rows, L.max_qlen, L.nh, L.nkv, L.hd, L.page_size, L.h, L.n_layers))
        return set_error("persist: unsupported shape");
    PArgs a;
    a.L = L;
    a.MT = gemm_token_tile(L.M);
    a.M_lm = L.n_logit_rows;
    a.MT_lm = a.M_lm > 0 ? gemm_token_tile(a.M_lm) : 16;
    a.MTmax = a.MT > a.MT_lm ? a.MT : a.MT_lm;
    const int G = L.nh / L.nkv, rows = L.max_qlen * G;
    const int rg = (rows + 15) / 16;
    a.rg_count = rg <= 1 ? 1 : (rg == 2 ? 2 : 4);
    a.kg_count = 4 / a.rg_count;
    a.page_shift = -1;
    for (int b = 0; b < 16; ++b)
        if ((1 << b) == L.page_size) a.page_shift = b;
    a.scale_log2 = 1.4426950408889634f / sqrtf((float)L.hd);
    a.n_phases = 3 + 5 * L.n_layers;
    a.counters = L.sync;
    a.part_flag = L.sync + kPersistMaxPhases;
    a.stage_bytes = kPWBytes + a.MTmax * 128;
    const size_t attn_bytes = (size_t)a.rg_count * 16 * L.hd * 2 + 4 * (size_t)kPKeyTile * L.hd * 2 +
                              ((size_t)L.split_keys / L.page_size + 2) * sizeof(int);
    const size_t merge_bytes = (size_t)a.rg_count * 16 * L.hd * 2 + (size_t)4 * 16 * (L.hd + 2) * 4;
    size_t aux = (size_t)a.MTmax * kPTile * 4;
    if (attn_bytes > aux) aux = attn_bytes;
    if (merge_bytes > aux) aux = merge_bytes;
    aux = (aux + 1023) / 1024 * 1024;
    a.aux_bytes = (int)aux;
    const int small = 2048;
    int stages = (int)((g_p_smem - 1024 - (int)aux - small) / a.stage_bytes);
    if (stages > 12) stages = 12;
    if (stages < 2) return set_error("persist: shared memory does not fit a two-stage ring");
    a.stages = stages;
    a.ahead = L.prefetch_ahead;
    const size_t smem = 1024 + (size_t)stages * a.stage_bytes + aux + small;
    CUtensorMap maps[5];
    if (make_tmap_bf16(&maps[0], L.resid_bf, L.M, L.h, L.h, a.MT)) return -1;
    if (make_tmap_bf16(&maps[1], L.attn, L.M, L.qdim, L.qdim, a.MT)) return -1;
    if (make_tmap_bf16(&maps[2], L.act, L.M, L.ffn, L.ffn, a.MT)) return -1;
    if (a.M_lm > 0) {
        const void* src = L.logit_rows ? (const void*)L.xsel : (const void*)L.resid_bf;
        if (make_tmap_bf16(&maps[3], src, a.M_lm, L.h, L.h, a.MT_lm)) return -1;
    } else {
        maps[3] = maps[0];
    }
    if (make_tmap_bf16(&maps[4], L.lm_head, L.vocab, L.h, L.h, 128)) return -1;
    const int grid = persist_num_ctas();
    ASD_CUDA(cudaMemsetAsync(L.sync, 0, sizeof(unsigned) * kPersistSyncWords, stream));
    int rc;
#define ASD_P(HD_)                                                                                    \
    (a.rg_count == 1 ? p_launch_t<HD_, 16>(maps, a, grid, smem, stream)                               \
                     : (a.rg_count == 2 ? p_launch_t<HD_, 32>(maps, a, grid, smem, stream)            \
                                        : p_launch_t<HD_, 64>(maps, a, grid, smem, stream)))
    rc = L.hd == 128 ? ASD_P(128) : ASD_P(64);
#undef ASD_P
    if (rc) return rc;
    count_launch(1);
    return 0;
}

}  // namespace asd
