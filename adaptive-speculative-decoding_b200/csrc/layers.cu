// Memory-bound glue kernels of the Qwen2 verify / draft forward.  Each one is fused with the
// deterministic reduction of the GEMM's K-split fp32 slices so partial sums never make an extra
// round trip:
//   add_norm    : resid += sum_s part[s]  (or resid = embedding row);  x = RMSNorm(resid) * w  (bf16)
//   qkv_rope    : qkv = sum_s part[s] + bias;  RoPE(q, k);  q -> bf16 buffer, k/v -> paged KV cache
//   gather_rows : pick the rows whose logits are wanted
// Model semantics follow HF Qwen2 (RMSNorm eps inside rsqrt, rotate_half RoPE, QKV bias), which is
// what the reference's Stage wraps through vLLM (docs/guides/RESEARCH_PROTOCOL.md:233-304).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "asd_internal.h"
#include "layers.h"
#include "ptx.cuh"

namespace asd {

static void glue_carveout();   // same shared-memory carve-out for every glue kernel (see prefer_max_smem)

constexpr int kNormThreads = 256;
constexpr int kNormMaxVec = 8;  // float4 per thread -> hidden <= 8192

__device__ __forceinline__ float block_sum(float v, float* scratch) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = 0.0f;
    for (int w = 0; w < nw; ++w) t += scratch[w];  // fixed order: deterministic
    return t;
}

// one CTA per token row
__global__ void __launch_bounds__(kNormThreads) add_norm_kernel(float* __restrict__ resid, const float* __restrict__ part,
                                                                 int nslices, size_t slice_stride,
                                                                 const int* __restrict__ tokens,
                                                                 const __nv_bfloat16* __restrict__ emb,
                                                                 const __nv_bfloat16* __restrict__ w,
                                                                 __nv_bfloat16* __restrict__ xnorm, int h, float eps,
                                                                 __nv_bfloat16* __restrict__ resid_bf,
                                                                 float* __restrict__ sumsq0) {
    __shared__ float scratch[kNormThreads / 32];
    grid_dep_wait();    // launched programmatically: the upstream kernel's results are needed from here on
    grid_dep_launch();  // lets the next GEMM start streaming its weights (it waits before reading x)
    const int m = blockIdx.x, nvec = h >> 2;
    float4 v[kNormMaxVec];
    float ss = 0.0f;
    float4* rrow = reinterpret_cast<float4*>(resid + (size_t)m * h);
#pragma unroll
    for (int i = 0; i < kNormMaxVec; ++i) {
        const int c = threadIdx.x + i * kNormThreads;
        if (c < nvec) {
            float4 x;
            if (tokens) {
                const __nv_bfloat162* e =
                    reinterpret_cast<const __nv_bfloat162*>(emb + (size_t)tokens[m] * h) + 2 * c;
                const float2 a = __bfloat1622float2(e[0]), b = __bfloat1622float2(e[1]);
                x = make_float4(a.x, a.y, b.x, b.y);
            } else {
                x = rrow[c];
                for (int s = 0; s < nslices; ++s) {
                    const float4 p = reinterpret_cast<const float4*>(part + s * slice_stride + (size_t)m * h)[c];
                    x.x += p.x;
                    x.y += p.y;
                    x.z += p.z;
                    x.w += p.w;
                }
            }
            rrow[c] = x;
            v[i] = x;
            ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        }
    }
    const float tot = block_sum(ss, scratch);
    if (resid_bf != nullptr) {   // fused-norm mode: the consumer GEMM normalises; publish bf16(resid) and sum x^2
#pragma unroll
        for (int i = 0; i < kNormMaxVec; ++i) {
            const int c = threadIdx.x + i * kNormThreads;
            if (c < nvec) {
                const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(w) + 2 * c;
                const float2 w0 = __bfloat1622float2(wp[0]), w1 = __bfloat1622float2(wp[1]);
                __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(resid_bf + (size_t)m * h) + 2 * c;
                op[0] = __floats2bfloat162_rn(v[i].x * w0.x, v[i].y * w0.y);
                op[1] = __floats2bfloat162_rn(v[i].z * w1.x, v[i].w * w1.y);
            }
        }
        if (threadIdx.x == 0) sumsq0[m] = tot;
    }
    const float inv = rsqrtf(tot / (float)h + eps);
    if (xnorm == nullptr) return;
#pragma unroll
    for (int i = 0; i < kNormMaxVec; ++i) {
        const int c = threadIdx.x + i * kNormThreads;
        if (c < nvec) {
            const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(w) + 2 * c;
            const float2 w0 = __bfloat1622float2(wp[0]), w1 = __bfloat1622float2(wp[1]);
            __nv_bfloat162 o0 = __floats2bfloat162_rn(v[i].x * inv * w0.x, v[i].y * inv * w0.y);
            __nv_bfloat162 o1 = __floats2bfloat162_rn(v[i].z * inv * w1.x, v[i].w * inv * w1.y);
            __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(xnorm + (size_t)m * h) + 2 * c;
            op[0] = o0;
            op[1] = o1;
        }
    }
}

int launch_add_norm(float* resid, const float* part, int nslices, size_t slice_stride, const int* tokens,
                    const __nv_bfloat16* emb, const __nv_bfloat16* w, __nv_bfloat16* xnorm, int M, int h, float eps,
                    cudaStream_t stream, __nv_bfloat16* resid_bf, float* sumsq0) {
    if (h % 4 || h > kNormThreads * kNormMaxVec * 4) return set_error("add_norm: hidden must be %%4 and <= 8192");
    if (M <= 0) return 0;
    glue_carveout();
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(M);
    cfg.blockDim = dim3(kNormThreads);
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = tuning().glue_pdl ? 1 : 0;
    ASD_CUDA(cudaLaunchKernelEx(&cfg, add_norm_kernel, resid, part, nslices, slice_stride, tokens, emb, w, xnorm, h, eps,
                                resid_bf, sumsq0));
    count_launch(1);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-parallel boundary, fused: one-shot all-reduce over NVLink peer memory + residual add + RMSNorm
// statistics in ONE kernel (replaces ncclAllReduce + add_norm).  Every rank's row-parallel GEMM leaves its
// fp32 partial [M, h] in a buffer that all peers have mapped (CUDA IPC); this kernel
//   1. publishes "my partial for epoch e is ready" into every peer's flag array (system-scope release),
//   2. waits until every peer's flag for this epoch has arrived (acquire, bounded spin),
//   3. loads the `world` partials straight from peer memory (NVLink P2P loads, L1 bypassed), adds them in rank
//      order - so all ranks compute bit-identical sums - plus the residual, and emits the fp32 residual, the
//      bf16 operand of the next GEMM and the per-token sum of squares exactly like add_norm_kernel.
// The partial buffers are double-buffered by the caller; a rank can only reach all-reduce n+1's flag after it
// finished reading all-reduce n, so when a GEMM overwrites buffer n%2 again (for all-reduce n+2) every peer is
// done with it.
struct TpPeers {
    const float* buf[8];   // this epoch's partial buffer of every rank (peer-mapped pointers)
    uint32_t* flags[8];    // flag array of every rank: flags[p][r] = last epoch rank r published to rank p
    int rank, world;
    uint32_t epoch;
    int* error;            // set to 1 if a peer never showed up
    // two-shot (world >= 4): row m is reduced by its home rank m % world only, which publishes the final row in
    // its broadcast buffer; the other ranks fetch that row instead of all `world` partials
    int two_shot;
    const float* bcast[8]; // broadcast buffer of every rank (peer-mapped), [M][h] fp32
    uint32_t* rowflags[8]; // rowflags[p][m] = last epoch whose final row m is readable in its home's buffer
};

__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}

__global__ void __launch_bounds__(kNormThreads) tp_allreduce_norm_kernel(
    const TpPeers tp, float* __restrict__ resid, const __nv_bfloat16* __restrict__ w,
    __nv_bfloat16* __restrict__ xnorm, int h, float eps, __nv_bfloat16* __restrict__ resid_bf,
    float* __restrict__ sumsq0) {
    __shared__ float scratch[kNormThreads / 32];
    grid_dep_launch();
    if (blockIdx.x == 0 && threadIdx.x < tp.world && (int)threadIdx.x != tp.rank) {
        __threadfence_system();   // the GEMM that produced my partial finished before this kernel started
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(tp.flags[threadIdx.x] + tp.rank), "r"(tp.epoch)
                     : "memory");
    }
    if (threadIdx.x < tp.world && (int)threadIdx.x != tp.rank) {
        const uint32_t* f = tp.flags[tp.rank] + threadIdx.x;
        uint32_t seen;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
            if ((int)(seen - tp.epoch) >= 0) break;
            if (clock64() - t0 > 4000000000LL) {   // ~2 s: a peer died; fail loudly instead of hanging the GPU
                *tp.error = 1;
                break;
            }
        } while (true);
    }
    __syncthreads();
    const int m = blockIdx.x, nvec = h >> 2;
    float4 v[kNormMaxVec];
    float ss = 0.0f;
    float4* rrow = reinterpret_cast<float4*>(resid + (size_t)m * h);
    const int home = tp.two_shot ? m % tp.world : tp.rank;
    if (home == tp.rank) {
        float4* brow = tp.two_shot ? reinterpret_cast<float4*>(const_cast<float*>(tp.bcast[tp.rank]) + (size_t)m * h) : nullptr;
#pragma unroll
        for (int i = 0; i < kNormMaxVec; ++i) {
            const int c = threadIdx.x + i * kNormThreads;
            if (c < nvec) {
                float4 x = rrow[c];
                for (int r = 0; r < tp.world; ++r) {
                    const float4 p = ld_peer_f4(tp.buf[r] + (size_t)m * h + 4 * c);
                    x.x += p.x;
                    x.y += p.y;
                    x.z += p.z;
                    x.w += p.w;
                }
                rrow[c] = x;
                if (brow) brow[c] = x;
                v[i] = x;
                ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
            }
        }
        if (tp.two_shot) {
            __syncthreads();   // the whole final row is written
            if (threadIdx.x < tp.world && (int)threadIdx.x != tp.rank) {
                __threadfence_system();
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(tp.rowflags[threadIdx.x] + m), "r"(tp.epoch)
                             : "memory");
            }
        }
    } else {
        if (threadIdx.x == 0) {
            const uint32_t* f = tp.rowflags[tp.rank] + m;
            uint32_t seen;
            const long long t0 = clock64();
            do {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
                if ((int)(seen - tp.epoch) >= 0) break;
                if (clock64() - t0 > 4000000000LL) {
                    *tp.error = 1;
                    break;
                }
            } while (true);
        }
        __syncthreads();
        const float* brow = tp.bcast[home] + (size_t)m * h;
#pragma unroll
        for (int i = 0; i < kNormMaxVec; ++i) {
            const int c = threadIdx.x + i * kNormThreads;
            if (c < nvec) {
                const float4 x = ld_peer_f4(brow + 4 * c);   // residual + all partials, summed by the home rank
                rrow[c] = x;
                v[i] = x;
                ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
            }
        }
    }
    const float tot = block_sum(ss, scratch);
    const float inv = rsqrtf(tot / (float)h + eps);
#pragma unroll
    for (int i = 0; i < kNormMaxVec; ++i) {
        const int c = threadIdx.x + i * kNormThreads;
        if (c < nvec && w != nullptr) {
            const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(w) + 2 * c;
            const float2 w0 = __bfloat1622float2(wp[0]), w1 = __bfloat1622float2(wp[1]);
            if (resid_bf != nullptr) {
                __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(resid_bf + (size_t)m * h) + 2 * c;
                op[0] = __floats2bfloat162_rn(v[i].x * w0.x, v[i].y * w0.y);
                op[1] = __floats2bfloat162_rn(v[i].z * w1.x, v[i].w * w1.y);
            }
            if (xnorm != nullptr) {
                __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(xnorm + (size_t)m * h) + 2 * c;
                op[0] = __floats2bfloat162_rn(v[i].x * inv * w0.x, v[i].y * inv * w0.y);
                op[1] = __floats2bfloat162_rn(v[i].z * inv * w1.x, v[i].w * inv * w1.y);
            }
        }
    }
    if (sumsq0 != nullptr && threadIdx.x == 0) sumsq0[m] = tot;
}

int launch_tp_allreduce_norm(const float* const* peer_bufs, uint32_t* const* peer_flags, int rank, int world,
                             uint32_t epoch, int* error, float* resid, const __nv_bfloat16* w, __nv_bfloat16* xnorm,
                             int M, int h, float eps, __nv_bfloat16* resid_bf, float* sumsq0, cudaStream_t stream,
                             const float* const* peer_bcast, uint32_t* const* peer_rowflags) {
    if (world > 8) return set_error("tp all-reduce: world size <= 8");
    if (h % 4 || h > kNormThreads * kNormMaxVec * 4) return set_error("tp all-reduce: hidden must be %%4 and <= 8192");
    TpPeers tp;
    for (int r = 0; r < 8; ++r) {
        tp.buf[r] = r < world ? peer_bufs[r] : nullptr;
        tp.flags[r] = r < world ? peer_flags[r] : nullptr;
    }
    tp.rank = rank;
    tp.world = world;
    tp.epoch = epoch;
    tp.error = error;
    tp.two_shot = peer_bcast != nullptr;
    for (int r = 0; r < 8; ++r) {
        tp.bcast[r] = (peer_bcast && r < world) ? peer_bcast[r] : nullptr;
        tp.rowflags[r] = (peer_rowflags && r < world) ? peer_rowflags[r] : nullptr;
    }
    glue_carveout();
    tp_allreduce_norm_kernel<<<M, kNormThreads, 0, stream>>>(tp, resid, w, xnorm, h, eps, resid_bf, sumsq0);
    ASD_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
}

// reduce K-split slices into slice 0 (used before a tensor-parallel all-reduce)
__global__ void reduce_slices_kernel(float* __restrict__ part, int nslices, size_t slice_stride, size_t n,
                                     float* __restrict__ dst) {
    grid_dep_wait();
    grid_dep_launch();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i * 4 >= n) return;
    float4 a = reinterpret_cast<float4*>(part)[i];
    for (int s = 1; s < nslices; ++s) {
        const float4 p = reinterpret_cast<const float4*>(part + s * slice_stride)[i];
        a.x += p.x;
        a.y += p.y;
        a.z += p.z;
        a.w += p.w;
    }
    reinterpret_cast<float4*>(dst)[i] = a;
}

int launch_reduce_slices(float* part, int nslices, size_t slice_stride, size_t n, cudaStream_t stream, float* dst) {
    if (dst == nullptr) dst = part;
    if ((nslices <= 1 && dst == part) || n == 0) return 0;
    const int threads = 256;
    const size_t blocks = (n / 4 + threads - 1) / threads;
    reduce_slices_kernel<<<(unsigned)blocks, threads, 0, stream>>>(part, nslices, slice_stride, n, dst);
    ASD_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
}

// one CTA per token: bias, RoPE, q store, paged K/V append
__global__ void __launch_bounds__(256) qkv_rope_kernel(const float* __restrict__ part, int nslices, size_t slice_stride,
                                                        const __nv_bfloat16* __restrict__ bias,
                                                        const int* __restrict__ positions,
                                                        const int* __restrict__ token_slot,
                                                        const int* __restrict__ page_table, int max_pages,
                                                        const float* __restrict__ inv_freq,
                                                        __nv_bfloat16* __restrict__ q_out,
                                                        __nv_bfloat16* __restrict__ k_cache,
                                                        __nv_bfloat16* __restrict__ v_cache, int nh, int nkv, int hd,
                                                        int page_size) {
    extern __shared__ float cs[];  // cos[hd/2], sin[hd/2]
    grid_dep_launch();
    const int m = blockIdx.x, half = hd >> 1;
    const int pos = positions[m];
    const int nqkv = (nh + 2 * nkv) * hd;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        float s, c;
        sincosf((float)pos * inv_freq[i], &s, &c);
        cs[i] = c;
        cs[half + i] = s;
    }
    __syncthreads();
    const int slot = token_slot[m];
    const int page = page_table[(size_t)slot * max_pages + pos / page_size];
    const int in_page = pos % page_size;
    const float* row = part + (size_t)m * nqkv;
    auto val = [&](int c) {
        float x = row[c];
        for (int s = 1; s < nslices; ++s) x += row[s * slice_stride + c];
        return x + __bfloat162float(bias[c]);
    };
    // rotary pairs of q and k heads
    const int npairs = (nh + nkv) * half;
    for (int t = threadIdx.x; t < npairs; t += blockDim.x) {
        const int head = t / half, i = t - head * half;
        const float x1 = val(head * hd + i), x2 = val(head * hd + i + half);
        const float c = cs[i], s = cs[half + i];
        const __nv_bfloat16 o1 = __float2bfloat16(x1 * c - x2 * s), o2 = __float2bfloat16(x2 * c + x1 * s);
        if (head < nh) {
            __nv_bfloat16* q = q_out + (size_t)m * nh * hd + head * hd;
            q[i] = o1;
            q[i + half] = o2;
        } else {
            const int g = head - nh;
            __nv_bfloat16* k = k_cache + (((size_t)page * nkv + g) * page_size + in_page) * hd;
            k[i] = o1;
            k[i + half] = o2;
        }
    }
    const int nv = nkv * hd;
    for (int t = threadIdx.x; t < nv; t += blockDim.x) {
        const int g = t / hd, d = t - g * hd;
        v_cache[(((size_t)page * nkv + g) * page_size + in_page) * hd + d] = __float2bfloat16(val((nh + nkv) * hd + t));
    }
}

int launch_qkv_rope(const float* part, int nslices, size_t slice_stride, const __nv_bfloat16* bias,
                    const int* positions, const int* token_slot, const int* page_table, int max_pages,
                    const float* inv_freq, __nv_bfloat16* q_out, __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, int M,
                    int nh, int nkv, int hd, int page_size, cudaStream_t stream) {
    if (M <= 0) return 0;
    qkv_rope_kernel<<<M, 256, hd * sizeof(float), stream>>>(part, nslices, slice_stride, bias, positions, token_slot,
                                                             page_table, max_pages, inv_freq, q_out, k_cache, v_cache,
                                                             nh, nkv, hd, page_size);
    ASD_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
}

// (cos, sin) of every token's rotary angles, once per forward (consumed by the fused QKV epilogue)
__global__ void rope_table_kernel(const int* __restrict__ positions, const float* __restrict__ inv_freq,
                                  float2* __restrict__ cs, int half) {
    const int m = blockIdx.x;
    const float pos = (float)positions[m];
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        float s, c;
        sincosf(pos * inv_freq[i], &s, &c);
        cs[(size_t)m * half + i] = make_float2(c, s);
    }
}

int launch_rope_table(const int* positions, const float* inv_freq, float2* cs, int M, int half, cudaStream_t stream) {
    if (M <= 0) return 0;
    glue_carveout();
    rope_table_kernel<<<M, 64, 0, stream>>>(positions, inv_freq, cs, half);
    ASD_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
}

__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, const int* __restrict__ rows,
                                   __nv_bfloat16* __restrict__ dst, int h, const float* __restrict__ ss_src,
                                   float* __restrict__ ss_dst, int parts, int ld) {
    const int r = blockIdx.x, sr = rows[r];
    if (ss_src != nullptr)
        for (int t = threadIdx.x; t < parts; t += blockDim.x) ss_dst[(size_t)t * ld + r] = ss_src[(size_t)t * ld + sr];
    const uint4* s = reinterpret_cast<const uint4*>(src + (size_t)sr * h);
    uint4* d = reinterpret_cast<uint4*>(dst + (size_t)r * h);
    for (int i = threadIdx.x; i < h / 8; i += blockDim.x) d[i] = s[i];
}

int launch_gather_rows(const __nv_bfloat16* src, const int* rows, __nv_bfloat16* dst, int n, int h,
                       cudaStream_t stream, const float* ss_src, float* ss_dst, int parts, int ld) {
    if (n <= 0) return 0;
    glue_carveout();
    gather_rows_kernel<<<n, 128, 0, stream>>>(src, rows, dst, h, ss_src, ss_dst, parts, ld);
    ASD_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
}

static void glue_carveout() {
    static PerDeviceOnce once;
    if (!once.need()) return;
    prefer_max_smem(add_norm_kernel);
    prefer_max_smem(tp_allreduce_norm_kernel);
    prefer_max_smem(reduce_slices_kernel);
    prefer_max_smem(qkv_rope_kernel);
    prefer_max_smem(rope_table_kernel);
    prefer_max_smem(gather_rows_kernel);
}

// Force the (lazily loaded) kernels of this file into the context now: a first launch that loads a kernel may need a
// context synchronisation, which deadlocks when another rank of the same process is spinning for this rank's launch.
int preload_layers() {
    cudaFuncAttributes fa;
    ASD_CUDA(cudaFuncGetAttributes(&fa, add_norm_kernel));
    ASD_CUDA(cudaFuncGetAttributes(&fa, tp_allreduce_norm_kernel));
    ASD_CUDA(cudaFuncGetAttributes(&fa, reduce_slices_kernel));
    ASD_CUDA(cudaFuncGetAttributes(&fa, qkv_rope_kernel));
    ASD_CUDA(cudaFuncGetAttributes(&fa, rope_table_kernel));
    ASD_CUDA(cudaFuncGetAttributes(&fa, gather_rows_kernel));
    return 0;
}

}  // namespace asd
