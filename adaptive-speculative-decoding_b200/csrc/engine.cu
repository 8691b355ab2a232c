// Qwen2-family forward engine: one opaque handle per (model, tensor-parallel rank).
// asd_engine_forward runs embedding -> L x [RMSNorm, QKV GEMM, RoPE + paged-KV append, attention,
// O GEMM, add+RMSNorm, gate|up GEMM with fused SwiGLU, down GEMM, add+RMSNorm] -> lm_head GEMM on
// the caller's stream with no host synchronisation (CUDA-graph capturable).
//
// This is the slot the reference fills with vllm.LLM(...).generate
// (/root/reference/src/serving/real_model_pipeline.py:98-108,135; Stage listing in
// docs/guides/RESEARCH_PROTOCOL.md:233-304).  Tensor parallelism mirrors what vLLM does for the
// reference's `tensor_parallel_size` (configs/qwen3_models.yaml:10,22,34,46): column-parallel
// QKV / gate|up, row-parallel O / down, one all-reduce after each row-parallel GEMM.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <string.h>

#include <unordered_map>
#include <vector>

#include "../../include/asd_b200.h"
#include "asd_internal.h"
#include "gemm.h"
#include "layers.h"
#include "persist.h"

namespace asd {

typedef int (*allreduce_fn_t)(const void* send, void* recv, size_t count, int dtype, int op, void* comm,
                              cudaStream_t stream);

struct Layer {
    const __nv_bfloat16 *wqkv = nullptr, *bqkv = nullptr, *wo = nullptr, *wgu = nullptr, *wdown = nullptr,
                        *ln1 = nullptr, *ln2 = nullptr;
    CUtensorMap t_qkv, t_o, t_gu, t_down;
};

struct Plans {
    GemmPlan qkv, o, gu, down;
    CUtensorMap x_norm, x_attn, x_act;
    // token counts above the HBM/tensor ridge: gate|up and down run on the tensor-bound CTA-pair kernel (gemm_tc.cu)
    bool use_tc = false;
    TcPlan gu_tc, down_tc, qkv_tc, o_tc;
    CUtensorMap x_norm_tc, x_act_tc, x_norm_qkv_tc, x_attn_tc;
};

struct Engine {
    asd_model_config c;
    int nqkv, qdim, ffp;  // local dims
    std::vector<Layer> layers;
    const __nv_bfloat16 *embed = nullptr, *final_norm = nullptr, *lm_head = nullptr;
    const float* inv_freq = nullptr;
    CUtensorMap t_lm;
    CUtensorMap kv_map;          // whole paged pool as rows of head_dim elements (tcgen05 attention)
    bool kv_map_ok = false;
    __nv_bfloat16* kv_pool = nullptr;
    size_t kv_half = 0;  // elements of one K (or V) pool of one layer
    const int* page_table = nullptr;
    int num_pages = 0, max_seqs = 0, max_pages = 0;
    // owned activations
    float *resid = nullptr, *part = nullptr, *o_part = nullptr, *ml_part = nullptr;
    int* tickets = nullptr;
    float2* rope_cs = nullptr;
    __nv_bfloat16* resid_bf = nullptr;
    float *sumsq = nullptr, *sumsq_sel = nullptr;
    __nv_bfloat16 *xnorm = nullptr, *q = nullptr, *attn = nullptr, *act = nullptr, *xsel = nullptr;
    size_t part_floats = 0, o_part_floats = 0;
    std::unordered_map<int, Plans> plans;
    std::unordered_map<int, std::pair<GemmPlan, CUtensorMap>> lm_plans;  // keyed by rows (xnorm) / -rows (xsel)
    std::unordered_map<int, std::pair<TcPlan, CUtensorMap>> lm_plans_tc;
    // options
    // attn_impl: 2 = tcgen05 kernel (attention_tc.cu; default: faster on every bench shape since P moved into TMEM -
    // 18.7 vs 22.8 us at 32B/prefix 512, 64 vs 180 us at 72B-TP4/prefix 4096, tools/bench_attn.py), 1 = mma.sync kernel
    // (also the fallback for head_dim 64 or more than 128 query rows), 0 = one-warp cross-check kernel
    int attn_impl = 2, pdl = 1, force_ksplit = 0, force_stages = 0, max_attn_splits = 16, reduce = 1,
        attn_target_ctas = 148, fuse_rope = 1, fuse_norm = 0, attn_min_split_keys = 1024, tp_fused = 1, tp_two_shot = 0;
    Tuning tune;   // per-engine knobs, installed for the calling thread by forward()
    // persistent forward kernel (forward_persist.cu): one cooperative launch per forward when the shape allows
    // measured on B200 (tools/trace_persist.py): the persistent kernel streams every GEMM at 87-95 % of HBM speed, but a
    // phase boundary through global memory (partials -> flags -> reduce -> epilogue -> counter -> X load) costs 15-18 us
    // against 10-12 us for a kernel boundary under programmatic dependent launch, so it is opt-in (option persist = 1)
    int persist = 0, persist_ahead = 0;
    // MB of the O-projection / gate|up weights that the attention kernel pulls into L2 while it runs (0 = off)
    int attn_prefetch_mb = 0;
    // token counts above 256: QKV and O also run on the CTA-pair kernel (0 = weight-streaming kernel as for small steps)
    int tc_qkvo = 1;
    PLayer* p_layers = nullptr;      // device array, rebuilt lazily after set_layer / set_kv
    bool p_layers_ok = false;
    unsigned* p_sync = nullptr;
    float* p_ws = nullptr;
    unsigned long long* p_trace = nullptr;   // optional (asd_debug_persist_trace)
    int device = 0;
    void* comm = nullptr;
    allreduce_fn_t allreduce = nullptr;
    // fused peer-memory all-reduce (CUDA IPC): double-buffered partials + flag array, local and peer views
    float* tp_buf[2] = {nullptr, nullptr};
    uint32_t* tp_flags = nullptr;
    int* tp_error = nullptr;
    const float* peer_buf[2][8] = {};
    uint32_t* peer_flags[8] = {};
    bool p2p = false;
    uint32_t tp_epoch = 0;
    // profiling (option "profile"): CUDA-event pairs around every launch, by kernel class
    int profile = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    std::vector<std::pair<int, int>> ev_used;  // (class, pool index)
    size_t ev_next = 0;
};

enum ProfClass { PROF_GEMM = 0, PROF_ATTN = 1, PROF_GLUE = 2, PROF_COMM = 3, PROF_LMHEAD = 4, PROF_NCLASS = 5 };

struct ProfScope {
    Engine* e;
    cudaStream_t s;
    int idx = -1;
    ProfScope(Engine* e_, int cls, cudaStream_t s_) : e(e_), s(s_) {
        if (!e->profile) return;
        if (e->ev_next == e->ev_pool.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            e->ev_pool.emplace_back(a, b);
        }
        idx = (int)e->ev_next++;
        e->ev_used.emplace_back(cls, idx);
        cudaEventRecord(e->ev_pool[idx].first, s);
    }
    ~ProfScope() {
        if (idx >= 0) cudaEventRecord(e->ev_pool[idx].second, s);
    }
};
#define PROF(cls) ProfScope _prof_scope(e, cls, s)

static int ensure_plans(Engine* e, int M, Plans** out) {
    auto it = e->plans.find(M);
    if (it != e->plans.end()) {
        *out = &it->second;
        return 0;
    }
    Plans p;
    const int h = e->c.hidden;
    if (gemm_plan(&p.qkv, M, e->nqkv, h, (e->reduce && e->fuse_rope) ? GEMM_OUT_QKV : GEMM_OUT_F32, e->force_ksplit,
                  e->force_stages, e->reduce))
        return -1;
    if (gemm_plan(&p.o, M, h, e->qdim, GEMM_OUT_F32, e->force_ksplit, e->force_stages, e->reduce)) return -1;
    if (gemm_plan(&p.gu, M, 2 * e->ffp, h, GEMM_OUT_SWIGLU, 0, e->force_stages, 0)) return -1;
    if (gemm_plan(&p.down, M, h, e->c.ffn, GEMM_OUT_F32, e->force_ksplit, e->force_stages, e->reduce)) return -1;
    const size_t need = (size_t)M * std::max(std::max((size_t)p.qkv.ksplit * e->nqkv, (size_t)p.o.ksplit * h),
                                             (size_t)p.down.ksplit * h);
    if (need > e->part_floats) return set_error("engine: split-K workspace too small for M = %d", M);
    if (make_tmap_bf16(&p.x_norm, e->fuse_norm ? e->resid_bf : e->xnorm, M, h, h, p.qkv.MT)) return -1;
    if (make_tmap_bf16(&p.x_attn, e->attn, M, e->qdim, e->qdim, p.o.MT)) return -1;
    if (make_tmap_bf16(&p.x_act, e->act, M, e->c.ffn, e->c.ffn, p.down.MT)) return -1;
    p.use_tc = e->tune.gemm_big && M > 256;
    if (p.use_tc) {
        if (gemm_tc_plan(&p.gu_tc, M, 2 * e->ffp, h, GEMM_OUT_SWIGLU, 0, e->force_stages)) return -1;
        if (gemm_tc_plan(&p.down_tc, M, h, e->c.ffn, GEMM_OUT_F32, 0, e->force_stages)) return -1;
        if ((size_t)p.down_tc.ksplit * M * h > e->part_floats) return set_error("engine: split-K workspace too small");
        if (make_tmap_bf16(&p.x_norm_tc, e->fuse_norm ? e->resid_bf : e->xnorm, M, h, h, p.gu_tc.MT / 2)) return -1;
        if (make_tmap_bf16(&p.x_act_tc, e->act, M, e->c.ffn, e->c.ffn, p.down_tc.MT / 2)) return -1;
        // QKV and O as well: at 576 tokens the weight-streaming kernel spends 59 / 120 us per layer on them (72B, TP4
        // shard; its per-thread epilogue loops are serial over the tokens), the CTA-pair kernel runs them at ~900 TFLOP/s
        if (gemm_tc_plan(&p.qkv_tc, M, e->nqkv, h, GEMM_OUT_F32, 0, e->force_stages)) return -1;
        if (gemm_tc_plan(&p.o_tc, M, h, e->qdim, GEMM_OUT_F32, 0, e->force_stages)) return -1;
        if ((size_t)p.qkv_tc.ksplit * M * e->nqkv > e->part_floats || (size_t)p.o_tc.ksplit * M * h > e->part_floats)
            return set_error("engine: split-K workspace too small");
        if (make_tmap_bf16(&p.x_norm_qkv_tc, e->fuse_norm ? e->resid_bf : e->xnorm, M, h, h, p.qkv_tc.MT / 2)) return -1;
        if (make_tmap_bf16(&p.x_attn_tc, e->attn, M, e->qdim, e->qdim, p.o_tc.MT / 2)) return -1;
    }
    auto r = e->plans.emplace(M, p);
    *out = &r.first->second;
    return 0;
}

static int tp_allreduce(Engine* e, float* buf, int nslices, size_t slice_stride, size_t n, cudaStream_t s) {
    if (launch_reduce_slices(buf, nslices, slice_stride, n, s)) return -1;
    const int rc = e->allreduce(buf, buf, n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, e->comm, s);
    if (rc != 0) return set_error("engine: all-reduce failed with ncclResult %d", rc);
    return 0;
}

static int forward(Engine* e, const int* tokens, const int* positions, const int* token_slot, int M, const int* cu_q,
                   const int* seq_slot, int nseq, int max_qlen, int max_kv_len, const int* logit_rows,
                   int n_logit_rows, float* logits_out, long long logits_ld, cudaStream_t s) {
    const asd_model_config& c = e->c;
    if (M <= 0) return 0;
    TuningScope tuning_scope(&e->tune);
    if (M > c.max_tokens) return set_error("engine: M = %d exceeds max_tokens = %d", M, c.max_tokens);
    if (!e->embed || !e->kv_pool || !e->inv_freq) return set_error("engine: weights / KV pool not set");
    for (auto& L : e->layers)
        if (!L.wqkv) return set_error("engine: a layer has no weights");
    const bool fn = e->fuse_norm != 0;   // RMSNorm fused into the GEMM epilogues (producer: bf16(resid * ln_w) + sum of squares; consumer: rstd)
    if (fn && !(e->reduce && e->fuse_rope)) return set_error("engine: fuse_norm needs reduce = 1 and fuse_rope = 1");
    Plans* P = nullptr;
    if (ensure_plans(e, M, &P)) return -1;
    const int h = c.hidden, nh = c.n_heads, nkv = c.n_kv_heads, hd = c.head_dim, Mx = c.max_tokens;
    const bool tp = c.tp_size > 1;
    if (tp && !e->allreduce && !e->p2p) return set_error("engine: tp_size > 1 but no all-reduce installed");

    // attention split selection: fill the SMs once, bounded workspace.  A split costs a ticket plus a merge
    // pass (~4 us, more than ten 64-key tiles of the key loop), so a sequence is only split into pieces of
    // at least attn_min_split_keys keys
    int nsplit = (e->attn_target_ctas + nseq * nkv / 2) / (nseq * nkv);
    const int max_by_len = max_kv_len / e->attn_min_split_keys > 0 ? max_kv_len / e->attn_min_split_keys : 1;
    if (nsplit > max_by_len) nsplit = max_by_len;
    if (nsplit > e->max_attn_splits) nsplit = e->max_attn_splits;
    if (nsplit < 1) nsplit = 1;
    int split_keys = ((max_kv_len + nsplit - 1) / nsplit + 127) / 128 * 128;   // whole key tiles of both kernels
    if (split_keys < 128) split_keys = 128;
    const int nsplit_max = (max_kv_len + split_keys - 1) / split_keys;
    if ((size_t)M * nh * nsplit_max * hd > e->o_part_floats) return set_error("engine: attention workspace too small");

    if (e->persist && !tp && fn && e->p_sync != nullptr &&
        persist_supported(M, n_logit_rows, max_qlen, nh, nkv, hd, c.page_size, h, c.n_layers)) {
        if (n_logit_rows > 0 && !logits_out) return set_error("engine: logits_out is NULL");
        if (logit_rows == nullptr && n_logit_rows > 0 && n_logit_rows != M)
            return set_error("engine: n_logit_rows must equal M when logit_rows is NULL");
        if (!e->p_layers_ok) {
            std::vector<PLayer> hl(c.n_layers);
            for (int l = 0; l < c.n_layers; ++l) {
                const Layer& Ly = e->layers[l];
                memset(&hl[l], 0, sizeof(PLayer));
                hl[l].t_qkv = Ly.t_qkv, hl[l].t_o = Ly.t_o, hl[l].t_gu = Ly.t_gu, hl[l].t_down = Ly.t_down;
                hl[l].bqkv = Ly.bqkv, hl[l].ln1 = Ly.ln1, hl[l].ln2 = Ly.ln2;
                hl[l].k_cache = e->kv_pool + (size_t)(2 * l) * e->kv_half;
                hl[l].v_cache = hl[l].k_cache + e->kv_half;
            }
            // stream-ordered after every forward already enqueued on this stream; a forward on another stream must
            // not be in flight while weights / the KV pool are being replaced (same rule as for the weights themselves)
            ASD_CUDA(cudaMemcpyAsync(e->p_layers, hl.data(), sizeof(PLayer) * c.n_layers, cudaMemcpyHostToDevice, s));
            ASD_CUDA(cudaStreamSynchronize(s));
            e->p_layers_ok = true;
        }
        PersistLaunch PL;
        PL.layers = e->p_layers;
        PL.n_layers = c.n_layers, PL.h = h, PL.nqkv = e->nqkv, PL.qdim = e->qdim, PL.ffn = c.ffn, PL.ffp = e->ffp;
        PL.vocab = c.vocab, PL.nh = nh, PL.nkv = nkv, PL.hd = hd, PL.page_size = c.page_size, PL.max_pages = e->max_pages;
        PL.Mx = Mx, PL.eps = c.rms_eps;
        PL.embed = e->embed, PL.final_norm = e->final_norm, PL.lm_head = e->lm_head, PL.inv_freq = e->inv_freq;
        PL.page_table = e->page_table;
        PL.M = M, PL.nseq = nseq, PL.max_qlen = max_qlen, PL.max_kv_len = max_kv_len;
        PL.tokens = tokens, PL.positions = positions, PL.token_slot = token_slot, PL.cu_q = cu_q, PL.seq_slot = seq_slot;
        PL.resid = e->resid, PL.resid_bf = e->resid_bf, PL.sumsq = e->sumsq, PL.sumsq_sel = e->sumsq_sel;
        PL.rope_cs = e->rope_cs, PL.q = e->q, PL.attn = e->attn, PL.act = e->act, PL.xsel = e->xsel;
        PL.split_keys = split_keys, PL.nsplit_max = nsplit_max;
        PL.o_part = e->o_part, PL.ml_part = e->ml_part, PL.tickets = e->tickets;
        PL.logit_rows = logit_rows, PL.n_logit_rows = n_logit_rows, PL.logits = logits_out;
        PL.logits_ld = logits_ld > 0 ? logits_ld : c.vocab;
        PL.prefetch_ahead = e->persist_ahead;
        PL.sync = e->p_sync, PL.part_ws = e->p_ws, PL.error = e->tp_error, PL.trace = e->p_trace;
        PROF(PROF_GEMM);
        return persist_launch(PL, s);
    }
    int parts = 1;   // rows of e->sumsq that currently describe the residual (fused-norm mode)
    NormFusion cons;  // consumer-side descriptor, refreshed before each consumer GEMM
    auto consumer = [&](const float* ss) -> const NormFusion* {
        if (!fn) return nullptr;
        cons = NormFusion{};
        cons.sumsq_in = ss;
        cons.parts = parts;
        cons.ld = Mx;
        cons.hidden = h;
        cons.eps = c.rms_eps;
        return &cons;
    };
    {
        PROF(PROF_GLUE);
        if (launch_add_norm(e->resid, nullptr, 0, 0, tokens, e->embed, e->layers[0].ln1, fn ? nullptr : e->xnorm, M, h,
                            c.rms_eps, s, fn ? e->resid_bf : nullptr, fn ? e->sumsq : nullptr))
            return -1;
        if (P->qkv.mode == GEMM_OUT_QKV && launch_rope_table(positions, e->inv_freq, e->rope_cs, M, hd / 2, s)) return -1;
    }
    // residual update after a row-parallel projection: fused into the GEMM when possible, else glue kernel
    // one kernel: all-reduce of the partials in tp_buf[epoch & 1] over NVLink peer memory + residual add + norm stats
    auto peer_allreduce = [&](const __nv_bfloat16* next_ln) -> int {
        PROF(PROF_COMM);
        parts = 1;
        const uint32_t ep = e->tp_epoch;   // the GEMM above wrote tp_buf[ep & 1]
        // two-shot (option tp_two_shot = 1): a row is reduced by its home rank and fetched once by the others, so a
        // rank moves 2 (W-1)/W payloads instead of W-1 (TP8, 96 tokens: 3.4 MB instead of 13.8 MB per boundary).
        // Measured on 8 B200: 12.4 ms per verify forward against 11.1 ms one-shot - at these sizes the exchange is
        // bound by the second flag hop (fence.sys + release), not by bytes - so one-shot stays the default.
        // It pays once bytes dominate: saved traffic (W-1 - 2 (W-1)/W) * M * hidden * 4 >= 16 MB per boundary
        // (72B, 576 tokens, TP4: 56 -> 28 MB per boundary).  tp_two_shot = 1 forces it, -1 disables it.
        const double payload = (double)M * h * 4.0, w1 = c.tp_size - 1;
        const bool two = (e->tp_two_shot == 1 && M <= kTpRowFlags) ||
                         (e->tp_two_shot == 0 && M <= kTpRowFlags &&
                          (w1 - 2.0 * w1 / c.tp_size) * payload >= 16.0 * (1 << 20));
        const float* bc[8] = {};
        uint32_t* rf[8] = {};
        for (int r = 0; r < c.tp_size; ++r) {
            bc[r] = e->peer_buf[ep & 1][r] + 2 * Mx * (size_t)h;   // slot 1 of the receive buffer: the final rows
            rf[r] = e->peer_flags[r] + 64;
        }
        return launch_tp_allreduce_norm(e->peer_buf[ep & 1], e->peer_flags, c.tp_rank, c.tp_size, ep, e->tp_error,
                                        e->resid, next_ln, fn ? nullptr : (next_ln ? e->xnorm : nullptr), M, h,
                                        c.rms_eps, fn ? e->resid_bf : nullptr, fn ? e->sumsq : nullptr, s,
                                        two ? bc : nullptr, two ? rf : nullptr);
    };
    // row-parallel projection on the tensor-bound kernel: fp32 K-split slices, summed in order by the glue kernels
    auto project_residual_tc = [&](const TcPlan& pl, const CUtensorMap& tw, const CUtensorMap& tx,
                                   const __nv_bfloat16* next_ln) -> int {
        const bool p2p_ok = tp && e->p2p;
        const size_t stride = (size_t)M * h;
        if (p2p_ok) ++e->tp_epoch;
        float* dst = (p2p_ok && pl.ksplit == 1) ? e->tp_buf[e->tp_epoch & 1] : e->part;
        {
            PROF(PROF_GEMM);
            if (gemm_tc_launch(pl, tw, tx, dst, h, h, stride, e->pdl, s, false, nullptr)) return -1;
        }
        if (p2p_ok) {
            if (pl.ksplit > 1) {
                PROF(PROF_GLUE);
                if (launch_reduce_slices(e->part, pl.ksplit, stride, stride, s, e->tp_buf[e->tp_epoch & 1])) return -1;
            }
            return peer_allreduce(next_ln);
        }
        int ns = pl.ksplit;
        if (tp) {
            PROF(PROF_COMM);
            if (tp_allreduce(e, e->part, ns, stride, stride, s)) return -1;
            ns = 1;
        }
        PROF(PROF_GLUE);
        parts = 1;
        return launch_add_norm(e->resid, e->part, ns, stride, nullptr, nullptr, next_ln,
                               fn ? nullptr : (next_ln ? e->xnorm : nullptr), M, h, c.rms_eps, s,
                               fn ? e->resid_bf : nullptr, fn ? e->sumsq : nullptr);
    };
    auto project_residual = [&](const GemmPlan& pl, const CUtensorMap& tw, const CUtensorMap& tx,
                                const __nv_bfloat16* next_ln) -> int {
        const bool fuse = !tp && (pl.reduce || pl.ksplit == 1);
        const bool pow2 = pl.ksplit >= 2 && (pl.ksplit & (pl.ksplit - 1)) == 0;
        const bool emit = fn && fuse && pl.reduce && pow2;
        const bool p2p_ok = tp && e->p2p && (pl.reduce || pl.ksplit == 1);
        if (p2p_ok) ++e->tp_epoch;
        // The fused exchange is a one-shot all-reduce whose words carry their own flag (2 x the bytes): measured on
        // B200 it beats the separate all-reduce kernel while (world - 1) * M * hidden * 8 bytes stays small (TP2:
        // verify 11.6 -> 10.3 ms); at TP4 and 96 tokens the 12 MB per GEMM saturate the links (12.4 vs 9.8 ms), so
        // larger exchanges keep the pull-based kernel.  tp_fused = 2 forces the fused path.
        const size_t push_bytes = (size_t)(c.tp_size - 1) * M * h * 8;
        const bool fused_pays = e->tp_fused == 2 || c.tp_size == 2 || push_bytes <= (size_t)5 << 20;
        if (p2p_ok && e->tp_fused && fused_pays && fn && pl.reduce && pow2) {
            // tensor parallel, fused: the GEMM's owner CTAs exchange their partial tiles over NVLink peer memory
            // and finish residual + norm statistics themselves - same five launches per layer as on one GPU
            PROF(PROF_GEMM);
            const uint32_t ep = e->tp_epoch;
            TpFusion tf;
            for (int r = 0; r < c.tp_size; ++r) tf.recv[r] = const_cast<float*>(e->peer_buf[ep & 1][r]);
            tf.rank = c.tp_rank;
            tf.world = c.tp_size;
            tf.epoch = ep;
            tf.error = e->tp_error;
            tf.slot_stride = Mx * (size_t)h;
            NormFusion prod;
            prod.sumsq_out = e->sumsq;
            prod.ld = Mx;
            prod.resid_bf = e->resid_bf;
            prod.ln_w = next_ln;
            if (gemm_launch(pl, tw, tx, e->resid, h, h, e->pdl, s, true, nullptr, &prod, &tf)) return -1;
            parts = pl.n_tiles;
            return 0;
        }
        NormFusion prod;
        if (emit) {
            prod.sumsq_out = e->sumsq;
            prod.ld = Mx;
            prod.resid_bf = e->resid_bf;
            prod.ln_w = next_ln;
        }
        {
            PROF(PROF_GEMM);
            void* dst = fuse ? (void*)e->resid : (p2p_ok ? (void*)e->tp_buf[e->tp_epoch & 1] : (void*)e->part);
            if (gemm_launch(pl, tw, tx, dst, h, h, e->pdl, s, fuse, nullptr, emit ? &prod : nullptr))
                return -1;
        }
        if (emit) {
            parts = pl.n_tiles;
            return 0;
        }
        int ns = fuse ? 0 : (pl.reduce ? 1 : pl.ksplit);
        if (tp && p2p_ok) return peer_allreduce(next_ln);
        if (tp) {
            PROF(PROF_COMM);
            if (tp_allreduce(e, e->part, ns, (size_t)M * h, (size_t)M * h, s)) return -1;
            ns = 1;
        }
        PROF(PROF_GLUE);
        parts = 1;
        return launch_add_norm(e->resid, e->part, ns, (size_t)M * h, nullptr, nullptr, next_ln,
                               fn ? nullptr : (next_ln ? e->xnorm : nullptr), M, h, c.rms_eps, s,
                               fn ? e->resid_bf : nullptr, fn ? e->sumsq : nullptr);
    };
    for (int l = 0; l < c.n_layers; ++l) {
        Layer& L = e->layers[l];
        __nv_bfloat16* kc = e->kv_pool + (size_t)(2 * l) * e->kv_half;
        __nv_bfloat16* vc = kc + e->kv_half;
        if (P->use_tc && e->tune.gemm_big >= 1 && e->tc_qkvo) {   // tensor-bound step: CTA-pair GEMM + one glue kernel
            {
                PROF(PROF_GEMM);
                if (gemm_tc_launch(P->qkv_tc, L.t_qkv, P->x_norm_qkv_tc, e->part, e->nqkv, e->nqkv, (size_t)M * e->nqkv,
                                   e->pdl, s, false, consumer(e->sumsq)))
                    return -1;
            }
            PROF(PROF_GLUE);
            if (launch_qkv_rope(e->part, P->qkv_tc.ksplit, (size_t)M * e->nqkv, L.bqkv, positions, token_slot,
                                e->page_table, e->max_pages, e->inv_freq, e->q, kc, vc, M, nh, nkv, hd, c.page_size, s))
                return -1;
        } else if (P->qkv.mode == GEMM_OUT_QKV) {   // bias + RoPE + q store + paged K/V append fused into the epilogue
            PROF(PROF_GEMM);
            QkvEpilogue q;
            q.cs = e->rope_cs;
            q.bias = L.bqkv;
            q.positions = positions;
            q.token_slot = token_slot;
            q.page_table = e->page_table;
            q.q_out = e->q;
            q.k_cache = kc;
            q.v_cache = vc;
            q.max_pages = e->max_pages;
            q.nh = nh;
            q.nkv = nkv;
            q.hd = hd;
            q.page_size = c.page_size;
            gemm_set_next(P->o, &L.t_o);
            if (gemm_launch(P->qkv, L.t_qkv, P->x_norm, nullptr, e->nqkv, e->nqkv, e->pdl, s, false, &q,
                            consumer(e->sumsq)))
                return -1;
        } else {
            {
                PROF(PROF_GEMM);
                if (gemm_launch(P->qkv, L.t_qkv, P->x_norm, e->part, e->nqkv, e->nqkv, e->pdl, s)) return -1;
            }
            PROF(PROF_GLUE);
            if (launch_qkv_rope(e->part, P->qkv.reduce ? 1 : P->qkv.ksplit, (size_t)M * e->nqkv, L.bqkv, positions,
                                token_slot, e->page_table, e->max_pages, e->inv_freq, e->q, kc, vc, M, nh, nkv, hd,
                                c.page_size, s))
                return -1;
        }
        AttnLaunch A;
        A.q = e->q;
        A.k_cache = kc;
        A.v_cache = vc;
        A.positions = positions;
        A.token_slot = token_slot;
        A.cu_q = cu_q;
        A.seq_slot = seq_slot;
        A.page_table = e->page_table;
        A.out = e->attn;
        A.o_part = e->o_part;
        A.ml_part = e->ml_part;
        A.tickets = e->tickets;
        A.M = M;
        A.nseq = nseq;
        A.max_qlen = max_qlen;
        A.nh = nh;
        A.nkv = nkv;
        A.hd = hd;
        A.page_size = c.page_size;
        A.max_pages = e->max_pages;
        A.split_keys = split_keys;
        A.nsplit_max = nsplit_max;
        A.impl = e->attn_impl;
        A.kv_map = e->kv_map_ok ? &e->kv_map : nullptr;
        A.k_row0 = (long long)(2 * l) * e->num_pages * nkv * c.page_size;
        A.v_row0 = A.k_row0 + (long long)e->num_pages * nkv * c.page_size;
        if (e->attn_prefetch_mb > 0) {
            // all of W_o first, what is left of the budget as the first k-blocks of every gate|up CTA
            long long budget = (long long)e->attn_prefetch_mb << 20;
            const GemmPlan* gp[2] = {&P->o, P->use_tc ? nullptr : &P->gu};
            const CUtensorMap* gm[2] = {&L.t_o, &L.t_gu};
            for (int i = 0; i < 2; ++i) {
                if (gp[i] == nullptr || gp[i]->m_tiles != 1 || budget <= 0) continue;
                const long long ncta = (long long)gp[i]->n_tiles * gp[i]->ksplit;
                const int per_cta = (gp[i]->kblocks + gp[i]->ksplit - 1) / gp[i]->ksplit;
                int kp = (int)(budget / (ncta * 16384));
                if (kp > per_cta) kp = per_cta;
                if (kp <= 0) continue;
                A.pf[i].tmap = gm[i];
                A.pf[i].ntiles = gp[i]->n_tiles;
                A.pf[i].ksplit = gp[i]->ksplit;
                A.pf[i].kblocks = gp[i]->kblocks;
                A.pf[i].kp = kp;
                budget -= (long long)kp * ncta * 16384;
            }
        }
        {
            PROF(PROF_ATTN);
            if (launch_attention(A, s)) return -1;
        }
        gemm_set_next(P->gu, &L.t_gu);
        if (P->use_tc && e->tc_qkvo ? project_residual_tc(P->o_tc, L.t_o, P->x_attn_tc, L.ln2)
                                    : project_residual(P->o, L.t_o, P->x_attn, L.ln2))
            return -1;
        if (P->use_tc) {
            PROF(PROF_GEMM);
            if (gemm_tc_launch(P->gu_tc, L.t_gu, P->x_norm_tc, e->act, c.ffn, c.ffn, 0, e->pdl, s, false, consumer(e->sumsq)))
                return -1;
        } else {
            PROF(PROF_GEMM);
            gemm_set_next(P->down, &L.t_down);
            if (gemm_launch(P->gu, L.t_gu, P->x_norm, e->act, c.ffn, c.ffn, e->pdl, s, false, nullptr,
                            consumer(e->sumsq)))
                return -1;
        }
        const bool last = l + 1 == c.n_layers;
        const __nv_bfloat16* wn = last ? e->final_norm : e->layers[l + 1].ln1;
        if (!last) gemm_set_next(P->qkv, &e->layers[l + 1].t_qkv);
        const __nv_bfloat16* down_ln = (last && n_logit_rows == 0 && !fn) ? nullptr : wn;
        if (P->use_tc ? project_residual_tc(P->down_tc, L.t_down, P->x_act_tc, down_ln)
                      : project_residual(P->down, L.t_down, P->x_act, down_ln))
            return -1;
    }
    if (n_logit_rows <= 0) return 0;
    if (!logits_out) return set_error("engine: logits_out is NULL");
    int rows = M;
    const __nv_bfloat16* src = fn ? e->resid_bf : e->xnorm;
    const float* ss = e->sumsq;
    int key = M;
    if (logit_rows) {
        if (launch_gather_rows(src, logit_rows, e->xsel, n_logit_rows, h, s, fn ? e->sumsq : nullptr, e->sumsq_sel, parts,
                               Mx))
            return -1;
        rows = n_logit_rows;
        src = e->xsel;
        ss = e->sumsq_sel;
        key = -rows;
    } else if (n_logit_rows != M) {
        return set_error("engine: n_logit_rows must equal M when logit_rows is NULL");
    }
    if (e->tune.gemm_big && rows > 256) {   // tensor-bound kernel; 152K vocabulary rows give plenty of units without a K split
        auto jt = e->lm_plans_tc.find(key);
        if (jt == e->lm_plans_tc.end()) {
            std::pair<TcPlan, CUtensorMap> v;
            if (gemm_tc_plan(&v.first, rows, c.vocab, h, GEMM_OUT_F32, 1, e->force_stages)) return -1;
            if (make_tmap_bf16(&v.second, src, rows, h, h, v.first.MT / 2)) return -1;
            jt = e->lm_plans_tc.emplace(key, v).first;
        }
        PROF(PROF_LMHEAD);
        const int ld = logits_ld > 0 ? (int)logits_ld : c.vocab;
        return gemm_tc_launch(jt->second.first, e->t_lm, jt->second.second, logits_out, ld, c.vocab, 0, e->pdl, s, false,
                              consumer(ss));
    }
    auto it = e->lm_plans.find(key);
    if (it == e->lm_plans.end()) {
        std::pair<GemmPlan, CUtensorMap> v;
        if (gemm_plan(&v.first, rows, c.vocab, h, GEMM_OUT_F32, 1, e->force_stages, 0)) return -1;
        if (make_tmap_bf16(&v.second, src, rows, h, h, v.first.MT)) return -1;
        it = e->lm_plans.emplace(key, v).first;
    }
    PROF(PROF_LMHEAD);
    return gemm_launch(it->second.first, e->t_lm, it->second.second, logits_out,
                       logits_ld > 0 ? (int)logits_ld : c.vocab, c.vocab, e->pdl, s, false, nullptr, consumer(ss));
}

}  // namespace asd

using namespace asd;

extern "C" {

asd_engine_t* asd_engine_create(const asd_model_config* cfg) {
    if (!cfg) {
        set_error("asd_engine_create: NULL config");
        return nullptr;
    }
    const asd_model_config& c = *cfg;
    if (c.hidden <= 0 || c.hidden % 8 || c.hidden > 8192 || c.n_layers <= 0 || c.n_heads <= 0 || c.n_kv_heads <= 0 ||
        c.n_heads % c.n_kv_heads || (c.head_dim != 64 && c.head_dim != 128) || c.ffn <= 0 || c.ffn % 8 ||
        c.vocab <= 0 || c.max_tokens <= 0 || c.page_size <= 0 || c.tp_size < 1) {
        set_error("asd_engine_create: unsupported model shape");
        return nullptr;
    }
    {   // lazy module loading + spinning tensor-parallel kernels in one process = deadlock: load everything now
        static PerDeviceOnce loaded;
        if (loaded.need() && (preload_layers() || preload_attention() || preload_attention_tc() || preload_gemm() ||
                              preload_gemm_tc()))
            return nullptr;
    }
    Engine* e = new Engine();
    e->c = c;
    cudaGetDevice(&e->device);
    e->nqkv = (c.n_heads + 2 * c.n_kv_heads) * c.head_dim;
    e->qdim = c.n_heads * c.head_dim;
    e->ffp = (c.ffn + 63) / 64 * 64;
    e->layers.resize(c.n_layers);
    const size_t Mx = c.max_tokens;
    e->part_floats = 16 * Mx * (size_t)std::max(e->nqkv, c.hidden);
    e->o_part_floats = Mx * c.n_heads * (size_t)e->max_attn_splits * c.head_dim;
    bool ok = true;
    auto alloc = [&](void** p, size_t bytes) {
        if (ok && cudaMalloc(p, bytes) != cudaSuccess) {
            ok = false;
            set_error("asd_engine_create: cudaMalloc of %zu bytes failed", bytes);
        }
    };
    alloc((void**)&e->resid, Mx * c.hidden * 4);
    alloc((void**)&e->part, e->part_floats * 4);
    alloc((void**)&e->o_part, e->o_part_floats * 4);
    alloc((void**)&e->ml_part, Mx * c.n_heads * (size_t)e->max_attn_splits * 2 * 4);
    alloc((void**)&e->rope_cs, Mx * (c.head_dim / 2) * sizeof(float2));
    alloc((void**)&e->resid_bf, Mx * c.hidden * 2);
    alloc((void**)&e->sumsq, Mx * (size_t)((c.hidden + 127) / 128) * 4);
    alloc((void**)&e->sumsq_sel, Mx * (size_t)((c.hidden + 127) / 128) * 4);
    alloc((void**)&e->tickets, Mx * c.n_kv_heads * 4);
    if (ok && cudaMemset(e->tickets, 0, Mx * c.n_kv_heads * 4) != cudaSuccess) ok = false;
    if (c.tp_size > 1) {
        // receive buffers of the fused GEMM + all-reduce: one [Mx, hidden] fp32 slot per source rank (the separate
        // all-reduce kernel uses slot 0 for the partial and slot 1 for the two-shot final rows); flags: 64 words of
        // per-rank epochs + kTpRowFlags per-row epochs (two-shot)
        const size_t flag_bytes = 256 + (size_t)kTpRowFlags * 4;
        alloc((void**)&e->tp_buf[0], (size_t)c.tp_size * Mx * c.hidden * 8);   // {value, epoch} word pairs
        alloc((void**)&e->tp_buf[1], (size_t)c.tp_size * Mx * c.hidden * 8);
        if (ok) {
            cudaMemset(e->tp_buf[0], 0, (size_t)c.tp_size * Mx * c.hidden * 8);   // epoch 0 never matches a launch
            cudaMemset(e->tp_buf[1], 0, (size_t)c.tp_size * Mx * c.hidden * 8);
        }
        alloc((void**)&e->tp_flags, flag_bytes);
        alloc((void**)&e->tp_error, 256);
        if (ok) {
            cudaMemset(e->tp_flags, 0, flag_bytes);
            cudaMemset(e->tp_error, 0, 256);
        }
    }
    if (c.tp_size == 1) {
        alloc((void**)&e->tp_error, 256);       // device-side wait bound exceeded (persistent kernel)
        if (ok) cudaMemset(e->tp_error, 0, 256);
        alloc((void**)&e->p_layers, sizeof(PLayer) * c.n_layers);
        alloc((void**)&e->p_sync, sizeof(unsigned) * kPersistSyncWords);
        alloc((void**)&e->p_ws, persist_part_ws_bytes());
    }
    alloc((void**)&e->xnorm, Mx * c.hidden * 2);
    alloc((void**)&e->xsel, Mx * c.hidden * 2);
    alloc((void**)&e->q, Mx * e->qdim * 2);
    alloc((void**)&e->attn, Mx * e->qdim * 2);
    alloc((void**)&e->act, Mx * (size_t)c.ffn * 2);
    if (!ok) {
        asd_engine_destroy(reinterpret_cast<asd_engine_t*>(e));
        return nullptr;
    }
    return reinterpret_cast<asd_engine_t*>(e);
}

void asd_engine_destroy(asd_engine_t* h) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e) return;
    void* bufs[] = {e->resid, e->part, e->o_part, e->ml_part, e->tickets, e->rope_cs, e->resid_bf, e->sumsq,
                    e->sumsq_sel, e->tp_buf[0], e->tp_buf[1], e->tp_flags, e->tp_error, e->xnorm, e->xsel, e->q, e->attn, e->act,
                    e->p_layers, e->p_sync, e->p_ws};
    for (void* b : bufs)
        if (b) cudaFree(b);
    delete e;
}

int asd_engine_set_layer(asd_engine_t* h, int layer, const void* wqkv, const void* bqkv, const void* wo,
                         const void* wgateup, const void* wdown, const void* ln1, const void* ln2) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e || layer < 0 || layer >= e->c.n_layers) return set_error("asd_engine_set_layer: bad handle or layer");
    if (!wqkv || !bqkv || !wo || !wgateup || !wdown || !ln1 || !ln2) return set_error("asd_engine_set_layer: NULL weight");
    Layer& L = e->layers[layer];
    L.wqkv = (const __nv_bfloat16*)wqkv;
    L.bqkv = (const __nv_bfloat16*)bqkv;
    L.wo = (const __nv_bfloat16*)wo;
    L.wgu = (const __nv_bfloat16*)wgateup;
    L.wdown = (const __nv_bfloat16*)wdown;
    L.ln1 = (const __nv_bfloat16*)ln1;
    L.ln2 = (const __nv_bfloat16*)ln2;
    const int hdn = e->c.hidden;
    if (make_tmap_bf16(&L.t_qkv, wqkv, e->nqkv, hdn, hdn, 128)) return -1;
    if (make_tmap_bf16(&L.t_o, wo, hdn, e->qdim, e->qdim, 128)) return -1;
    if (make_tmap_bf16(&L.t_gu, wgateup, 2 * e->ffp, hdn, hdn, 128)) return -1;
    if (make_tmap_bf16(&L.t_down, wdown, hdn, e->c.ffn, e->c.ffn, 128)) return -1;
    e->p_layers_ok = false;
    return 0;
}

int asd_engine_set_globals(asd_engine_t* h, const void* embed, const void* final_norm, const void* lm_head,
                           const float* inv_freq) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e || !embed || !final_norm || !lm_head || !inv_freq) return set_error("asd_engine_set_globals: NULL argument");
    e->embed = (const __nv_bfloat16*)embed;
    e->final_norm = (const __nv_bfloat16*)final_norm;
    e->lm_head = (const __nv_bfloat16*)lm_head;
    e->inv_freq = inv_freq;
    return make_tmap_bf16(&e->t_lm, lm_head, e->c.vocab, e->c.hidden, e->c.hidden, 128);
}

int asd_engine_set_kv(asd_engine_t* h, void* kv_pool, int num_pages, const int32_t* page_table, int max_seqs,
                      int max_pages_per_seq) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e || !kv_pool || !page_table || num_pages <= 0) return set_error("asd_engine_set_kv: bad argument");
    e->kv_pool = (__nv_bfloat16*)kv_pool;
    e->num_pages = num_pages;
    e->page_table = page_table;
    e->max_seqs = max_seqs;
    e->max_pages = max_pages_per_seq;
    e->kv_half = (size_t)num_pages * e->c.n_kv_heads * e->c.page_size * e->c.head_dim;
    e->kv_map_ok = false;
    e->p_layers_ok = false;
    const unsigned long long rows = 2ull * e->c.n_layers * num_pages * e->c.n_kv_heads * e->c.page_size;
    const int ps = e->c.page_size;
    if (e->c.head_dim == 128 && (ps == 16 || ps == 32 || ps == 64 || ps == 128) && rows < (1ull << 31)) {
        if (make_tmap_bf16(&e->kv_map, kv_pool, rows, 128, 128, ps)) return -1;   // box = one page x 64 elements
        e->kv_map_ok = true;
    }
    return 0;
}

size_t asd_engine_kv_pool_bytes(const asd_model_config* c, int num_pages) {
    return (size_t)2 * c->n_layers * num_pages * c->n_kv_heads * c->page_size * c->head_dim * 2;
}

int asd_engine_set_allreduce(asd_engine_t* h, void* comm, void* nccl_allreduce_fn) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e) return set_error("asd_engine_set_allreduce: NULL handle");
    e->comm = comm;
    e->allreduce = reinterpret_cast<allreduce_fn_t>(nccl_allreduce_fn);
    return 0;
}

int asd_engine_ipc_export(asd_engine_t* h, void* handles_out) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e || !handles_out || !e->tp_buf[0]) return set_error("asd_engine_ipc_export: engine has no TP buffers");
    cudaIpcMemHandle_t* out = static_cast<cudaIpcMemHandle_t*>(handles_out);
    ASD_CUDA(cudaIpcGetMemHandle(&out[0], e->tp_buf[0]));
    ASD_CUDA(cudaIpcGetMemHandle(&out[1], e->tp_buf[1]));
    ASD_CUDA(cudaIpcGetMemHandle(&out[2], e->tp_flags));
    return 0;
}

int asd_engine_ipc_import(asd_engine_t* h, const void* all_handles) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e || !all_handles || !e->tp_buf[0]) return set_error("asd_engine_ipc_import: engine has no TP buffers");
    const int world = e->c.tp_size, rank = e->c.tp_rank;
    if (world > 8) return set_error("asd_engine_ipc_import: tp_size <= 8");
    const cudaIpcMemHandle_t* hs = static_cast<const cudaIpcMemHandle_t*>(all_handles);
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            e->peer_buf[0][r] = e->tp_buf[0];
            e->peer_buf[1][r] = e->tp_buf[1];
            e->peer_flags[r] = e->tp_flags;
            continue;
        }
        void *b0 = nullptr, *b1 = nullptr, *fl = nullptr;
        ASD_CUDA(cudaIpcOpenMemHandle(&b0, hs[r * 3 + 0], cudaIpcMemLazyEnablePeerAccess));
        ASD_CUDA(cudaIpcOpenMemHandle(&b1, hs[r * 3 + 1], cudaIpcMemLazyEnablePeerAccess));
        ASD_CUDA(cudaIpcOpenMemHandle(&fl, hs[r * 3 + 2], cudaIpcMemLazyEnablePeerAccess));
        e->peer_buf[0][r] = static_cast<const float*>(b0);
        e->peer_buf[1][r] = static_cast<const float*>(b1);
        e->peer_flags[r] = static_cast<uint32_t*>(fl);
    }
    e->p2p = true;
    return 0;
}

// In-process tensor parallelism: all ranks' engines live in ONE process on different devices (what
// Stage(tensor_parallel_size = t, gpu_ids = [...]) builds).  With peer access enabled a cudaMalloc pointer of one
// device is directly usable in kernels of the others (unified addressing), so no IPC handles are needed.
int asd_engine_peer_connect(asd_engine_t** hs, int n) {
    if (!hs || n < 1 || n > 8) return set_error("asd_engine_peer_connect: need 1..8 engines");
    Engine* es[8];
    for (int r = 0; r < n; ++r) {
        es[r] = reinterpret_cast<Engine*>(hs[r]);
        if (!es[r] || !es[r]->tp_buf[0] || es[r]->c.tp_size != n || es[r]->c.tp_rank != r)
            return set_error("asd_engine_peer_connect: engine %d is not rank %d of a %d-way group", r, r, n);
        for (int q = 0; q < r; ++q)
            if (es[q]->device == es[r]->device)
                return set_error("asd_engine_peer_connect: ranks %d and %d share device %d", q, r, es[r]->device);
    }
    int prev = 0;
    ASD_CUDA(cudaGetDevice(&prev));
    for (int r = 0; r < n; ++r) {
        ASD_CUDA(cudaSetDevice(es[r]->device));
        for (int q = 0; q < n; ++q) {
            if (q == r) continue;
            int can = 0;
            ASD_CUDA(cudaDeviceCanAccessPeer(&can, es[r]->device, es[q]->device));
            if (!can) {
                cudaSetDevice(prev);
                return set_error("asd_engine_peer_connect: device %d cannot access device %d", es[r]->device, es[q]->device);
            }
            const cudaError_t pe = cudaDeviceEnablePeerAccess(es[q]->device, 0);
            if (pe == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
            } else if (pe != cudaSuccess) {
                cudaSetDevice(prev);
                return set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", es[r]->device, es[q]->device,
                                 cudaGetErrorString(pe));
            }
        }
    }
    ASD_CUDA(cudaSetDevice(prev));
    for (int r = 0; r < n; ++r) {
        for (int q = 0; q < n; ++q) {
            es[r]->peer_buf[0][q] = es[q]->tp_buf[0];
            es[r]->peer_buf[1][q] = es[q]->tp_buf[1];
            es[r]->peer_flags[q] = es[q]->tp_flags;
        }
        es[r]->p2p = true;
        es[r]->plans.clear();
    }
    return 0;
}

int asd_engine_tp_error(asd_engine_t* h) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e || !e->tp_error) return 0;
    int v = 0, prev = 0;
    cudaGetDevice(&prev);
    if (prev != e->device) cudaSetDevice(e->device);
    const cudaError_t rc = cudaMemcpy(&v, e->tp_error, sizeof(int), cudaMemcpyDeviceToHost);
    if (prev != e->device) cudaSetDevice(prev);
    return rc == cudaSuccess ? v : -1;
}

int asd_engine_set_option(asd_engine_t* h, const char* name, int value) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e || !name) return set_error("asd_engine_set_option: NULL argument");
    if (!strcmp(name, "attn_impl")) e->attn_impl = value;
    else if (!strcmp(name, "pdl")) e->pdl = value;
    else if (!strcmp(name, "ksplit")) e->force_ksplit = value;
    else if (!strcmp(name, "stages")) e->force_stages = value;
    else if (!strcmp(name, "reduce")) e->reduce = value;
    else if (!strcmp(name, "attn_target_ctas")) e->attn_target_ctas = value;
    else if (!strcmp(name, "attn_min_split_keys")) e->attn_min_split_keys = value < 64 ? 64 : value;
    else if (!strcmp(name, "fuse_rope")) e->fuse_rope = value;
    else if (!strcmp(name, "glue_pdl")) e->tune.glue_pdl = value;
    else if (!strcmp(name, "attn_wide")) e->tune.attn_wide = value;
    else if (!strcmp(name, "attn_dbg")) e->tune.attn_dbg = value;
    else if (!strcmp(name, "tc_qkvo")) e->tc_qkvo = value;
    else if (!strcmp(name, "attn_prefetch_mb")) e->attn_prefetch_mb = value;
    else if (!strcmp(name, "gemm_big")) e->tune.gemm_big = value;
    else if (!strcmp(name, "persist")) e->persist = value;
    else if (!strcmp(name, "persist_ahead")) e->persist_ahead = value;
    else if (!strcmp(name, "fuse_norm")) e->fuse_norm = value;
    else if (!strcmp(name, "early_trigger")) e->tune.gemm_early_trigger = value;
    else if (!strcmp(name, "headroom")) e->tune.gemm_headroom = value;
    else if (!strcmp(name, "recv_dedicated")) e->tune.gemm_recv_dedicated = value;
    else if (!strcmp(name, "next_prefetch_mb")) e->tune.gemm_next_mb = value;
    else if (!strcmp(name, "tp_fused")) e->tp_fused = value;
    else if (!strcmp(name, "tp_two_shot")) e->tp_two_shot = value;
    else if (!strcmp(name, "p2p")) e->p2p = value != 0 && e->peer_flags[0] != nullptr;
    else if (!strcmp(name, "profile")) {
        e->profile = value;
        e->ev_used.clear();
        e->ev_next = 0;
        return 0;
    }
    else return set_error("asd_engine_set_option: unknown option %s", name);
    e->plans.clear();
    e->lm_plans.clear();
    e->lm_plans_tc.clear();
    return 0;
}

// diagnostics: per-CTA globaltimer stamps of the persistent forward kernel (buf: [CTAs][643] u64 on the device;
// NULL switches tracing off).  Slot 0: kernel entry; slot p + 1: this CTA's share of phase p done (phase 0 embedding,
// 1 + 5 l + {0 QKV, 1 attention, 2 O, 3 gate|up, 4 down} of layer l, then the logit-row gather and the lm_head).
int asd_engine_persist_trace(asd_engine_t* h, unsigned long long* buf) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e) return set_error("asd_engine_persist_trace: NULL handle");
    e->p_trace = buf;
    return persist_num_ctas();
}

int asd_engine_profile_read(asd_engine_t* h, float* ms_by_class, int* launches_by_class, int nclass) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e || !ms_by_class || nclass < PROF_NCLASS) return set_error("asd_engine_profile_read: bad argument");
    ASD_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < nclass; ++i) {
        ms_by_class[i] = 0.0f;
        if (launches_by_class) launches_by_class[i] = 0;
    }
    for (auto& u : e->ev_used) {
        float ms = 0.0f;
        ASD_CUDA(cudaEventElapsedTime(&ms, e->ev_pool[u.second].first, e->ev_pool[u.second].second));
        ms_by_class[u.first] += ms;
        if (launches_by_class) launches_by_class[u.first] += 1;
    }
    e->ev_used.clear();
    e->ev_next = 0;
    return 0;
}

int asd_engine_forward(asd_engine_t* h, const int32_t* tokens, const int32_t* positions, const int32_t* token_slot,
                       int M, const int32_t* cu_q, const int32_t* seq_slot, int nseq, int max_qlen, int max_kv_len,
                       const int32_t* logit_rows, int n_logit_rows, float* logits_out, long long logits_ld,
                       void* stream) {
    Engine* e = reinterpret_cast<Engine*>(h);
    if (!e) return set_error("asd_engine_forward: NULL handle");
    if (!tokens || !positions || !token_slot || !cu_q || !seq_slot || nseq <= 0 || max_qlen <= 0 || max_kv_len <= 0)
        return set_error("asd_engine_forward: bad argument");
    if ((long long)max_kv_len > (long long)e->max_pages * e->c.page_size)
        return set_error("asd_engine_forward: max_kv_len = %d exceeds the page table (%d pages of %d positions)",
                         max_kv_len, e->max_pages, e->c.page_size);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (e->c.tp_size > 1 && e->p2p) {
        // the peer-memory all-reduce numbers its exchanges with a host-side epoch baked into the kernel arguments:
        // a replayed graph would re-use stale epochs, so tensor-parallel forwards are not graph-capturable
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone)
            return set_error("asd_engine_forward: tensor-parallel forwards cannot be captured into a CUDA graph");
    }
    int prev = 0;
    ASD_CUDA(cudaGetDevice(&prev));
    if (prev != e->device) ASD_CUDA(cudaSetDevice(e->device));
    const int rc = forward(e, tokens, positions, token_slot, M, cu_q, seq_slot, nseq, max_qlen, max_kv_len, logit_rows,
                           n_logit_rows, logits_out, logits_ld, s);
    if (prev != e->device) cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
