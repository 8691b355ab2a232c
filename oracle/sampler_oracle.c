/*
 * ORACLE (test infrastructure only - never imported by the product path).
 *
 * PARITY UNPINNED: the reference contains no token-level rejection sampler (SURVEY.md
 * section 0 fact 2, section 8c), so there is no reference implementation, test or golden
 * vector to pin this file to.  It restates the canonical speculative-sampling rule the
 * reference cites (docs/papers/FINAL_PAPER.md:22,285):
 *     accept x_i  iff  u_i <= min(1, p_i(x_i)/q_i(x_i));
 *     on the first reject resample from normalise(max(0, p - q)); bonus token from p_{k+1};
 * and the reference's only vocab-wide arithmetic, softmax + log(probs[token])
 * (src/training/generate_training_data.py:128-134), for the per-row features.
 * Cross-checked in tests/ against an independent float64 numpy softmax / p-over-q test.
 *
 * Because fp32 sums are order dependent, the ARITHMETIC CONTRACT below is part of the
 * definition (SURVEY.md section 7 hard part 2).  The CUDA kernel implements the same
 * contract; this file states it with plain loops:
 *
 *   c1   = (T > 0) ? (1.0f / T) * LOG2E : LOG2E                       (fp32, rn)
 *   a_v  = z_v * c1                       m1 = max a_v, m2 = second largest (multiset)
 *   t_v  = fmaf(z_v, c1, -m1)             e_v = exp2p(t_v)   (degree-5 polynomial below)
 *   Z    = csum(e_v)                      S  = csum(e_v * t_v)
 *   csum / cscan: element v belongs to lane ((v / 4) mod 4096); a lane adds its
 *     elements in increasing v; the 4096 lane values are then scanned: Hillis-Steele
 *     inside each group of 32 lanes, a sequential chain over the 16 groups of a
 *     512-lane block, a sequential chain over the 8 blocks.  csum = the scan's total.
 *   accept:  (u * (double)eq_x) * (double)Zp  <=  (double)ep_x * (double)Zq    (binary64)
 *   residual weight r_v = max(0, fmaf(ep_v, Zq, -(eq_v * Zp)))   (bonus row: r_v = ep_v)
 *   resample: tau = ur * R (R = cscan total of lane sums of r); pick the first lane whose
 *     inclusive scan value exceeds tau, then the first element inside that lane whose
 *     running sum (started from the previous lane's scan value) exceeds tau.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LANES 4096
#define GROUP 32
#define BLOCK_LANES 512
#define NFEAT 6

static const float LOG2E_F = 0x1.715476p+0f;
static const float LN2_F = 0x1.62e43p-1f;

static float exp2p(float t)
{
    /* 2^t for t <= ~0: clamp, round-to-nearest split, degree-5 minimax polynomial on
     * [-0.5, 0.5] (Horner, fused multiply-add), exponent insertion by integer add. */
    float tc = fmaxf(t, -125.0f);
    float r = tc + 12582912.0f;
    float nf = r - 12582912.0f;
    float f = tc - nf;
    float p = 0x1.5c08e6p-10f;
    uint32_t rb, pb;
    p = fmaf(p, f, 0x1.3d0c52p-7f);
    p = fmaf(p, f, 0x1.c6b6e4p-5f);
    p = fmaf(p, f, 0x1.ebf918p-3f);
    p = fmaf(p, f, 0x1.62e428p-1f);
    p = fmaf(p, f, 0x1.000002p+0f);
    memcpy(&rb, &r, 4);
    memcpy(&pb, &p, 4);
    pb += rb << 23;
    memcpy(&p, &pb, 4);
    return p;
}

static int lane_of(int v) { return (v >> 2) & (LANES - 1); }

/* canonical scan over the 4096 lane values: P inclusive, X = previous lane's P. */
static float cscan(const float *val, float *P, float *X)
{
    static float hs[LANES];
    float tmp[GROUP];
    float coff = 0.0f;
    int c, w, i, d;
    for (w = 0; w < LANES / GROUP; ++w) {
        float *g = hs + w * GROUP;
        memcpy(g, val + w * GROUP, sizeof(float) * GROUP);
        for (d = 1; d < GROUP; d <<= 1) {
            memcpy(tmp, g, sizeof(tmp));
            for (i = d; i < GROUP; ++i)
                g[i] = tmp[i] + tmp[i - d];
        }
    }
    for (c = 0; c < LANES / BLOCK_LANES; ++c) {
        float off = 0.0f;
        for (w = 0; w < BLOCK_LANES / GROUP; ++w) {
            const float *g = hs + c * BLOCK_LANES + w * GROUP;
            for (i = 0; i < GROUP; ++i) {
                float q = off + g[i];
                P[c * BLOCK_LANES + w * GROUP + i] = coff + q;
            }
            off = off + g[GROUP - 1];
        }
        coff = coff + off;
    }
    if (X) {
        X[0] = 0.0f;
        for (i = 1; i < LANES; ++i)
            X[i] = P[i - 1];
    }
    return coff;
}

typedef struct {
    float m1, m2, Z, S;
    int argmax;
} row_stats_t;

/* passes 1+2 for one row: e[] and t[] out. */
static void row_softmax(const float *z, int V, float c1, int want_top2, float *e, float *t,
                        row_stats_t *st)
{
    static float laneZ[LANES], laneS[LANES], P[LANES];
    float m1 = -INFINITY, m2 = -INFINITY;
    int v, amax = 0;
    for (v = 0; v < V; ++v) {
        float a = z[v] * c1;
        if (a > m1) {
            m2 = m1;
            m1 = a;
            amax = v;
        } else if (a > m2) {
            m2 = a;
        }
    }
    memset(laneZ, 0, sizeof(laneZ));
    memset(laneS, 0, sizeof(laneS));
    for (v = 0; v < V; ++v) {
        int l = lane_of(v);
        t[v] = fmaf(z[v], c1, -m1);
        e[v] = exp2p(t[v]);
        laneZ[l] = laneZ[l] + e[v];
        if (want_top2)
            laneS[l] = laneS[l] + e[v] * t[v];
    }
    st->m1 = m1;
    st->m2 = m2;
    st->argmax = amax;
    st->Z = cscan(laneZ, P, NULL);
    st->S = want_top2 ? cscan(laneS, P, NULL) : 0.0f;
}

/* inverse-CDF pick over residual weights r[] in canonical lane order; -1 when R == 0 */
static int pick(const float *r, int V, float ur)
{
    static float laneR[LANES], P[LANES], X[LANES];
    float R, tau, c;
    int v, l, sel = -1, last_pos = -1;
    memset(laneR, 0, sizeof(laneR));
    for (v = 0; v < V; ++v)
        laneR[lane_of(v)] = laneR[lane_of(v)] + r[v];
    R = cscan(laneR, P, X);
    if (!(R > 0.0f))
        return -1;
    tau = ur * R;
    for (l = 0; l < LANES; ++l)
        if (P[l] > tau && X[l] <= tau)
            break;
    if (l == LANES)
        return -1; /* unreachable: P is monotone and P[last] = R > tau */
    c = X[l];
    for (v = 0; v < V; ++v) {
        if (lane_of(v) != l)
            continue;
        if (r[v] > 0.0f)
            last_pos = v;
        c = c + r[v];
        if (c > tau) {
            sel = v;
            break;
        }
    }
    return sel >= 0 ? sel : last_pos;
}

static float clamp_u(double u)
{
    float ur = (float)u;
    if (!(ur >= 0.0f))
        ur = 0.0f;
    if (ur >= 1.0f)
        ur = 0x1.fffffep-1f;
    return ur;
}

/*
 * target_logits fp32 [B, k+1, V]; draft_logits fp32 [B, k, V] (NULL allowed when k == 0
 * or temperature <= 0); draft_tokens int32 [B, k]; u_accept fp64 [B, k]; u_resid fp64 [B].
 * Outputs: accept_mask u8 [B, k]; accepted_len i32 [B]; out_tokens i32 [B, k+1] (-1 pad);
 * out_logprobs fp32 [B, k+1] (0 pad); features fp32 [B, k+1, 6] =
 * {lse, p_max, margin p1-p2, entropy, ln p(draft token), ln p(resample candidate)}.
 * Returns 0, or -1 on bad arguments.
 */
int oracle_reject_sample(const float *target_logits, const float *draft_logits,
                         const int32_t *draft_tokens, const double *u_accept,
                         const double *u_resid, int B, int k, int V, float temperature,
                         uint8_t *accept_mask, int32_t *accepted_len, int32_t *out_tokens,
                         float *out_logprobs, float *features)
{
    const int greedy = !(temperature > 0.0f);
    const float c1 = greedy ? LOG2E_F : (1.0f / temperature) * LOG2E_F;
    float *ep, *tp, *eq, *tq, *r;
    int b, i, v;
    if (B < 0 || k < 0 || V < 4 || (V & 3))
        return -1;
    ep = malloc(sizeof(float) * V);
    tp = malloc(sizeof(float) * V);
    eq = malloc(sizeof(float) * V);
    tq = malloc(sizeof(float) * V);
    r = malloc(sizeof(float) * V);
    for (b = 0; b < B; ++b) {
        int n = 0, rejected = 0;
        int cand[65];
        float lp_x[65], lp_y[65];
        if (k > 64)
            return -1;
        for (i = 0; i <= k; ++i) {
            const float *zt = target_logits + ((size_t)b * (k + 1) + i) * V;
            const int has_draft = i < k;
            const int x = has_draft ? draft_tokens[b * k + i] : -1;
            const int x_ok = has_draft && x >= 0 && x < V;
            row_stats_t sp, sq;
            float *f = features + ((size_t)b * (k + 1) + i) * NFEAT;
            float logZ;
            int acc = 0, y;
            row_softmax(zt, V, c1, 1, ep, tp, &sp);
            logZ = logf(sp.Z);
            if (greedy) {
                acc = x_ok && x == sp.argmax;
                y = sp.argmax;
            } else {
                if (has_draft) {
                    const float *zq = draft_logits + ((size_t)b * k + i) * V;
                    row_softmax(zq, V, c1, 0, eq, tq, &sq);
                    if (x_ok) {
                        double lhs = (u_accept[b * k + i] * (double)eq[x]) * (double)sp.Z;
                        double rhs = (double)ep[x] * (double)sq.Z;
                        acc = lhs <= rhs;
                    }
                    for (v = 0; v < V; ++v) {
                        float wq = eq[v] * sp.Z;
                        r[v] = fmaxf(fmaf(ep[v], sq.Z, -wq), 0.0f);
                    }
                } else {
                    for (v = 0; v < V; ++v)
                        r[v] = ep[v];
                }
                y = pick(r, V, clamp_u(u_resid[b]));
                if (y < 0) /* R == 0: p == q on this row; the draft token itself is a valid draw */
                    y = x_ok ? x : 0;
            }
            f[0] = (sp.m1 + log2f(sp.Z)) * LN2_F;
            f[1] = 1.0f / sp.Z;
            f[2] = f[1] - exp2p(sp.m2 - sp.m1) / sp.Z;
            f[3] = (log2f(sp.Z) - sp.S / sp.Z) * LN2_F;
            f[4] = x_ok ? logf(ep[x]) - logZ : -INFINITY;
            f[5] = logf(ep[y]) - logZ;
            cand[i] = y;
            lp_x[i] = f[4];
            lp_y[i] = f[5];
            if (has_draft) {
                accept_mask[b * k + i] = (uint8_t)(acc && !rejected);
                if (acc && !rejected)
                    n++;
                else
                    rejected = 1;
            }
        }
        accepted_len[b] = n;
        for (i = 0; i <= k; ++i) {
            out_tokens[b * (k + 1) + i] = i < n ? draft_tokens[b * k + i] : (i == n ? cand[n] : -1);
            out_logprobs[b * (k + 1) + i] = i < n ? lp_x[i] : (i == n ? lp_y[n] : 0.0f);
        }
    }
    free(ep);
    free(tp);
    free(eq);
    free(tq);
    free(r);
    return 0;
}

/* exposed for unit tests of the contract pieces */
float oracle_exp2p(float t) { return exp2p(t); }
float oracle_cscan_total(const float *lane_vals) { static float P[LANES]; return cscan(lane_vals, P, NULL); }
