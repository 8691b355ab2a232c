/*
 * ORACLE (test infrastructure only - never imported by the product path).
 *
 * PARITY UNPINNED against the reference: it contains no token-level rejection sampler (SURVEY.md
 * section 0 fact 2, section 8c), so there is no reference implementation, test or golden vector
 * to pin this file to.  It restates the canonical speculative-sampling rule the reference cites
 * (docs/papers/FINAL_PAPER.md:22,285):
 *     accept x_i  iff  u_i <= min(1, p_i(x_i)/q_i(x_i));
 *     on the first reject resample from normalise(max(0, p - q)); bonus token from p_{k+1};
 * and the reference's only vocab-wide arithmetic, softmax + log(probs[token])
 * (src/training/generate_training_data.py:128-134), for the per-row features.
 * Independent pins: tests/test_oracle_sampler.py (float64 numpy softmax / p-over-q, chi-square of the
 * resampler) and tests/test_sampler_vllm_gpu.py (the accept rule of the installed vLLM kernels).
 *
 * Because fp32 sums are order dependent, the ARITHMETIC CONTRACT below is part of the definition
 * (SURVEY.md section 7 hard part 2).  The CUDA kernels (csrc/sampler.cu) implement the same contract;
 * this file states it with plain loops.  Contract v2 ("chunked streaming": every logit is read from
 * HBM once by an independent 4096-element chunk, no vocabulary-wide exchange before the exponentials):
 *
 *   c1 = (T > 0) ? (1.0f / T) * LOG2E : LOG2E                              (fp32, rn)
 *   A row is cut into chunks of 4096 consecutive elements.  Inside chunk c, element with local index u
 *   belongs to lane ((u / 4) mod 256) and a lane visits its (up to 16) elements in increasing u.
 *     a_v  = z_v * c1;   m_c = max a_v over the chunk;  (m_c, m2_c) = two largest of the multiset
 *     t_v  = fmaf(z_v, c1, -m_c);     e_v = exp2p(t_v)      (degree-5 polynomial below)
 *     lane sums of e_v and of e_v * t_v: a lane keeps two running sums, one over its elements with even local
 *     index and one over the odd ones (each left to right from 0), and adds them (even + odd) at the end;
 *     then wsum: Hillis-Steele inclusive scan inside each group of 32 lanes, a sequential chain over the 8
 *     groups (its last value is the total).
 *   Row merge, sequential over the chunks:  M = max m_c;  s_c = exp2p(m_c - M);
 *     Z = sum_c Z_c * s_c;   S = sum_c fmaf(Z_c, m_c - M, S_c) * s_c;   second maximum by the usual merge.
 *   A probability is never formed; P_v = e_v * s_c is "p_v * Z".
 *   accept:  (u * (double)Q_x) * (double)Zp  <=  (double)P_x * (double)Zq                 (binary64)
 *   Only the row that emits the sequence's new token (the first rejected position, or the bonus row)
 *   draws from the residual  r_v = max(0, fmaf(P_v, Zq, -(Q_v * Zp)))  (bonus row: r_v = P_v):
 *     chunk totals R_c by lane sums (ONE running sum per lane here, left to right) + wsum, R = sequential sum, tau = ur * R (0 if that is not < R);
 *     the first chunk whose running total exceeds tau, inside it the first lane whose inclusive scan value
 *     X + (off_group + hs_lane) exceeds tau (X = running total before the chunk), inside the lane the
 *     first element whose running sum (started from the previous lane's scan value) exceeds tau.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CHUNK 4096
#define LANES 256
#define GROUP 32
#define NFEAT 6
#define MAXCH 64

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

static int lane_of(int u) { return (u >> 2) & (LANES - 1); }

/* canonical scan of the 256 lane values of a chunk: hs = Hillis-Steele inclusive inside each group of 32,
 * off[g] = sequential sum of the totals of the groups before g; returns the chunk total. */
static float wscan(const float *val, float *hs, float *off)
{
    float tmp[GROUP];
    float o = 0.0f;
    int w, i, d;
    for (w = 0; w < LANES / GROUP; ++w) {
        float *g = hs + w * GROUP;
        memcpy(g, val + w * GROUP, sizeof(float) * GROUP);
        for (d = 1; d < GROUP; d <<= 1) {
            memcpy(tmp, g, sizeof(tmp));
            for (i = d; i < GROUP; ++i)
                g[i] = tmp[i] + tmp[i - d];
        }
        off[w] = o;
        o = o + g[GROUP - 1];
    }
    return o;
}

static float wsum(const float *val)
{
    float hs[LANES], off[LANES / GROUP];
    return wscan(val, hs, off);
}

typedef struct {
    int nc;
    float m[MAXCH], m2[MAXCH], Z[MAXCH], S[MAXCH], s[MAXCH]; /* per chunk; s = exp2p(m_c - M) */
    float M, M2, Zt, St;
    int argmax;
} row_t;

/* chunk statistics + row merge of one row of logits; e[] = chunk-relative exponentials */
static void row_stats(const float *z, int V, float c1, float *e, row_t *R)
{
    int c, v, i;
    R->nc = (V + CHUNK - 1) / CHUNK;
    for (c = 0; c < R->nc; ++c) {
        const int v0 = c * CHUNK, n = (V - v0) < CHUNK ? (V - v0) : CHUNK;
        float laneZ[2][LANES], laneS[2][LANES], lz[LANES], ls[LANES];
        float m1 = -INFINITY, m2 = -INFINITY;
        for (i = 0; i < n; ++i) {
            float a = z[v0 + i] * c1;
            if (a > m1) {
                m2 = m1;
                m1 = a;
            } else if (a > m2) {
                m2 = a;
            }
        }
        memset(laneZ, 0, sizeof(laneZ));
        memset(laneS, 0, sizeof(laneS));
        for (i = 0; i < n; ++i) {
            const int l = lane_of(i);
            const float t = fmaf(z[v0 + i], c1, -m1);
            const float ex = exp2p(t);
            e[v0 + i] = ex;
            laneZ[i & 1][l] = laneZ[i & 1][l] + ex;
            laneS[i & 1][l] = laneS[i & 1][l] + ex * t;
        }
        for (i = 0; i < LANES; ++i) {
            lz[i] = laneZ[0][i] + laneZ[1][i];
            ls[i] = laneS[0][i] + laneS[1][i];
        }
        R->m[c] = m1;
        R->m2[c] = m2;
        R->Z[c] = wsum(lz);
        R->S[c] = wsum(ls);
    }
    R->M = -INFINITY;
    R->M2 = -INFINITY;
    for (c = 0; c < R->nc; ++c) {
        R->M2 = fmaxf(fminf(R->M, R->m[c]), fmaxf(R->M2, R->m2[c]));
        R->M = fmaxf(R->M, R->m[c]);
    }
    R->Zt = 0.0f;
    R->St = 0.0f;
    for (c = 0; c < R->nc; ++c) {
        const float d = R->m[c] - R->M;
        R->s[c] = exp2p(d);
        R->Zt = R->Zt + R->Z[c] * R->s[c];
        R->St = R->St + fmaf(R->Z[c], d, R->S[c]) * R->s[c];
    }
    /* arg-max: lowest index attaining the row maximum */
    R->argmax = 0;
    for (v = 0; v < V; ++v)
        if (z[v] * c1 == R->M) {
            R->argmax = v;
            break;
        }
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

/* inverse-CDF draw over the residual weights r[] (already in canonical form); -1 when their total is 0 */
static int pick(const float *r, int V, float ur)
{
    const int nc = (V + CHUNK - 1) / CHUNK;
    float Rc[MAXCH], lane[LANES], hs[LANES], off[LANES / GROUP];
    float R = 0.0f, tau, pre = 0.0f, X = 0.0f, cc;
    int c, i, l, cstar = -1, sel = -1, last_pos = -1, first = -1, lstar = -1;
    for (c = 0; c < nc; ++c) {
        const int v0 = c * CHUNK, n = (V - v0) < CHUNK ? (V - v0) : CHUNK;
        memset(lane, 0, sizeof(lane));
        for (i = 0; i < n; ++i)
            lane[lane_of(i)] = lane[lane_of(i)] + r[v0 + i];
        Rc[c] = wsum(lane);
        R = R + Rc[c];
    }
    if (!(R > 0.0f))
        return -1;
    tau = ur * R;
    if (!(tau < R))
        tau = 0.0f;
    for (c = 0; c < nc; ++c) {
        const float nx = pre + Rc[c];
        if (nx > tau) {
            cstar = c;
            X = pre;
            break;
        }
        pre = nx;
    }
    if (cstar < 0)
        return -1; /* unreachable: the running total ends at R > tau */
    {
        const int v0 = cstar * CHUNK, n = (V - v0) < CHUNK ? (V - v0) : CHUNK;
        float Pprev = 0.0f;
        memset(lane, 0, sizeof(lane));
        for (i = 0; i < n; ++i)
            lane[lane_of(i)] = lane[lane_of(i)] + r[v0 + i];
        wscan(lane, hs, off);
        for (l = 0; l < LANES; ++l) {
            const float P = X + (off[l / GROUP] + hs[l]);
            /* scan value of the previous lane; for the first lane of a group: X + off (same expression as
             * the last lane of the previous group) */
            const float Xl = (l % GROUP == 0) ? X + off[l / GROUP] : Pprev;
            Pprev = P;
            if (P > tau) {
                lstar = l;
                cc = Xl;
                break;
            }
        }
        if (lstar < 0)
            return -1; /* unreachable: the last lane's value equals the running total > tau */
        for (i = 0; i < n; ++i) {
            if (lane_of(i) != lstar)
                continue;
            if (first < 0)
                first = v0 + i;
            if (r[v0 + i] > 0.0f)
                last_pos = v0 + i;
            cc = cc + r[v0 + i];
            if (cc > tau) {
                sel = v0 + i;
                break;
            }
        }
    }
    return sel >= 0 ? sel : (last_pos >= 0 ? last_pos : first);
}

/*
 * target_logits fp32 [B, k+1, V]; draft_logits fp32 [B, k, V] (NULL allowed when k == 0
 * or temperature <= 0); draft_tokens int32 [B, k]; u_accept fp64 [B, k]; u_resid fp64 [B].
 * Outputs: accept_mask u8 [B, k]; accepted_len i32 [B]; out_tokens i32 [B, k+1] (-1 pad);
 * out_logprobs fp32 [B, k+1] (0 pad); features fp32 [B, k+1, 6] =
 * {lse, p_max, margin p1-p2, entropy, ln p(draft token), ln p(token emitted at this position) (0 past it)}.
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
    float *ep, *eq, *r;
    int b, i, v;
    if (B < 0 || k < 0 || k > 64 || V < 4 || (V & 3) || V > CHUNK * MAXCH)
        return -1;
    ep = malloc(sizeof(float) * V);
    eq = malloc(sizeof(float) * V);
    r = malloc(sizeof(float) * V);
    for (b = 0; b < B; ++b) {
        int n = 0, rejected = 0, y;
        int amax[65];
        float lp_x[65], lp_amax[65], lp_y;
        for (i = 0; i <= k; ++i) {
            const float *zt = target_logits + ((size_t)b * (k + 1) + i) * V;
            const int has_draft = i < k;
            const int x = has_draft ? draft_tokens[b * k + i] : -1;
            const int x_ok = has_draft && x >= 0 && x < V;
            row_t sp, sq;
            float *f = features + ((size_t)b * (k + 1) + i) * NFEAT;
            float logZ, log2Z, Px = 0.0f;
            int acc = 0;
            row_stats(zt, V, c1, ep, &sp);
            logZ = logf(sp.Zt);
            log2Z = log2f(sp.Zt);
            if (x_ok)
                Px = ep[x] * sp.s[x / CHUNK];
            if (greedy) {
                acc = x_ok && x == sp.argmax;
            } else if (has_draft && x_ok) {
                const float *zq = draft_logits + ((size_t)b * k + i) * V;
                float Qx;
                double lhs, rhs;
                row_stats(zq, V, c1, eq, &sq);
                Qx = eq[x] * sq.s[x / CHUNK];
                lhs = (u_accept[b * k + i] * (double)Qx) * (double)sp.Zt;
                rhs = (double)Px * (double)sq.Zt;
                acc = lhs <= rhs;
            }
            f[0] = (sp.M + log2Z) * LN2_F;
            f[1] = 1.0f / sp.Zt;
            f[2] = f[1] - exp2p(sp.M2 - sp.M) / sp.Zt;
            f[3] = (log2Z - sp.St / sp.Zt) * LN2_F;
            f[4] = x_ok ? logf(Px) - logZ : -INFINITY;
            f[5] = 0.0f;
            amax[i] = sp.argmax;
            lp_x[i] = f[4];
            lp_amax[i] = logf(ep[sp.argmax] * sp.s[sp.argmax / CHUNK]) - logZ;
            if (has_draft) {
                accept_mask[b * k + i] = (uint8_t)(acc && !rejected);
                if (acc && !rejected)
                    n++;
                else
                    rejected = 1;
            }
        }
        /* the emitting row n: arg-max (greedy) or a draw from the residual / the bonus distribution */
        {
            const float *zt = target_logits + ((size_t)b * (k + 1) + n) * V;
            const int has_draft = n < k;
            const int x = has_draft ? draft_tokens[b * k + n] : -1;
            const int x_ok = has_draft && x >= 0 && x < V;
            row_t sp, sq;
            if (greedy) {
                y = amax[n];
                lp_y = lp_amax[n];
            } else {
                row_stats(zt, V, c1, ep, &sp);
                if (has_draft) {
                    const float *zq = draft_logits + ((size_t)b * k + n) * V;
                    row_stats(zq, V, c1, eq, &sq);
                    for (v = 0; v < V; ++v) {
                        const float P = ep[v] * sp.s[v / CHUNK];
                        const float Q = eq[v] * sq.s[v / CHUNK];
                        const float wq = Q * sp.Zt;
                        r[v] = fmaxf(fmaf(P, sq.Zt, -wq), 0.0f);
                    }
                } else {
                    for (v = 0; v < V; ++v)
                        r[v] = ep[v] * sp.s[v / CHUNK];
                }
                y = pick(r, V, clamp_u(u_resid[b]));
                if (y < 0) /* residual mass is zero: p == q on this row; the draft token itself is a valid draw */
                    y = x_ok ? x : 0;
                lp_y = logf(ep[y] * sp.s[y / CHUNK]) - logf(sp.Zt);
            }
        }
        accepted_len[b] = n;
        for (i = 0; i <= k; ++i) {
            out_tokens[b * (k + 1) + i] = i < n ? draft_tokens[b * k + i] : (i == n ? y : -1);
            out_logprobs[b * (k + 1) + i] = i < n ? lp_x[i] : (i == n ? lp_y : 0.0f);
            features[((size_t)b * (k + 1) + i) * NFEAT + 5] = out_logprobs[b * (k + 1) + i];
        }
    }
    free(ep);
    free(eq);
    free(r);
    return 0;
}

/* exposed for unit tests of the contract pieces */
float oracle_exp2p(float t) { return exp2p(t); }
float oracle_wsum(const float *lane_vals) { return wsum(lane_vals); }
