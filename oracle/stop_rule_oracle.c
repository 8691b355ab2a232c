/*
 * ORACLE (test infrastructure only - never imported by the product path).
 *
 * CPU restatement, in plain C, of the reference's optimal-stopping policy:
 *   optimal_stopping_rule      /root/reference/src/algorithms/dp_solver.py:12-71
 *   compute_expected_cost      /root/reference/src/algorithms/dp_solver.py:74-103
 *   bayesian_adjustment        /root/reference/src/algorithms/dp_solver.py:106-130
 *   derive_optimal_policy      /root/reference/src/theory/optimal_stopping.py:45-91
 *
 * CPython floats are IEEE binary64 and every operation below is a single
 * correctly-rounded binary64 add / sub / mul / div, so replaying the same
 * operations in the same order in C (compiled with -ffp-contract=off) is
 * bit-exact with the reference.  Pinned by tests/golden/stop_rule_golden.json,
 * which oracle/gen_golden.py produced by importing the reference's own
 * dp_solver.py / optimal_stopping.py by file path.
 */
#include <stddef.h>
#include <stdint.h>

#define ORACLE_MAX_STAGES 64

/* dp_solver.py:106-130.  n_obs is an int in the reference; int*float promotes to
 * binary64, so a double argument is equivalent for every int below 2^53. */
double oracle_bayesian_adjustment(double p_hat, double n_obs, double alpha, double beta)
{
    double posterior_alpha = n_obs * p_hat + alpha;          /* :122 */
    double posterior_beta = n_obs * (1 - p_hat) + beta;      /* :123 */
    return posterior_alpha / (posterior_alpha + posterior_beta); /* :126 */
}

/* dp_solver.py:12-71.  Returns k_star, or -1 when L is out of range.
 * J must have room for L+1 doubles. */
int oracle_optimal_stopping_rule(const double *p_in, const double *C, int L, double lam,
                                 int risk_adjustment, double alpha, double beta, double *J)
{
    double p[ORACLE_MAX_STAGES];
    double p_bar[ORACLE_MAX_STAGES + 1];
    int stop_decision[ORACLE_MAX_STAGES];
    int i;
    if (L < 1 || L > ORACLE_MAX_STAGES)
        return -1;
    for (i = 0; i < L; ++i)                                  /* :40-41 */
        p[i] = risk_adjustment ? oracle_bayesian_adjustment(p_in[i], 100.0, alpha, beta) : p_in[i];
    p_bar[0] = 1.0;                                          /* :44-46 */
    for (i = 0; i < L; ++i)
        p_bar[i + 1] = p_bar[i] * p[i];
    for (i = 0; i <= L; ++i)                                 /* :49 */
        J[i] = 0.0;
    for (i = L - 1; i >= 0; --i) {                           /* :53-66 */
        double cost_if_stop = C[i] + lam * (1 - p_bar[i + 1]);
        double cost_if_continue = C[i] + J[i + 1];
        if (cost_if_stop <= cost_if_continue) {
            stop_decision[i] = 1;
            J[i] = cost_if_stop;
        } else {
            stop_decision[i] = 0;
            J[i] = cost_if_continue;
        }
    }
    for (i = 0; i < L; ++i)                                  /* :69 */
        if (stop_decision[i])
            return i;
    return L - 1;
}

/* dp_solver.py:74-103.  Python's sum() starts from int 0 and adds left to right. */
double oracle_compute_expected_cost(const double *p, const double *C, double lam, int stopping_stage)
{
    double p_bar = 1.0, computation_cost = 0.0;
    int i;
    for (i = 0; i <= stopping_stage; ++i)
        p_bar *= p[i];
    for (i = 0; i <= stopping_stage; ++i)
        computation_cost = computation_cost + C[i];
    return computation_cost + lam * (1 - p_bar);
}

/* optimal_stopping.py:45-91 (threshold policy).  quality_bounds q / cost_ratios c of
 * length n <= ORACLE_MAX_STAGES; thresholds[n] out.  Operation order follows :62-80
 * and :84-91; V is the reference's np.zeros(n + 1) value table. */
void oracle_derive_optimal_policy(const double *q, const double *c, int n, double lam,
                                  double *thresholds)
{
    double V[ORACLE_MAX_STAGES + 1];
    int s;
    for (s = 0; s <= n; ++s)
        V[s] = 0.0;
    for (s = n - 1; s >= 0; --s) {
        double r_stop = q[s] - lam * c[s];                               /* :62 */
        if (s < n - 1) {
            double p_improve = 0.6 * (1 - q[s]);                         /* :91 */
            double r_continue = p_improve * V[s + 1] + (1 - p_improve) * r_stop; /* :68 */
            V[s] = r_stop >= r_continue ? r_stop : r_continue;           /* :72 max() */
            thresholds[s] = (V[s + 1] + lam * c[s]) / (1 + lam * (c[s + 1] - c[s])); /* :76 */
        } else {
            V[s] = r_stop;                                               /* r_continue = -inf */
            thresholds[s] = 0.0;                                         /* :78 */
        }
    }
}
