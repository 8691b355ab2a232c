"""ORACLE (test infrastructure / CPU baseline only): the stop rule exactly as the reference executes it - CPython
floats, one request at a time - restated from /root/reference/src/algorithms/dp_solver.py:12-71 (backward induction)
and :106-130 (Beta-posterior shrinkage).  Pinned by tests/golden/stop_rule_golden.json (reference-generated, hex
floats).  ``bench.py --impl reference`` times it over 10^5 (p, C, lambda) triples on one core, which is what
the reference's own per-request call costs; /root/reference itself does not exist on the
GPU box."""
from __future__ import annotations

import time


def bayesian_adjustment(p_hat, n_obs, alpha=1.0, beta=1.0):
    a = n_obs * p_hat + alpha                       # :122
    b = n_obs * (1 - p_hat) + beta                  # :123
    return a / (a + b)                              # :126


def optimal_stopping_rule(p, C, lam, risk_adjustment=False, alpha=1.0, beta=1.0):
    L = len(C)
    if len(p) != L:
        raise ValueError("p and C must have the same length")          # :34-35
    if risk_adjustment:
        p = [bayesian_adjustment(x, 100, alpha, beta) for x in p]      # :40-41
    p_bar = [1.0] * (L + 1)
    for i in range(L):
        p_bar[i + 1] = p_bar[i] * p[i]                                 # :44-46
    J = [0.0] * (L + 1)
    stop = [False] * L
    for i in reversed(range(L)):                                       # :53-66
        cost_if_stop = C[i] + lam * (1 - p_bar[i + 1])
        cost_if_continue = C[i] + J[i + 1]
        if cost_if_stop <= cost_if_continue:
            J[i], stop[i] = cost_if_stop, True
        else:
            J[i] = cost_if_continue
    k_star = next((i for i, s in enumerate(stop) if s), L - 1)         # :69
    return k_star, J


def make_triples(n=100_000, seed=7):
    """L in {3, 4}, C = [1, 4.5, 10] / [1, 2, 4.5, 10], lambda in {0.1 .. 10}"""
    import numpy as np
    rng = np.random.default_rng(seed)
    lams = rng.choice([0.1, 0.5, 1.0, 2.0, 5.0, 10.0], n)
    four = rng.random(n) < 0.5
    p = rng.random((n, 4))
    return p, four, lams


def time_triples(n=100_000, seed=7):
    p, four, lams = make_triples(n, seed)
    C3, C4 = [1.0, 4.5, 10.0], [1.0, 2.0, 4.5, 10.0]
    rows = [(list(p[i, :4]) if four[i] else list(p[i, :3]), C4 if four[i] else C3, float(lams[i])) for i in range(n)]
    t0 = time.perf_counter()
    acc = 0
    for pr, C, lam in rows:
        pr = [bayesian_adjustment(x, 100, 1.0, 1.0) for x in pr[:-1]] + [1.0]     # pipeline.py:235-242
        acc += optimal_stopping_rule(pr, C, lam)[0]
    dt = time.perf_counter() - t0
    return {"decisions_per_second": n / dt, "seconds": dt, "n": n, "cores": 1, "checksum": int(acc),
            "what": "bayesian_adjustment + optimal_stopping_rule per request in CPython (dp_solver.py:12-71,106-130)"}
