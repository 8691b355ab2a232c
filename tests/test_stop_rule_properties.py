"""Size-independent properties of the stop rule (hypothesis): the library's host function agrees bit for bit
with the C oracle on arbitrary inputs, and the DP's structural invariants hold."""
import math

from hypothesis import given, settings
from hypothesis import strategies as st

import oracle
from asd_b200.algorithms import dp_solver

prob = st.floats(min_value=0.0, max_value=1.0, allow_nan=False)
cost = st.floats(min_value=0.0, max_value=100.0, allow_nan=False)
lam_s = st.floats(min_value=0.0, max_value=1000.0, allow_nan=False)


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 8).flatmap(lambda n: st.tuples(st.lists(prob, min_size=n, max_size=n),
                                                      st.lists(cost, min_size=n, max_size=n))), lam_s, st.booleans())
def test_library_equals_oracle_bitwise(pc, lam, risk):
    p, C = pc
    k0, J0 = oracle.optimal_stopping_rule(p, C, lam, risk, 2.0, 3.0)
    k1, J1 = dp_solver.optimal_stopping_rule(p, C, lam, risk, 2.0, 3.0)
    assert k0 == k1 and [x.hex() for x in J0] == [x.hex() for x in J1]


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 8).flatmap(lambda n: st.tuples(st.lists(prob, min_size=n, max_size=n),
                                                      st.lists(cost, min_size=n, max_size=n))), lam_s)
def test_dp_invariants(pc, lam):
    p, C = pc
    L = len(C)
    k, J = dp_solver.optimal_stopping_rule(p, C, lam)
    assert 0 <= k <= L - 1 and len(J) == L + 1 and J[L] == 0.0
    pbar = 1.0
    for i in range(L):
        pbar *= p[i]
        stop, cont = C[i] + lam * (1 - pbar), C[i] + J[i + 1]
        assert J[i] == min(stop, cont)                 # Bellman equation, `<=` picks stop on ties
        assert J[i] >= C[i] - 1e-12
    # the last stage never flags "stop" unless the penalty is exactly zero, so k falls back to L-1 or earlier
    first_stop = next((i for i in range(L) if C[i] + lam * (1 - math.prod(p[:i + 1])) <= C[i] + J[i + 1]), L - 1)
    assert k == first_stop


@settings(max_examples=200, deadline=None)
@given(prob, st.integers(1, 100000), st.floats(0.1, 10.0), st.floats(0.1, 10.0))
def test_bayes_between_prior_and_estimate(p_hat, n, a, b):
    out = dp_solver.bayesian_adjustment(p_hat, n, a, b)
    assert out.hex() == oracle.bayesian_adjustment(p_hat, n, a, b).hex()
    lo, hi = sorted([p_hat, a / (a + b)])
    assert lo - 1e-12 <= out <= hi + 1e-12
