"""The C oracle of the stop rule must be bit-exact with the golden vectors that
oracle/gen_golden.py produced from the reference's own dp_solver.py / optimal_stopping.py."""
import pytest

import oracle
from oracle import ref_loader
from conftest import fh


def test_stopping_rule_golden(stop_rule_golden):
    for g in stop_rule_golden["stopping_rule"]:
        k, J = oracle.optimal_stopping_rule([fh(x) for x in g["p"]], [fh(x) for x in g["C"]], fh(g["lam"]),
                                            g["risk_adjustment"], fh(g["alpha"]), fh(g["beta"]))
        assert k == g["k_star"]
        assert [x.hex() for x in J] == g["J"]


def test_documented_example(stop_rule_golden):
    # docs/guides/GETTING_STARTED.md:61-65
    k, J = oracle.optimal_stopping_rule([0.3, 0.5, 0.7, 0.9], [1.0, 1.6, 4.2, 8.8], 1.0)
    assert k == 0 and J == [1.7, 2.45, 5.095000000000001, 8.8, 0.0]
    k, J = oracle.optimal_stopping_rule([0.6, 1.0], [1.0, 4.5], 12.0)
    assert k == 1 and J == [5.5, 4.5, 0.0]


def test_bayesian_golden(stop_rule_golden):
    for g in stop_rule_golden["bayesian"]:
        out = oracle.bayesian_adjustment(fh(g["p_hat"]), g["n_obs"], fh(g["alpha"]), fh(g["beta"]))
        assert out.hex() == g["out"]


def test_expected_cost_golden(stop_rule_golden):
    for g in stop_rule_golden["expected_cost"]:
        out = oracle.compute_expected_cost([fh(x) for x in g["p"]], [fh(x) for x in g["C"]], fh(g["lam"]), g["stage"])
        assert out.hex() == g["out"]


def test_policy_golden(stop_rule_golden):
    for g in stop_rule_golden["policy"]:
        th = oracle.derive_optimal_policy([fh(x) for x in g["quality_bounds"]], [fh(x) for x in g["cost_ratios"]],
                                          fh(g["lam"]))
        assert {str(s): v.hex() for s, v in th.items()} == g["thresholds"]


def test_length_mismatch_raises():
    with pytest.raises(ValueError):
        oracle.optimal_stopping_rule([0.5], [1.0, 2.0], 1.0)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference():
    import random
    dp = ref_loader.dp_solver()
    rng = random.Random(99)
    for _ in range(2000):
        L = rng.randint(1, 6)
        p = [rng.random() for _ in range(L)]
        C = sorted(rng.uniform(0.1, 20) for _ in range(L))
        lam = rng.uniform(0, 30)
        ra = rng.random() < 0.5
        k0, J0 = dp.optimal_stopping_rule(list(p), list(C), lam, risk_adjustment=ra)
        k1, J1 = oracle.optimal_stopping_rule(p, C, lam, ra)
        assert k0 == k1 and [x.hex() for x in J0] == [x.hex() for x in J1]
