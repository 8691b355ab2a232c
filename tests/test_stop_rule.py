"""Policy seam: asd_b200.algorithms.dp_solver (C-ABI host function and CUDA batch kernel)
against the reference-generated goldens and the C oracle.  Bar: bit-exact."""
import numpy as np
import pytest

import oracle
from asd_b200.algorithms import dp_solver
from conftest import fh


def test_host_rule_matches_reference_goldens(stop_rule_golden):
    for g in stop_rule_golden["stopping_rule"]:
        k, J = dp_solver.optimal_stopping_rule([fh(x) for x in g["p"]], [fh(x) for x in g["C"]], fh(g["lam"]),
                                               g["risk_adjustment"], fh(g["alpha"]), fh(g["beta"]))
        assert k == g["k_star"] and [x.hex() for x in J] == g["J"]


def test_host_bayes_and_cost_match_goldens(stop_rule_golden):
    for g in stop_rule_golden["bayesian"]:
        out = dp_solver.bayesian_adjustment(fh(g["p_hat"]), g["n_obs"], fh(g["alpha"]), fh(g["beta"]))
        assert out.hex() == g["out"]
    for g in stop_rule_golden["expected_cost"]:
        out = dp_solver.compute_expected_cost([fh(x) for x in g["p"]], [fh(x) for x in g["C"]], fh(g["lam"]),
                                              g["stage"])
        assert float(out).hex() == g["out"]


def test_errors_and_edge_cases():
    with pytest.raises(ValueError):
        dp_solver.optimal_stopping_rule([0.5], [1.0, 2.0], 1.0)
    assert dp_solver.optimal_stopping_rule([1.0], [1.0], 5.0) == (0, [1.0, 0.0])
    assert dp_solver.optimal_stopping_rule([], [], 1.0) == (-1, [0.0])


def test_table_and_solver_helpers():
    t = dp_solver.OptimalStoppingTable([0.5, 1.0, 12.0], num_stages=2)
    t.precompute([1.0, 4.5], [[0.6, 1.0], [0.2, 1.0]])
    assert t.lookup([0.6, 1.0], 11.0) == 1 and t.lookup([0.6, 1.0], 1.1) == 0
    assert t.lookup([0.33, 0.9], 1.0) == dp_solver.optimal_stopping_rule([0.33, 0.9], [1.0, 1.6], 1.0)[0]
    s = dp_solver.DynamicProgrammingSolver(3, [1.0, 4.5, 10.0])
    assert s.should_stop(2, 0.0, 0.1, 1.0) is True
    assert s.should_stop(0, 0.0, 0.95, 1.0) is True
    assert s.should_stop(0, 0.0, 0.05, 50.0) is False


def test_rows_host_matches_scalar_rule():
    rng = np.random.default_rng(3)
    n, L = 257, 4
    p, C, lam = rng.random((n, L)), rng.uniform(0.5, 12.0, (n, L)), rng.choice([0.1, 0.5, 1, 2, 5, 10], n)
    for risk in (False, True):
        k, J = dp_solver.stop_rule_rows(p, C, lam, risk, 2.0, 0.5)
        for r in range(n):
            k1, J1 = dp_solver.optimal_stopping_rule(list(p[r]), list(C[r]), float(lam[r]), risk, 2.0, 0.5)
            assert k1 == k[r] and [x.hex() for x in J1] == [float(x).hex() for x in J[r]]


def test_table_and_adaptive_match_live_reference():
    """OptimalStoppingTable (one batched call) and AdaptiveStopping against the reference's own classes"""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference not present (GPU box)")
    ref = ref_loader.dp_solver()
    rng = np.random.default_rng(11)
    lams = [0.1, 0.5, 1.0, 2.0, 5.0, 10.0]
    grid = [[round(a, 2), round(b, 2), round(c, 2), 1.0] for a in np.linspace(0.05, 0.95, 7)
            for b in np.linspace(0.1, 0.9, 5) for c in (0.3, 0.725, 0.9)]
    grid += [[0.123, 0.456, 0.789, 1.0], [0.125, 0.455, 0.785, 1.0]]       # colliding rounded keys
    costs = [1.0, 2.0, 4.5, 10.0]
    mine, theirs = dp_solver.OptimalStoppingTable(lams), ref.OptimalStoppingTable(lams)
    mine.precompute(costs, grid)
    theirs.precompute(costs, grid)
    assert mine.table == theirs.table
    for _ in range(300):
        pr = [round(float(x), int(rng.integers(2, 5))) for x in rng.random(4)]
        lam = float(rng.uniform(0.05, 12.0))
        for fb in (True, False):
            assert mine.lookup(pr, lam, fb) == theirs.lookup(pr, lam, fb)
    a, b = dp_solver.AdaptiveStopping(1.5, 0.05), ref.AdaptiveStopping(1.5, 0.05)
    for _ in range(200):
        st, q, lat = int(rng.integers(0, 4)), float(rng.random()), float(rng.uniform(50, 3000))
        a.update_statistics(st, q, lat)
        b.update_statistics(st, q, lat)
        for s_ in range(4):
            assert a.get_confidence_bounds(s_) == b.get_confidence_bounds(s_)
            assert a.should_explore(s_) == b.should_explore(s_)
    assert np.array_equal(a.stage_counts, b.stage_counts) and np.array_equal(a.stage_rewards, b.stage_rewards)
    assert a.total_steps == b.total_steps


@pytest.mark.gpu
def test_device_rows_bit_exact():
    import torch
    rng = np.random.default_rng(5)
    n, L = 5000, 3
    p, C, lam = rng.random((n, L)), np.tile([1.0, 4.5, 10.0], (n, 1)), rng.choice([0.1, 0.5, 1, 2, 5, 10], n).astype(np.float64)
    k_h, J_h = dp_solver.stop_rule_rows(p, C, lam)
    k_d, J_d = dp_solver.stop_rule_rows(torch.from_numpy(p).cuda(), torch.from_numpy(C).cuda(), torch.from_numpy(lam).cuda())
    assert np.array_equal(k_h, k_d.cpu().numpy())
    assert np.array_equal(J_h.view(np.uint64), J_d.cpu().numpy().view(np.uint64))


@pytest.mark.gpu
def test_device_batch_bit_exact(stop_rule_golden):
    import torch
    rng = np.random.default_rng(7)
    for L, C in [(3, [1.0, 4.5, 10.0]), (4, [1.0, 2.0, 4.5, 10.0]), (1, [1.0]), (8, list(np.linspace(1, 20, 8)))]:
        for lam in (0.1, 1.0, 10.0):
            for risk in (False, True):
                n = 4099
                p = rng.random((n, L))
                p[::3, -1] = 1.0
                Cm = np.tile(np.asarray(C, np.float64), (n, 1)) * rng.uniform(0.5, 2.0, (n, 1))
                k_d, J_d = dp_solver.stop_rule_batch(torch.from_numpy(p).cuda(), torch.from_numpy(Cm).cuda(), lam,
                                                     risk, 2.0, 0.5)
                k_d, J_d = k_d.cpu().numpy(), J_d.cpu().numpy()
                for r in range(0, n, 41):
                    k_o, J_o = oracle.optimal_stopping_rule(p[r], Cm[r], lam, risk, 2.0, 0.5)
                    assert k_o == k_d[r]
                    assert np.array_equal(np.asarray(J_o).view(np.uint64), J_d[r].view(np.uint64))
    # the reference goldens themselves, through the device kernel
    gs = [g for g in stop_rule_golden["stopping_rule"] if len(g["p"]) == 4 and not g["risk_adjustment"]]
    for lam_hex in sorted({g["lam"] for g in gs})[:6]:
        sel = [g for g in gs if g["lam"] == lam_hex]
        p = torch.tensor([[fh(x) for x in g["p"]] for g in sel], dtype=torch.float64).cuda()
        C = torch.tensor([[fh(x) for x in g["C"]] for g in sel], dtype=torch.float64).cuda()
        k_d, J_d = dp_solver.stop_rule_batch(p, C, fh(lam_hex))
        assert k_d.cpu().tolist() == [g["k_star"] for g in sel]
        assert [[float(x).hex() for x in row] for row in J_d.cpu().tolist()] == [g["J"] for g in sel]
