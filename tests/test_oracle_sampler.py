"""Sampler oracle cross-checked against an independent float64 numpy statement of canonical speculative sampling
(the reference has no sampler of its own; the pin to vLLM's kernels is tests/test_oracle_sampler_vllm.py)."""
import numpy as np
import pytest

import oracle


def f64_reference(tl, dl, dt, ua, T):
    """accept iff u <= p/q in float64; returns accept decisions and their margins."""
    B, k1, V = tl.shape
    k = k1 - 1
    def sm(z):
        z = z.astype(np.float64) / T
        z = z - z.max(-1, keepdims=True)
        e = np.exp(z)
        return e / e.sum(-1, keepdims=True)
    p = sm(tl[:, :k])
    q = sm(dl)
    idx = dt[..., None].astype(np.int64)
    px = np.take_along_axis(p, idx, -1)[..., 0]
    qx = np.take_along_axis(q, idx, -1)[..., 0]
    ratio = px / qx
    return ratio >= ua, np.abs(ratio - ua) / np.maximum(ua, 1e-300), p, q


def make_case(B, k, V, seed, T=0.7):
    rng = np.random.default_rng(seed)
    tl = (rng.standard_normal((B, k + 1, V)) * 2).astype(np.float32)
    dl = (tl[:, :k] + rng.standard_normal((B, k, V))).astype(np.float32)
    # draft tokens drawn from q so that accepts and rejects both occur
    z = dl.astype(np.float64) / T
    g = rng.gumbel(size=z.shape)
    dt = np.argmax(z + g, -1).astype(np.int32)
    ua = rng.random((B, k))
    ur = rng.random(B)
    return tl, dl, dt, ua, ur


@pytest.mark.parametrize("B,k,V", [(3, 4, 1024), (2, 5, 20000), (1, 1, 16388), (4, 8, 4100)])
def test_accepts_match_float64(B, k, V):
    T = 0.7
    tl, dl, dt, ua, ur = make_case(B, k, V, seed=B * 100 + k)
    out = oracle.reject_sample(tl, dl, dt, ua, ur, T)
    acc64, margin, p, q = f64_reference(tl, dl, dt, ua, T)
    # raw (un-prefixed) decisions agree wherever the decision margin exceeds fp32 error
    for b in range(B):
        n = int(out["accepted_len"][b])
        assert out["accept_mask"][b, :n].all() and not out["accept_mask"][b, n:].any()
        lead = np.cumprod(acc64[b]).astype(bool)
        safe = margin[b] > 1e-4
        if safe.all():
            assert n == int(lead.sum())
        assert (out["out_tokens"][b, :n] == dt[b, :n]).all()
        assert out["out_tokens"][b, n] >= 0 and (out["out_tokens"][b, n + 1:] == -1).all()
        y = out["out_tokens"][b, n]
        if n < k:   # resampled token must carry residual mass
            assert p[b, n, y] > q[b, n, y] * (1 - 1e-4)


def test_features_match_float64():
    T = 0.7
    tl, dl, dt, ua, ur = make_case(2, 3, 8192, seed=5)
    out = oracle.reject_sample(tl, dl, dt, ua, ur, T)
    z = tl.astype(np.float64) / T
    lse = np.log(np.exp(z - z.max(-1, keepdims=True)).sum(-1)) + z.max(-1)
    p = np.exp(z - lse[..., None])
    srt = np.sort(p, -1)
    ent = -(p * np.log(p)).sum(-1)
    f = out["features"]
    np.testing.assert_allclose(f[..., 0], lse, rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(f[..., 1], srt[..., -1], rtol=5e-6)
    np.testing.assert_allclose(f[..., 2], srt[..., -1] - srt[..., -2], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(f[..., 3], ent, rtol=1e-5)
    lpx = np.log(np.take_along_axis(p[:, :3], dt[..., None].astype(np.int64), -1)[..., 0])
    np.testing.assert_allclose(f[:, :3, 4], lpx, rtol=1e-5, atol=1e-5)


def test_all_accept_and_all_reject():
    B, k, V = 2, 4, 4096
    rng = np.random.default_rng(0)
    tl = (rng.standard_normal((B, k + 1, V)) * 2).astype(np.float32)
    dt = rng.integers(0, V, (B, k)).astype(np.int32)
    ur = rng.random(B)
    # q == p  ->  ratio 1 >= u always
    out = oracle.reject_sample(tl, tl[:, :k].copy(), dt, rng.random((B, k)), ur, 0.7)
    assert (out["accepted_len"] == k).all() and out["accept_mask"].all()
    assert (out["out_tokens"][:, k] >= 0).all()
    # disjoint support: draft mass on tokens the target gives ~0
    dl = np.full((B, k, V), -30.0, np.float32)
    tl2 = np.full((B, k + 1, V), -30.0, np.float32)
    dl[..., :10] = 5.0
    tl2[..., 100:110] = 5.0
    dt2 = rng.integers(0, 10, (B, k)).astype(np.int32)
    out = oracle.reject_sample(tl2, dl, dt2, rng.random((B, k)) * 0.9 + 0.05, ur, 0.7)
    assert (out["accepted_len"] == 0).all() and not out["accept_mask"].any()
    assert ((out["out_tokens"][:, 0] >= 100) & (out["out_tokens"][:, 0] < 110)).all()
    assert (out["out_tokens"][:, 1:] == -1).all()


def test_greedy():
    B, k, V = 3, 4, 5000
    rng = np.random.default_rng(3)
    tl = rng.standard_normal((B, k + 1, V)).astype(np.float32)
    am = tl.argmax(-1).astype(np.int32)
    dt = am[:, :k].copy()
    dt[1, 2] = (dt[1, 2] + 1) % V          # first mismatch at position 2
    dt[2, 0] = (dt[2, 0] + 7) % V
    out = oracle.reject_sample(tl, None, dt, np.zeros((B, k)), np.zeros(B), 0.0)
    assert out["accepted_len"].tolist() == [k, 2, 0]
    assert out["out_tokens"][0].tolist() == am[0].tolist()
    assert out["out_tokens"][1].tolist() == [am[1, 0], am[1, 1], am[1, 2], -1, -1]
    assert out["out_tokens"][2].tolist() == [am[2, 0], -1, -1, -1, -1]


def test_resample_distribution():
    """bonus-row sampling (k = 0) follows softmax(z/T): chi-square on a small vocab."""
    V, N, T = 8, 4000, 1.0
    z = np.array([0.0, 1.0, 2.0, -1.0, 0.5, 1.5, -2.0, 0.25], np.float32)
    tl = np.tile(z, (N, 1, 1))
    ur = np.random.default_rng(11).random(N)
    out = oracle.reject_sample(tl, None, np.zeros((N, 0), np.int32), np.zeros((N, 0)), ur, T)
    cnt = np.bincount(out["out_tokens"][:, 0], minlength=V)
    p = np.exp(z.astype(np.float64)); p /= p.sum()
    chi2 = ((cnt - N * p) ** 2 / (N * p)).sum()
    assert chi2 < 30.0   # 7 dof; P(chi2 > 30) ~ 1e-4


def test_exp2p_accuracy():
    ts = np.linspace(-60, 0, 4001)
    got = np.array([oracle.exp2p(t) for t in ts])
    np.testing.assert_allclose(got, 2.0 ** ts.astype(np.float32).astype(np.float64), rtol=4e-7)
    assert oracle.exp2p(-1000.0) > 0.0


def test_bad_vocab_rejected():
    with pytest.raises(ValueError):
        oracle.reject_sample(np.zeros((1, 1, 6), np.float32), None, np.zeros((1, 0), np.int32),
                             np.zeros((1, 0)), np.zeros(1), 1.0)
