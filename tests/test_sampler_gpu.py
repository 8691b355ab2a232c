"""GPU parity of the fused rejection-sampling kernel against the C oracle: accept masks,
accepted lengths, emitted tokens bit-exact; features within fp32 tolerance (stated below)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

FEAT_RTOL, FEAT_ATOL = 2e-5, 2e-6   # log/div on device vs libm; everything else is bit-exact


def make_case(B, k, V, seed, T, spread=2.0, noise=1.0):
    rng = np.random.default_rng(seed)
    tl = (rng.standard_normal((B, k + 1, V)) * spread).astype(np.float32)
    dl = (tl[:, :k] + rng.standard_normal((B, k, V)) * noise).astype(np.float32)
    Tn = max(T, 1e-3)
    dt = np.argmax(dl.astype(np.float64) / Tn + rng.gumbel(size=dl.shape), -1).astype(np.int32)
    return tl, dl, dt, rng.random((B, k)), rng.random(B)


def run_gpu(tl, dl, dt, ua, ur, T):
    import torch
    from asd_b200.ops import RejectionSampler
    B, k1, V = tl.shape
    s = RejectionSampler(B, k1 - 1)
    out = None
    for _ in range(2):   # second call checks the self-resetting workspace
        out = s(torch.from_numpy(tl).cuda(), None if dl is None else torch.from_numpy(dl).cuda(),
                torch.from_numpy(dt).cuda(), torch.from_numpy(ua).cuda(), torch.from_numpy(ur).cuda(), T)
        torch.cuda.synchronize()
    return {k_: v.cpu().numpy() for k_, v in out.items()}


def compare(got, ref):
    for key in ("accept_mask", "accepted_len", "out_tokens"):
        assert np.array_equal(got[key], ref[key]), key
    np.testing.assert_allclose(got["out_logprobs"], ref["out_logprobs"], rtol=FEAT_RTOL, atol=2e-5)
    g, r = got["features"], ref["features"]
    fin = np.isfinite(r)
    assert np.array_equal(np.isfinite(g), fin)
    np.testing.assert_allclose(g[fin], r[fin], rtol=FEAT_RTOL, atol=2e-5)


@pytest.mark.parametrize("B,k,V,T", [
    (1, 1, 152064, 0.7), (3, 4, 152064, 0.7), (16, 5, 152064, 0.7), (2, 8, 151936, 1.0),
    (5, 3, 4096, 0.7), (2, 2, 16388, 0.3), (7, 1, 32768, 1.5), (1, 8, 212992, 0.7), (40, 2, 50000, 0.7)])
def test_parity_sampling(B, k, V, T):
    tl, dl, dt, ua, ur = make_case(B, k, V, seed=B * 1000 + k * 10 + V % 7, T=T)
    compare(run_gpu(tl, dl, dt, ua, ur, T), oracle.reject_sample(tl, dl, dt, ua, ur, T))


def test_tail_chunks_and_u_extremes():
    """vocabularies that end inside a chunk / inside a lane group, and uniforms at the ends of [0, 1]"""
    for V in (4, 4100, 8188, 4096 * 3 + 36):
        tl, dl, dt, ua, ur = make_case(3, 2, V, seed=V, T=0.9)
        ur[0], ur[1] = 0.0, 1.0
        ua[0, 0], ua[1, 1] = 0.0, 1.0
        compare(run_gpu(tl, dl, dt, ua, ur, 0.9), oracle.reject_sample(tl, dl, dt, ua, ur, 0.9))


def test_parity_all_accept_all_reject_and_bad_tokens():
    B, k, V, T = 4, 5, 152064, 0.7
    tl, _, dt, ua, ur = make_case(B, k, V, seed=1, T=T)
    dl = tl[:, :k].copy()                      # q == p: always accept, residual mass is zero
    got, ref = run_gpu(tl, dl, dt, ua, ur, T), oracle.reject_sample(tl, dl, dt, ua, ur, T)
    compare(got, ref)
    assert (got["accepted_len"] == k).all()
    dl = np.full((B, k, V), -30.0, np.float32)  # disjoint support: always reject
    tl2 = np.full((B, k + 1, V), -30.0, np.float32)
    dl[..., :10] = 5.0
    tl2[..., 1000:1010] = 5.0
    dt2 = (dt % 10).astype(np.int32)
    got, ref = run_gpu(tl2, dl, dt2, ua * 0.9 + 0.05, ur, T), oracle.reject_sample(tl2, dl, dt2, ua * 0.9 + 0.05, ur, T)
    compare(got, ref)
    assert (got["accepted_len"] == 0).all()
    dt3 = dt.copy()                             # out-of-range draft ids are rejected, not read
    dt3[0, 1], dt3[1, 0] = -5, V + 3
    tl, dl, _, ua, ur = make_case(B, k, V, seed=2, T=T)
    compare(run_gpu(tl, dl, dt3, ua, ur, T), oracle.reject_sample(tl, dl, dt3, ua, ur, T))


@pytest.mark.parametrize("B,k,V", [(3, 4, 152064), (1, 8, 151936), (6, 2, 8192)])
def test_parity_greedy(B, k, V):
    tl, _, _, ua, ur = make_case(B, k, V, seed=B + k, T=1.0)
    dt = tl.argmax(-1).astype(np.int32)[:, :k].copy()
    if B > 1:
        dt[1, k // 2] = (dt[1, k // 2] + 1) % V
    tl[0, 0, 7] = tl[0, 0, 9] = tl[0, 0].max() + 1.0   # tie: lowest index wins
    dt[0, 0] = 7
    compare(run_gpu(tl, None, dt, ua, ur, 0.0), oracle.reject_sample(tl, None, dt, ua, ur, 0.0))


def test_parity_row_sampling_k0_and_host_entry():
    from asd_b200.ops import reject_sample_host
    B, V, T = 33, 152064, 0.7
    tl, _, _, _, ur = make_case(B, 0, V, seed=9, T=T)
    dt, ua = np.zeros((B, 0), np.int32), np.zeros((B, 0))
    ref = oracle.reject_sample(tl, None, dt, ua, ur, T)
    compare(run_gpu(tl, None, dt, ua, ur, T), ref)
    compare(reject_sample_host(tl, None, dt, ua, ur, T), ref)


def test_first_reject_prefix_property_full_size():
    """size-independent property at the BASELINE config-2 maximum (B=256, k=8): accepted prefix
    is a prefix, emitted tokens = draft prefix + one token, rest -1; draft == target accepts all."""
    import torch
    from asd_b200.ops import RejectionSampler
    B, k, V, T = 256, 8, 152064, 0.7
    g = torch.Generator(device="cuda").manual_seed(4321)
    tl = torch.randn(B, k + 1, V, device="cuda", generator=g) * 2
    dl = tl[:, :k] + torch.randn(B, k, V, device="cuda", generator=g)
    dt = torch.argmax(dl / T - torch.log(-torch.log(torch.rand(B, k, V, device="cuda", generator=g))), -1).int()
    ua = torch.rand(B, k, dtype=torch.float64, device="cuda", generator=g)
    ur = torch.rand(B, dtype=torch.float64, device="cuda", generator=g)
    s = RejectionSampler(B, k)
    out = {k_: v.cpu().numpy() for k_, v in s(tl.contiguous(), dl.contiguous(), dt, ua, ur, T).items()}
    n = out["accepted_len"]
    assert ((0 <= n) & (n <= k)).all() and 0 < n.sum() < B * k
    ar = np.arange(k)[None]
    assert np.array_equal(out["accept_mask"].astype(bool), ar < n[:, None])
    toks, dtn = out["out_tokens"], dt.cpu().numpy()
    ar1 = np.arange(k + 1)[None]
    assert np.array_equal(toks[:, :k][ar < n[:, None]], dtn[ar < n[:, None]])
    assert (toks[ar1 == n[:, None]] >= 0).all() and (toks[ar1 > n[:, None]] == -1).all()
    out2 = s(tl.contiguous(), tl[:, :k].contiguous(), dt, ua, ur, T)
    assert (out2["accepted_len"].cpu().numpy() == k).all()
    # spot-check 4 sequences of the full-size run against the oracle
    idx = [0, 77, 130, 255]
    ref = oracle.reject_sample(tl[idx].cpu().numpy(), dl[idx].cpu().numpy(), dtn[idx], ua[idx].cpu().numpy(),
                               ur[idx].cpu().numpy(), T)
    assert np.array_equal(ref["accepted_len"], n[idx]) and np.array_equal(ref["out_tokens"], toks[idx])


@pytest.mark.parametrize("B", [64, 128, 256])
def test_full_batch_every_row_against_oracle(B):
    """BASELINE config-2 batch sizes at k = 8, V = 152064: EVERY sequence compared with the C oracle"""
    k, V, T = 8, 152064, 0.7
    tl, dl, dt, ua, ur = make_case(B, k, V, seed=500 + B, T=T)
    compare(run_gpu(tl, dl, dt, ua, ur, T), oracle.reject_sample(tl, dl, dt, ua, ur, T))
