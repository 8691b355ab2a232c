"""CPU oracle: test infrastructure only.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package ``asd_b200``
never does.  The C restatements (``stop_rule_oracle.c``, ``sampler_oracle.c``) are built
into ``oracle/liboracle.so`` by ``oracle/Makefile`` (``__graft_entry__.build()`` runs it).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NFEAT = 6


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("stop_rule_oracle.c", "sampler_oracle.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        d, i, vp = ctypes.c_double, ctypes.c_int, ctypes.c_void_p
        L.oracle_bayesian_adjustment.restype = d
        L.oracle_bayesian_adjustment.argtypes = [d, d, d, d]
        L.oracle_optimal_stopping_rule.restype = i
        L.oracle_optimal_stopping_rule.argtypes = [vp, vp, i, d, i, d, d, vp]
        L.oracle_compute_expected_cost.restype = d
        L.oracle_compute_expected_cost.argtypes = [vp, vp, d, i]
        L.oracle_derive_optimal_policy.restype = None
        L.oracle_derive_optimal_policy.argtypes = [vp, vp, i, d, vp]
        L.oracle_reject_sample.restype = i
        L.oracle_reject_sample.argtypes = [vp, vp, vp, vp, vp, i, i, i, ctypes.c_float, vp, vp, vp, vp, vp]
        L.oracle_exp2p.restype = ctypes.c_float
        L.oracle_exp2p.argtypes = [ctypes.c_float]
        L.oracle_wsum.restype = ctypes.c_float
        L.oracle_wsum.argtypes = [vp]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def bayesian_adjustment(p_hat, n_obs, alpha=1.0, beta=1.0) -> float:
    return float(lib().oracle_bayesian_adjustment(float(p_hat), float(n_obs), float(alpha), float(beta)))


def optimal_stopping_rule(p, C, lam, risk_adjustment=False, alpha=1.0, beta=1.0):
    if len(p) != len(C):
        raise ValueError("p and C must have the same length")
    pa = np.ascontiguousarray(p, dtype=np.float64)
    ca = np.ascontiguousarray(C, dtype=np.float64)
    J = np.zeros(len(ca) + 1, dtype=np.float64)
    k = lib().oracle_optimal_stopping_rule(_p(pa), _p(ca), len(ca), float(lam), int(bool(risk_adjustment)),
                                           float(alpha), float(beta), _p(J))
    if k < 0:
        raise ValueError("bad stage count")
    return int(k), [float(x) for x in J]


def compute_expected_cost(p, C, lam, stopping_stage) -> float:
    pa = np.ascontiguousarray(p, dtype=np.float64)
    ca = np.ascontiguousarray(C, dtype=np.float64)
    return float(lib().oracle_compute_expected_cost(_p(pa), _p(ca), float(lam), int(stopping_stage)))


def derive_optimal_policy(quality_bounds, cost_ratios, lam):
    q = np.ascontiguousarray(quality_bounds, dtype=np.float64)
    c = np.ascontiguousarray(cost_ratios, dtype=np.float64)
    th = np.zeros(len(q), dtype=np.float64)
    lib().oracle_derive_optimal_policy(_p(q), _p(c), len(q), float(lam), _p(th))
    return {s: float(th[s]) for s in range(len(q))}


def reject_sample(target_logits, draft_logits, draft_tokens, u_accept, u_resid, temperature):
    """target_logits [B,k+1,V] f32; draft_logits [B,k,V] f32 or None; draft_tokens [B,k] i32;
    u_accept [B,k] f64; u_resid [B] f64.  Returns dict of numpy outputs."""
    tl = np.ascontiguousarray(target_logits, dtype=np.float32)
    B, k1, V = tl.shape
    k = k1 - 1
    dl = None if draft_logits is None else np.ascontiguousarray(draft_logits, dtype=np.float32)
    dt = np.ascontiguousarray(draft_tokens, dtype=np.int32).reshape(B, k)
    ua = np.ascontiguousarray(u_accept, dtype=np.float64).reshape(B, k)
    ur = np.ascontiguousarray(u_resid, dtype=np.float64).reshape(B)
    if dl is not None:
        assert dl.shape == (B, k, V)
    elif k > 0 and temperature > 0:
        raise ValueError("draft_logits required when temperature > 0 and k > 0")
    out = dict(
        accept_mask=np.zeros((B, k), np.uint8), accepted_len=np.zeros(B, np.int32),
        out_tokens=np.zeros((B, k + 1), np.int32), out_logprobs=np.zeros((B, k + 1), np.float32),
        features=np.zeros((B, k + 1, NFEAT), np.float32))
    rc = lib().oracle_reject_sample(_p(tl), _p(dl), _p(dt), _p(ua), _p(ur), B, k, V, float(temperature),
                                    _p(out["accept_mask"]), _p(out["accepted_len"]), _p(out["out_tokens"]),
                                    _p(out["out_logprobs"]), _p(out["features"]))
    if rc != 0:
        raise ValueError("oracle_reject_sample: bad arguments")
    return out


def exp2p(t: float) -> float:
    return float(lib().oracle_exp2p(float(t)))
