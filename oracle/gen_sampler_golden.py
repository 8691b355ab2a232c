"""Golden vectors that pin oracle/sampler_oracle.c to an INDEPENDENT implementation of speculative rejection
sampling: vLLM's ``vllm/v1/sample/rejection_sampler.py`` (the engine the reference's Stage wraps,
/root/reference/src/serving/real_model_pipeline.py:98-108; the reference itself has no token-level sampler).

vLLM's sampler is a set of Triton kernels; this script runs them ON THE CPU through the Triton interpreter
(``TRITON_INTERPRET=1``) with OUR uniforms and records, per case, what vLLM decided:
  * ``rejection_random_sample_kernel`` (accept iff draft_prob > 0 and target_prob / draft_prob >= uniform, :810):
    the accepted length of every sequence and the ratio p/q at every draft token (so that the test can set aside
    the decisions that sit within fp32 rounding of a tie, where two correct implementations may differ);
  * ``rejection_greedy_sample_kernel`` (:745): the full output row of every sequence (draft prefix that matches the
    arg-max, first mismatch replaced by the arg-max, bonus token after a full accept).
Inputs are NOT stored: they are regenerated from the recorded seeds with ``numpy.random.default_rng`` (a float64
checksum of each array is stored to notice a change of the generator).  tests/test_oracle_sampler_vllm.py replays the
cases through the C oracle in the CPU suite; tests/test_sampler_vllm_gpu.py compares the CUDA kernel with the same
vLLM kernels on the GPU.

    python oracle/gen_sampler_golden.py        # needs vllm + triton (this image), ~1 minute, writes tests/golden/
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "sampler_vllm_golden.json")

RANDOM_CASES = [  # (B, k, V, T, seed, noise)
    (8, 4, 152064, 0.7, 101, 1.0), (16, 5, 152064, 0.7, 102, 1.0), (6, 8, 151936, 1.0, 103, 0.5),
    (32, 8, 32000, 0.3, 104, 1.0), (5, 3, 4096, 1.0, 105, 2.0), (64, 1, 8192, 0.7, 106, 1.0),
]
GREEDY_CASES = [(8, 4, 152064, 201), (33, 8, 32000, 202), (5, 3, 4096, 203)]  # (B, k, V, seed)


def make_case(B, k, V, T, seed, noise=1.0):
    """the inputs of one case; every consumer of the golden file must call exactly this"""
    rng = np.random.default_rng(seed)
    tl = (rng.standard_normal((B, k + 1, V)) * 2).astype(np.float32)
    dl = (tl[:, :k] + rng.standard_normal((B, k, V)) * noise).astype(np.float32)
    dt = np.argmax(dl / max(T, 1e-6) + rng.gumbel(size=dl.shape), -1).astype(np.int32)
    ua, ur = rng.random((B, k)), rng.random(B)
    return tl, dl, dt, ua, ur


def make_greedy_case(B, k, V, seed):
    rng = np.random.default_rng(seed)
    tl = (rng.standard_normal((B, k + 1, V)) * 2).astype(np.float32)
    am = tl.argmax(-1).astype(np.int32)
    flip = rng.random((B, k)) < 0.2
    dt = np.where(flip, (am[:, :k] + 1) % V, am[:, :k]).astype(np.int32)
    return tl, dt, am, rng.random((B, k)), rng.random(B)


def checksum(*arrays):
    return [float(np.asarray(a, dtype=np.float64).sum()) for a in arrays]


def main():
    os.environ["TRITON_INTERPRET"] = "1"      # before triton is imported: run vLLM's kernels on the CPU
    import torch
    from vllm.v1.sample import rejection_sampler as rs
    import vllm
    gold = {"source": f"vllm {vllm.__version__} vllm/v1/sample/rejection_sampler.py, Triton interpreter on CPU",
            "random": [], "greedy": []}
    for B, k, V, T, seed, noise in RANDOM_CASES:
        tl, dl, dt, ua, ur = make_case(B, k, V, T, seed, noise)
        p = torch.softmax(torch.from_numpy(tl[:, :k].reshape(B * k, V)) / T, -1, dtype=torch.float32).contiguous()
        q = torch.softmax(torch.from_numpy(dl.reshape(B * k, V)) / T, -1, dtype=torch.float32).contiguous()
        out = torch.full((B, k + 1), -1, dtype=torch.int32)
        cu = torch.arange(k, (B + 1) * k, k, dtype=torch.int32)
        bonus = torch.full((B, 1), 7, dtype=torch.int32)
        recovered = torch.full((B * k,), -7, dtype=torch.int32)      # marks the first rejected position
        rs.rejection_random_sample_kernel[(B,)](out, cu, torch.from_numpy(dt.reshape(-1)), q, p, bonus, recovered,
                                                torch.from_numpy(ua.reshape(-1)), torch.zeros(B, dtype=torch.bool), k, V,
                                                None, NO_DRAFT_PROBS=False, SYNTHETIC_MODE=False)
        o = out.numpy()
        n = [next((i for i in range(k) if o[b, i] == -7), k) for b in range(B)]
        for b in range(B):
            assert np.array_equal(o[b, :n[b]], dt[b, :n[b]])
        rows = np.arange(B * k)
        ratio = (p.numpy()[rows, dt.reshape(-1)].astype(np.float64) / q.numpy()[rows, dt.reshape(-1)].astype(np.float64))
        gold["random"].append({"B": B, "k": k, "V": V, "T": T, "seed": seed, "noise": noise,
                               "checksum": checksum(tl, dl, dt, ua, ur), "accepted_len": [int(x) for x in n],
                               "ratio": [float.hex(float(x)) for x in ratio]})
        print("random", B, k, V, T, "accepted", sum(n), "of", B * k, flush=True)
    for B, k, V, seed in GREEDY_CASES:
        tl, dt, am, ua, ur = make_greedy_case(B, k, V, seed)
        out = torch.full((B, k + 1), -1, dtype=torch.int32)
        cu = torch.arange(k, (B + 1) * k, k, dtype=torch.int32)
        rs.rejection_greedy_sample_kernel[(B,)](out, cu, torch.from_numpy(dt.reshape(-1)),
                                                torch.from_numpy(am[:, :k].reshape(-1).astype(np.int64)),
                                                torch.from_numpy(am[:, k:k + 1].copy()), None, k, None, None,
                                                SYNTHETIC_MODE=False)
        gold["greedy"].append({"B": B, "k": k, "V": V, "seed": seed, "checksum": checksum(tl, dt),
                               "out_tokens": out.numpy().tolist()})
        print("greedy", B, k, V, flush=True)
    with open(OUT, "w") as f:
        json.dump(gold, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
