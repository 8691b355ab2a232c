"""Training rows for the quality predictor, emitted from what the B200 engine computes.

Mirrors the on-disk format of the reference's src/training/generate_training_data.py: a JSON list of
``TrainingSample`` records (:35-47, written at :348-356) plus ``feature_stats.json`` (:361-371), and its 64-d
``extract_features`` vector (:148-205) - with one difference that is the point of this path: the per-token
logprobs come from the fused sampling kernel (``Stage.generate`` -> ``FusedLogprobs``) instead of a
``softmax`` + ``log(probs[token]).item()`` host loop (:128-134), and four otherwise-unused padding slots carry
the kernel's full-vocabulary statistics (mean entropy, mean p_max, mean margin, min emitted logprob) so the MLP
can be trained on exactly what the device produces.  Quality labels need reference outputs (the reference
scores BLEU >= 0.7, :208-225); ``quality_fn`` is supplied by the caller because the ``evaluate`` package and
its metric files are not available offline."""
from __future__ import annotations

import json
import os
import time
from dataclasses import asdict, dataclass
from typing import Callable, Dict, List, Optional

import numpy as np


@dataclass
class TrainingSample:
    prompt: str
    stage_id: int
    model_output: str
    reference_output: str
    features: List[float]
    quality_score: float
    bleu_score: float
    generation_time: float
    prompt_tokens: int
    completion_tokens: int


def extract_features(prompt: str, output: str, metadata: Dict, stage_id: int) -> List[float]:
    """generate_training_data.py:148-205, same slots 0..21; slots 22..25 = fused-kernel statistics."""
    f: List[float] = []
    pw, ow = prompt.split(), output.split()
    f += [len(pw), len(prompt), len(ow), len(output), len(ow) / max(len(pw), 1)]
    lps = list(metadata.get("logprobs", []))
    if lps:
        f += [float(np.mean(lps)), float(np.std(lps)), float(np.min(lps)), float(np.percentile(lps, 25)),
              float(np.median(lps))]
    else:
        f += [0.0] * 5
    f.append(len(set(ow)) / max(len(ow), 1))
    one_hot = [0.0] * 4
    one_hot[min(stage_id, 3)] = 1.0
    f += one_hot
    f.append(metadata.get("completion_tokens", 0) / max(metadata.get("generation_time", 1.0), 0.001))
    f.append(int("def " in prompt or "```" in prompt or "import " in prompt))
    f.append(int(any(c in prompt for c in "+=*/<>")))
    f.append(sum(1 for w in ["what", "why", "how", "when", "where", "which"] if w in prompt.lower()))
    fused = metadata.get("fused")
    if fused is not None and len(fused):
        fu = np.asarray(fused, dtype=np.float64)
        f += [float(fu[:, 3].mean()), float(fu[:, 1].mean()), float(fu[:, 2].mean()), float(np.min(lps)) if lps else 0.0]
    while len(f) < 64:
        f.append(0.0)
    return [float(x) for x in f[:64]]


def generate_samples(stages, prompts: List[str], references: Optional[List[str]] = None,
                     quality_fn: Optional[Callable[[str, str], float]] = None, max_tokens: int = 64,
                     temperature: float = 0.7) -> List[TrainingSample]:
    """one sample per (prompt, stage); ``stages`` is a list of asd_b200.models.stage.Stage"""
    references = references or [""] * len(prompts)
    quality_fn = quality_fn or (lambda out, ref: 0.0)
    samples = []
    for sid, stage in enumerate(stages):
        for prompt, ref in zip(prompts, references):
            t0 = time.time()
            texts, lps, _ = stage.generate([prompt], max_tokens=max_tokens, temperature=temperature)
            dt = time.time() - t0
            lp = np.asarray(lps[0])
            meta = {"logprobs": lp[:, 0].tolist() if lp.size else [], "generation_time": dt,
                    "completion_tokens": int(lp.shape[0]) if lp.size else 0, "fused": getattr(lps[0], "fused", None)}
            q = float(quality_fn(texts[0], ref))
            samples.append(TrainingSample(prompt, sid, texts[0], ref, extract_features(prompt, texts[0], meta, sid),
                                          float(q >= 0.7), q, dt, len(stage._encode(prompt)), meta["completion_tokens"]))
    return samples


def save_training_data(samples: List[TrainingSample], output_dir: str) -> str:
    """training_data.json + feature_stats.json, as generate_training_data.py:348-371"""
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, "training_data.json")
    with open(path, "w") as fh:
        json.dump([asdict(s) for s in samples], fh, indent=2)
    feats = np.array([s.features for s in samples], dtype=np.float64)
    with open(os.path.join(output_dir, "feature_stats.json"), "w") as fh:
        json.dump({"mean": feats.mean(0).tolist(), "std": feats.std(0).tolist(), "min": feats.min(0).tolist(),
                   "max": feats.max(0).tolist()}, fh, indent=2)
    return path
