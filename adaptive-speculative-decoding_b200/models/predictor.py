"""Scorer seam.  ``QualityPredictor`` / ``FeatureExtractor`` are imported by the reference
(src/serving/pipeline.py:15,225-231; src/serving/server.py:24,168-179) from ``src/models/predictor.py``,
which it does not ship; reconstructed from docs/guides/RESEARCH_PROTOCOL.md:308-409.

The feature vector keeps the listing's 5 live dimensions in the listing's order (entropy, prompt length
/ 2048, output length / 512, mean max-logprob, stage / 4, zero-padded to 256).  When the logprobs come
from the B200 engine (``FusedLogprobs.fused``) entropy and max-logprob are the sampling kernel's
full-vocabulary values (no top-5 truncation, no logits round trip) and three extra dimensions carry
the fused margin, mean emitted-token logprob and its minimum (the logprob statistics of
src/training/generate_training_data.py:148-205)."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn as nn


class FeatureExtractor:
    FEATURE_DIM = 256

    def extract(self, prompt: str, draft_output: str, draft_logprobs, stage_id: int) -> np.ndarray:
        features = []
        fused = getattr(draft_logprobs, "fused", None)
        n = 0 if draft_logprobs is None else len(draft_logprobs)
        if fused is not None and len(fused) > 0:
            f = np.asarray(fused, dtype=np.float64)
            features.append(float(np.mean(f[-32:, 3])))                     # full-vocab entropy, last 32 tokens
        elif n > 0:
            features.append(float(-np.mean([np.sum(np.exp(lp) * lp) for lp in draft_logprobs[-32:]])))  # :379-384
        else:
            features.append(0.0)
        features.append(len(prompt.split()) / 2048)                          # :389
        features.append(len(draft_output.split()) / 512)                     # :393
        if fused is not None and len(fused) > 0:
            features.append(float(np.mean(np.log(np.clip(f[:, 1], 1e-30, 1.0)))))
        elif n > 0:
            features.append(float(np.mean([np.max(lp) for lp in draft_logprobs])))   # :396-398
        else:
            features.append(-10.0)
        features.append(stage_id / 4.0)                                      # :403
        if fused is not None and len(fused) > 0:
            lp_tok = np.asarray(draft_logprobs)[:, 0].astype(np.float64)
            features += [float(np.mean(f[:, 2])), float(np.mean(lp_tok)), float(np.min(lp_tok))]
        features = np.array(features, dtype=np.float64)
        return np.pad(features, (0, self.FEATURE_DIM - len(features)), "constant")   # :406-407


class QualityPredictor(nn.Module):
    """Linear(feature_dim, 128)-ReLU-Dropout(0.1)-Linear(128, 1)-Sigmoid (RESEARCH_PROTOCOL.md:325-331).
    ``predict`` accepts the pipeline's keyword call (pipeline.py:225-231) and the positional
    single-vector call of real_model_pipeline.py:406."""

    def __init__(self, feature_dim=256):
        super().__init__()
        if isinstance(feature_dim, dict):                # variant B: QualityPredictor(config: dict)
            feature_dim = int(feature_dim.get("feature_dim", feature_dim.get("input_dim", 128)))
        self.feature_dim = feature_dim
        self.feature_extractor = FeatureExtractor()
        self.mlp = nn.Sequential(nn.Linear(feature_dim, 128), nn.ReLU(), nn.Dropout(0.1), nn.Linear(128, 1),
                                 nn.Sigmoid())
        self.feature_cache = {}
        self.eval()

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        return self.mlp(features)

    def load_model(self, path: str):
        self.load_state_dict(torch.load(path, map_location="cpu"))
        self.eval()

    def predict(self, prompt=None, draft_output: Optional[str] = None, draft_logprobs=None, stage_id: int = 0,
                feature_extractor: Optional[FeatureExtractor] = None) -> float:
        if isinstance(prompt, np.ndarray) and draft_output is None:      # predict(features)
            features = prompt
        else:
            fx = feature_extractor or self.feature_extractor
            features = fx.extract(prompt, draft_output, draft_logprobs, stage_id)
        features = np.asarray(features, dtype=np.float32)
        if features.shape[-1] != self.feature_dim:
            features = np.pad(features, (0, max(0, self.feature_dim - features.shape[-1])))[:self.feature_dim]
        with torch.no_grad():
            return float(self.mlp(torch.from_numpy(features)).item())
