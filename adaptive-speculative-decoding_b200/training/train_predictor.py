"""Training of the quality predictor on the rows `generate_training_data.save_training_data` writes
(mirrors the surface of the reference's `scripts/train_predictor.py`: model :57-91, trainer
:261-466, result dictionary :420-458).

Same recipe - standardised features, K-fold cross-validation (shuffle, seed 42), AdamW + cosine schedule, MSE
on the quality score, best-validation checkpoint per fold, R^2 / MSE / MAE per fold - and the same keys in the
returned dictionary, so the reference's reporting code reads it unchanged.  Everything runs on the host (torch
CPU) unless a device is passed: this is the step *after* the hot path, its input is what the fused kernel
emits.  The standardisation constants and the best fold's weights are returned so that the scorer seam
(`asd_b200.models.predictor`) can load them."""
from __future__ import annotations

import json
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import nn


class ResearchQualityPredictor(nn.Module):
    """input -> [Linear, BatchNorm1d, ReLU, Dropout] per hidden width -> Linear(1) -> Sigmoid
    (train_predictor.py:57-91; defaults 128 -> 256 -> 128 -> 64 -> 1)."""

    def __init__(self, input_dim: int = 128, hidden_layers: Optional[List[int]] = None, dropout: float = 0.2):
        super().__init__()
        widths = [input_dim] + list(hidden_layers if hidden_layers is not None else [256, 128, 64])
        blocks: List[nn.Module] = []
        for a, b in zip(widths[:-1], widths[1:]):
            blocks += [nn.Linear(a, b), nn.BatchNorm1d(b), nn.ReLU(), nn.Dropout(dropout)]
        blocks += [nn.Linear(widths[-1], 1), nn.Sigmoid()]
        self.network = nn.Sequential(*blocks)

    def forward(self, x):
        return self.network(x).squeeze(-1)


def load_training_rows(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """features [N, 64] float32 and quality scores [N] float32 from training_data.json"""
    with open(path) as fh:
        rows = json.load(fh)
    feats = np.asarray([r["features"] for r in rows], dtype=np.float32)
    labels = np.asarray([r["bleu_score"] if "bleu_score" in r else r["quality_score"] for r in rows], dtype=np.float32)
    return feats, labels


def _r2(y: np.ndarray, p: np.ndarray) -> float:
    ss_res, ss_tot = float(((y - p) ** 2).sum()), float(((y - y.mean()) ** 2).sum())
    return 1.0 - ss_res / ss_tot if ss_tot > 0 else 0.0


def train_quality_predictor(features: np.ndarray, labels: np.ndarray, config: Dict[str, Any],
                            device: str = "cpu", seed: int = 42) -> Dict[str, Any]:
    pc = config.get("predictor", {})
    mc, tc = pc.get("model", {}), pc.get("training", {})
    input_dim = mc.get("input_dim", features.shape[1])
    hidden, dropout = mc.get("hidden_layers", [256, 128, 64]), mc.get("dropout", 0.2)
    bs, lr = tc.get("batch_size", 256), tc.get("learning_rate", 0.001)
    epochs, wd = tc.get("num_epochs", 100), tc.get("weight_decay", 0.01)
    folds = pc.get("data", {}).get("cv_folds", 5)
    if features.shape[1] != input_dim:
        raise ValueError(f"features have {features.shape[1]} columns, model input_dim is {input_dim}")

    mean, std = features.mean(0), features.std(0)
    std = np.where(std > 0, std, 1.0)                       # StandardScaler's handling of constant columns
    x_all = torch.from_numpy(((features - mean) / std).astype(np.float32))
    y_all = torch.from_numpy(labels.astype(np.float32))

    order = np.random.RandomState(seed).permutation(len(features))   # KFold(shuffle=True, random_state=42) analogue
    chunks = np.array_split(order, folds)
    torch.manual_seed(seed)
    fold_results, fold_states = [], []
    for k in range(folds):
        val_idx = chunks[k]
        tr_idx = np.concatenate([chunks[j] for j in range(folds) if j != k])
        model = ResearchQualityPredictor(input_dim, hidden, dropout).to(device)
        opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=wd)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=epochs)
        loss_fn = nn.MSELoss()
        xtr, ytr = x_all[tr_idx].to(device), y_all[tr_idx].to(device)
        xva, yva = x_all[val_idx].to(device), y_all[val_idx].to(device)
        best, best_state = float("inf"), None
        for _ in range(epochs):
            model.train()
            perm = torch.randperm(len(tr_idx))
            for i in range(0, len(perm), bs):
                b = perm[i:i + bs]
                if len(b) < 2:                               # BatchNorm needs more than one row
                    continue
                opt.zero_grad()
                loss_fn(model(xtr[b]), ytr[b]).backward()
                opt.step()
            sched.step()
            model.eval()
            with torch.no_grad():
                v = float(loss_fn(model(xva), yva))
            if v < best:
                best, best_state = v, {n: t.detach().cpu().clone() for n, t in model.state_dict().items()}
        model.load_state_dict(best_state)
        model.eval()
        with torch.no_grad():
            pred = model(xva).cpu().numpy()
        yv = yva.cpu().numpy()
        fold_results.append({"fold": k + 1, "r2_score": _r2(yv, pred), "mse": float(((pred - yv) ** 2).mean()),
                             "mae": float(np.abs(pred - yv).mean()), "best_val_loss": best})
        fold_states.append(best_state)
    r2s, mses = [f["r2_score"] for f in fold_results], [f["mse"] for f in fold_results]
    best_fold = int(np.argmin(mses))
    return {
        "cross_validation": {"mean_r2": float(np.mean(r2s)), "std_r2": float(np.std(r2s)), "mean_mse": float(np.mean(mses)),
                             "std_mse": float(np.std(mses)), "fold_results": fold_results},
        "model_config": {"architecture": "mlp", "input_dim": input_dim, "hidden_layers": list(hidden), "dropout": dropout,
                         "total_parameters": sum(p.numel() for p in ResearchQualityPredictor(input_dim, hidden, dropout).parameters())},
        "training_config": {"num_samples": int(len(features)), "batch_size": bs, "learning_rate": lr, "num_epochs": epochs,
                            "cv_folds": folds},
        # additions: what a scorer needs to use the result
        "scaler": {"mean": mean.tolist(), "std": std.tolist()},
        "best_fold": best_fold + 1,
        "state_dict": fold_states[best_fold],
    }
