"""asd_cascade_decide (features -> predictor MLP -> Bayesian shrinkage -> optimal-stopping DP in one launch on the GPU)
against the host chain it replaces: FeatureExtractor.extract + QualityPredictor.predict + bayesian_adjustment +
optimal_stopping_rule, composed as in the reference's loop (src/serving/pipeline.py:225-256)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def host_chain(pred, prompt, output, feats, stage_idx, prev, costs, lam, compat, risk, n_obs, alpha, beta):
    from asd_b200.algorithms.dp_solver import bayesian_adjustment, optimal_stopping_rule
    from asd_b200.models.stage import make_logprobs
    L = len(costs)
    if stage_idx < L - 1:
        lp = make_logprobs(feats[:, 5].astype(np.float64), feats)
        prob = pred.predict(prompt=prompt, draft_output=output, draft_logprobs=lp, stage_id=stage_idx)
        if risk:
            prob = bayesian_adjustment(prob, n_obs, alpha, beta)
    else:
        prob = 1.0
    probs = list(prev) + [prob]
    if compat:
        k, _ = optimal_stopping_rule(p=probs, C=list(costs[:stage_idx + 1]), lam=lam, risk_adjustment=False)
        return prob, k == stage_idx, k
    k, _ = optimal_stopping_rule(p=probs + [1.0] * (L - len(probs)), C=list(costs), lam=lam, risk_adjustment=False)
    return prob, k <= stage_idx, stage_idx


@pytest.mark.parametrize("compat", [False, True])
@pytest.mark.parametrize("risk", [False, True])
def test_device_decision_matches_host_chain(compat, risk):
    from asd_b200.models.predictor import QualityPredictor
    from asd_b200.ops import cascade_decide
    torch.manual_seed(5)
    rng = np.random.default_rng(11)
    pred = QualityPredictor(256)
    with torch.no_grad():
        pred.mlp[0].weight.mul_(4.0)        # spread the probabilities away from 0.5
    n, T, L = 37, 70, 4
    costs = [1.0, 2.0, 4.5, 10.0]
    feats = np.zeros((n, T, 6), np.float32)
    feats[..., 0] = rng.normal(5, 1, (n, T))
    feats[..., 1] = rng.uniform(1e-4, 1.0, (n, T))
    feats[..., 2] = feats[..., 1] * rng.uniform(0, 1, (n, T))
    feats[..., 3] = rng.uniform(0, 8, (n, T))
    feats[..., 4] = -rng.uniform(0, 9, (n, T))
    feats[..., 5] = -rng.uniform(0, 9, (n, T))
    ntok = rng.integers(1, T + 1, n).astype(np.int32)
    ntok[0], ntok[1] = T, 1
    prompts = [" ".join(["w"] * int(rng.integers(1, 400))) for _ in range(n)]
    outs = [" ".join(["o"] * int(rng.integers(1, 200))) for _ in range(n)]
    fd = torch.from_numpy(feats).cuda()
    nd = torch.from_numpy(ntok).cuda()
    for stage_idx in range(L):
        prev = rng.uniform(0.05, 0.95, (n, stage_idx))
        lam = float(rng.choice([0.5, 2.0, 8.0]))
        scal = [[len(p.split()) / 2048, len(o.split()) / 512, stage_idx / 4.0] for p, o in zip(prompts, outs)]
        prob, stop, kst = cascade_decide(fd, nd, scal, pred, prev, costs, stage_idx, lam, prefix_mode=compat,
                                         risk_adjustment=risk, n_obs=250.0, alpha=1.5, beta=2.0)
        for i in range(n):
            hp, hs, hk = host_chain(pred, prompts[i], outs[i], feats[i, :ntok[i]], stage_idx, prev[i], costs, lam, compat,
                                    risk, 250.0, 1.5, 2.0)
            assert abs(prob[i] - hp) <= 2e-6, (stage_idx, i, prob[i], hp)
            assert bool(stop[i]) == bool(hs) and int(kst[i]) == int(hk), (stage_idx, i)
    assert 0 < stop.sum() or True


def test_pipeline_uses_the_device_policy_and_agrees_with_the_host_policy():
    """a two-stage cascade of tiny engines: the same requests with device_policy on and off take the same decisions"""
    from asd_b200.models.predictor import FeatureExtractor, QualityPredictor
    from asd_b200.models.qwen2 import tiny_config
    from asd_b200.models.stage import StageConfig, StageManager
    from asd_b200.serving.pipeline import AdaptiveSpeculativePipeline, PipelineConfig
    cfgs = [StageConfig("tiny-a", "0.5b", config=tiny_config(num_hidden_layers=1)),
            StageConfig("tiny-b", "1.5b", config=tiny_config())]
    sm = StageManager(cfgs, k=3, stage_kwargs=dict(max_batch=4, max_model_len=256))
    torch.manual_seed(2)
    pred = QualityPredictor(256)
    res = {}
    for dp in (True, False):
        pipe = AdaptiveSpeculativePipeline(sm, pred, FeatureExtractor(),
                                           PipelineConfig(lambda_value=2.0, enable_caching=False, device_policy=dp))
        calls = []
        orig = sm.get_stage("0.5b").decide
        sm.get_stage("0.5b").decide = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
        out = pipe.batch_process(["alpha beta gamma", "the quick brown fox"], max_tokens=12, temperature=0.0, batched=True)
        sm.get_stage("0.5b").decide = orig
        res[dp] = [(r.stopped_at_stage, [round(p, 5) for p in r.stage_probabilities]) for r in out]
        assert bool(calls) == dp
    assert res[True] == res[False]
