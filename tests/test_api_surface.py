"""Host-side mirror of the reference API: pipeline control flow against goldens produced by the
reference's UNMODIFIED src/serving/pipeline.py (oracle/gen_golden.py), threshold policy, cache manager."""
import asyncio
import types

import numpy as np
import pytest

from asd_b200.algorithms import dp_solver
from asd_b200.serving.cache_manager import KVCacheManager
from asd_b200.serving.pipeline import (AdaptiveSpeculativePipeline, PipelineConfig, RequestResult)
from asd_b200.theory.optimal_stopping import OptimalStoppingTheory, TheoreticalParameters
from conftest import fh


class FakeStage:
    def __init__(self, name, cost):
        self.name, self.cost_per_token, self.calls = name, cost, 0

    def generate(self, prompts, max_tokens=512, temperature=0.7, top_p=0.9, return_logprobs=True):
        self.calls += 1
        return ([f"{self.name}-says: " + p[:16] for p in prompts], [np.full((3, 5), -1.0)], {"generation_time_ms": 1.0})


class FakeManager:
    def __init__(self, names=("8b", "13b", "34b", "70b"), costs=(1.0, 1.6, 4.2, 8.8)):
        self.stages = {n: FakeStage(n, c) for n, c in zip(names, costs)}
        self._names = list(names)

    def get_stage(self, name):
        return self.stages[name]

    def stage_names(self):
        return list(self._names)


class SeqPredictor:
    def __init__(self, seq):
        self.seq, self.i = seq, 0

    def predict(self, prompt, draft_output, draft_logprobs, stage_id, feature_extractor=None):
        v = self.seq[self.i % len(self.seq)]
        self.i += 1
        return v


def test_reference_compat_matches_unmodified_reference_pipeline(stop_rule_golden):
    for g in stop_rule_golden["pipeline"]:
        cfg = PipelineConfig(lambda_value=fh(g["lam"]), risk_adjustment=g["risk_adjustment"], enable_caching=False,
                             reference_compat=True)
        pipe = AdaptiveSpeculativePipeline(FakeManager(), SeqPredictor([fh(x) for x in g["predictions"]]), None, cfg)
        r = pipe.process_request("What is the capital of France?", max_tokens=8, temperature=0.7, request_id="golden")
        pipe.shutdown()
        assert r.output == g["output"] and r.stopped_at_stage == g["stopped_at_stage"]
        assert [float(x).hex() for x in r.stage_probabilities] == g["stage_probabilities"]
        assert [float(x).hex() for x in r.stage_costs] == g["stage_costs"]
        assert r.total_tokens == g["total_tokens"] and r.cache_hits == g["cache_hits"]


def test_default_mode_escalates_and_stops_with_the_full_vector_rule():
    mgr = FakeManager(("7b", "32b", "72b"), (1.0, 4.5, 10.0))
    # confident draft: stop at stage 0
    pipe = AdaptiveSpeculativePipeline(mgr, SeqPredictor([0.99]), None, PipelineConfig(lambda_value=1.0,
                                                                                       risk_adjustment=False))
    assert pipe.process_request("q").stopped_at_stage == 0
    # unsure draft, high quality weight: escalate.  The rule's penalty lam * (1 - prod p) can only shrink by
    # reaching the last stage (J[L] = 0), so once stage 0 continues the cascade runs to the end.
    pipe = AdaptiveSpeculativePipeline(mgr, SeqPredictor([0.2, 0.95]), None, PipelineConfig(lambda_value=20.0,
                                                                                            risk_adjustment=False))
    r = pipe.process_request("q")
    assert r.stopped_at_stage == 2 and r.output.startswith("72b-says") and r.stage_costs == [1.0, 4.5, 10.0]
    assert mgr.stages["32b"].calls == 1
    # nobody is confident: run to the last stage
    pipe = AdaptiveSpeculativePipeline(mgr, SeqPredictor([0.1, 0.1]), None, PipelineConfig(lambda_value=50.0,
                                                                                           risk_adjustment=False))
    r = pipe.process_request("q")
    assert r.stopped_at_stage == 2 and r.stage_probabilities[-1] == 1.0
    # decisions agree with the DP evaluated by hand
    k, _ = dp_solver.optimal_stopping_rule([0.2, 1.0, 1.0], [1.0, 4.5, 10.0], 20.0)
    assert k > 0
    stats = pipe.get_stats()
    assert stats["total_requests"] == 1 and stats["stage_stops"][2] == 1 and "cache_stats" in stats
    assert r["latency_ms"] == r.latency_ms and r["costs"] == r.stage_costs
    pipe.update_lambda(3.0)
    assert pipe.lambda_value == 3.0
    pipe.reset_stats()
    assert pipe.get_stats()["total_requests"] == 0
    pipe.shutdown()


def test_async_batch_and_error_path():
    mgr = FakeManager(("7b", "32b"), (1.0, 4.5))
    pipe = AdaptiveSpeculativePipeline(mgr, SeqPredictor([0.9]), None, PipelineConfig())
    r = asyncio.run(pipe.process_request_async("hello"))
    assert isinstance(r, RequestResult)
    assert len(pipe.batch_process(["a", "b", "c"])) == 3

    class Boom(FakeStage):
        def generate(self, *a, **k):
            raise RuntimeError("engine failure")
    mgr.stages["7b"] = Boom("7b", 1.0)
    with pytest.raises(RuntimeError):
        pipe.process_request("x")
    assert pipe.get_stats()["error_count"] == 1 and pipe.get_stats()["active_requests"] == 0
    pipe.shutdown()


def test_cache_manager_api():
    cm = KVCacheManager(max_cache_size_gb=1e-6, cleanup_interval=3600)      # ~1 KB budget
    assert cm.allocate("r1", 0, {"output": "x" * 300, "logprobs": np.zeros(10)})
    assert cm.get_cache("r1", 0)["output"].startswith("x") and cm.get_cache("r1", 1) is None
    assert cm.allocate("r2", 0, {"output": "y" * 600})
    assert cm.allocate("r3", 0, {"output": "z" * 600})                      # forces LRU eviction
    assert cm.get_stats()["current_size_bytes"] <= cm.max_cache_size_bytes
    assert not cm.allocate("big", 0, {"output": "w" * 5000})
    cm.allocate("r4", 0, {"output": "a"}); cm.allocate("r4", 1, {"output": "b"}); cm.allocate("r4", 2, {"output": "c"})
    cm.truncate_at_stage("r4", 0)
    assert cm.get_cache("r4", 0) is not None and cm.get_cache("r4", 2) is None
    cm.cleanup_request("r4")                                                # str payloads do not crash
    st = cm.get_stats()
    assert st["hit_rate"] > 0 and "utilization" in st
    cm.shutdown()


def test_threshold_policy_matches_reference(stop_rule_golden):
    for g in stop_rule_golden["policy"]:
        q, c = [fh(x) for x in g["quality_bounds"]], [fh(x) for x in g["cost_ratios"]]
        th = OptimalStoppingTheory(TheoreticalParameters(n_stages=len(q), quality_bounds=q, cost_ratios=c,
                                                         lambda_param=fh(g["lam"]))).derive_optimal_policy()
        assert {str(s): float(v).hex() for s, v in th.items()} == g["thresholds"]
    d = OptimalStoppingTheory(TheoreticalParameters()).derive_optimal_policy()
    assert d == {3: 0, 2: -0.7076923076923076, 1: -0.4714285714285714, 0: -0.09999999999999998}


def test_feature_extractor_and_predictor_follow_the_listing():
    from asd_b200.models.predictor import FeatureExtractor, QualityPredictor
    from asd_b200.models.stage import make_logprobs
    lp = np.log(np.array([[0.5, 0.2, 0.1, 0.1, 0.1], [0.9, 0.05, 0.02, 0.02, 0.01]]))
    f = FeatureExtractor().extract("a b c", "d e", lp, 2)
    assert f.shape == (256,) and (f[5:] == 0).all()
    ent = -np.mean([np.sum(np.exp(r) * r) for r in lp])
    np.testing.assert_allclose(f[:5], [ent, 3 / 2048, 2 / 512, np.mean(lp.max(1)), 0.5])
    assert FeatureExtractor().extract("a", "", np.zeros((0, 5)), 0)[3] == -10.0
    fused = np.array([[1.0, 0.5, 0.3, 1.2, -1.0, -0.7], [1.0, 0.9, 0.85, 0.3, -0.1, -0.1]], np.float32)
    flp = make_logprobs(np.array([-0.7, -0.1]), fused)
    assert flp.shape == (2, 5) and flp.fused is not None and flp[-32:].fused is not None
    g = FeatureExtractor().extract("a b c", "d e", flp, 1)
    np.testing.assert_allclose(g[0], 0.75, rtol=1e-6)          # mean fused entropy
    np.testing.assert_allclose(g[5:8], [0.575, -0.4, -0.7], rtol=1e-6)
    p = QualityPredictor(feature_dim=256)
    v = p.predict(prompt="a b", draft_output="c", draft_logprobs=flp, stage_id=0, feature_extractor=FeatureExtractor())
    assert 0.0 < v < 1.0
    assert sum(x.numel() for x in p.parameters()) == 33025
    assert 0.0 < QualityPredictor({"feature_dim": 128}).predict(np.zeros(128, np.float32)) < 1.0


def test_install_as_src_aliases():
    import asd_b200
    asd_b200.install_as_src()
    from src.algorithms.dp_solver import optimal_stopping_rule
    from src.serving.pipeline import AdaptiveSpeculativePipeline as P2
    from src.models.stage import Stage, StageManager  # noqa: F401
    from src.models.predictor import QualityPredictor, FeatureExtractor  # noqa: F401
    assert P2 is AdaptiveSpeculativePipeline and optimal_stopping_rule([1.0], [1.0], 5.0) == (0, [1.0, 0.0])


def test_lambda_optimizer_follows_the_reference_search():
    from asd_b200.algorithms.optimizer import GridSearchOptimizer, LambdaOptimizer, find_optimal_lambda
    # latency falls and quality falls as lambda grows (cheaper stages chosen): bisection finds the boundary
    ev = lambda lam: (1000.0 / (1.0 + lam), 1.0 / (1.0 + 0.1 * lam))
    r = LambdaOptimizer(latency_constraint=100.0, lambda_bounds=(0.01, 100.0)).optimize_for_latency_constraint(ev)
    assert r.constraint_satisfied and abs(r.optimal_lambda - 9.0) < 0.01 and r.iterations <= 50
    with pytest.raises(ValueError):
        LambdaOptimizer().optimize_for_latency_constraint(ev)
    front = LambdaOptimizer().optimize_pareto_front(ev, num_points=8)
    assert len(front) == 8 and all(front[i][1] <= front[i + 1][1] for i in range(7))
    b = LambdaOptimizer(lambda_bounds=(0.01, 10.0)).find_balanced_lambda(ev, quality_weight=0.5)
    assert 0.01 <= b.optimal_lambda <= 10.0
    mgr = FakeManager(("7b", "32b"), (1.0, 4.5))
    pipe = AdaptiveSpeculativePipeline(mgr, SeqPredictor([0.9]), None, PipelineConfig(enable_caching=False))
    lam = find_optimal_lambda(pipe, ["a b c", "d e f"], "latency", 1e9, num_evaluations=2)
    assert 0.01 <= lam <= 100.0 and pipe.lambda_value > 0
    g = GridSearchOptimizer([0.5, 5.0]).search(pipe, [{"prompt": "x y"}])
    assert g["best_lambda"] in (0.5, 5.0) and set(g["results"]) == {0.5, 5.0}
    pipe.shutdown()


def test_training_feature_vector_matches_the_reference_layout(tmp_path):
    from asd_b200.training.generate_training_data import TrainingSample, extract_features, save_training_data
    meta = {"logprobs": [-0.5, -1.5, -0.25, -2.0], "generation_time": 0.5, "completion_tokens": 4}
    f = extract_features("how does import x work = ?", "it works it works", meta, 2)
    assert len(f) == 64
    lp = np.array(meta["logprobs"])
    np.testing.assert_allclose(f[:11], [7, 26, 4, 17, 4 / 7, lp.mean(), lp.std(), lp.min(), np.percentile(lp, 25),
                                        np.median(lp), 0.5])
    assert f[11:15] == [0.0, 0.0, 1.0, 0.0] and f[15] == 8.0 and f[16:19] == [1.0, 1.0, 1.0] and f[22:] == [0.0] * 42
    fused = np.array([[1, 0.5, 0.3, 1.2, 0, 0], [1, 0.9, 0.8, 0.3, 0, 0]], np.float32)
    g = extract_features("a", "b", dict(meta, fused=fused), 0)
    np.testing.assert_allclose(g[19:23], [0.75, 0.7, 0.55, -2.0], rtol=1e-6)
    s = TrainingSample("p", 0, "o", "r", f, 1.0, 0.8, 0.1, 3, 4)
    path = save_training_data([s, s], str(tmp_path))
    import json
    rows = json.load(open(path))
    assert len(rows) == 2 and set(rows[0]) == {"prompt", "stage_id", "model_output", "reference_output", "features",
                                              "quality_score", "bleu_score", "generation_time", "prompt_tokens",
                                              "completion_tokens"}
    stats = json.load(open(tmp_path / "feature_stats.json"))
    assert set(stats) == {"mean", "std", "min", "max"} and len(stats["mean"]) == 64


def test_http_server_mirrors_the_reference_endpoints():
    """same routes, schemas, validation bounds and status codes as src/serving/server.py:223-376"""
    from fastapi.testclient import TestClient
    from asd_b200.serving.server import create_app

    # before the pipeline exists every pipeline endpoint answers 503, /cache_stats 404, /health works
    cold = TestClient(create_app(None))
    assert cold.get("/health").json()["status"] == "healthy"
    for method, path, body in (("post", "/generate", {"prompt": "x"}), ("get", "/stats", None),
                               ("post", "/reset_stats", None), ("get", "/models", None),
                               ("post", "/update_lambda", {"lambda_value": 2.0}),
                               ("post", "/batch_generate", {"prompts": ["a"]})):
        r = getattr(cold, method)(path, json=body) if body is not None else getattr(cold, method)(path)
        assert r.status_code == 503, (path, r.status_code)
    assert cold.get("/cache_stats").status_code == 404

    cache = KVCacheManager(max_cache_size_gb=1, cleanup_interval=3600)
    pipe = AdaptiveSpeculativePipeline(FakeManager(), SeqPredictor([0.2, 0.9]), None,
                                       PipelineConfig(lambda_value=50.0, enable_caching=False), cache_manager=cache)
    c = TestClient(create_app(pipe, cache))
    r = c.post("/generate", json={"prompt": "What is the capital of France?", "max_tokens": 8, "request_id": "r-1"})
    assert r.status_code == 200
    body = r.json()
    assert set(body) == {"request_id", "output", "stopped_at_stage", "latency_ms", "stage_probabilities",
                         "stage_costs", "total_tokens", "tokens_per_second", "cache_hits"}
    assert body["request_id"] == "r-1" and body["output"]
    # validation bounds of the reference schemas (422 from pydantic)
    assert c.post("/generate", json={"prompt": "x", "max_tokens": 0}).status_code == 422
    assert c.post("/generate", json={"prompt": "x", "temperature": 2.5}).status_code == 422
    assert c.post("/update_lambda", json={"lambda_value": 0.001}).status_code == 422
    rb = c.post("/batch_generate", json={"prompts": ["a", "b", "c"], "max_tokens": 4})
    assert rb.status_code == 200 and len(rb.json()["results"]) == 3
    st = c.get("/stats").json()
    assert st["total_requests"] == 4 and abs(sum(st["stage_distribution"]) - 1.0) < 1e-9
    up = c.post("/update_lambda", json={"lambda_value": 2.5}).json()
    assert up == {"message": "Lambda updated successfully", "old_lambda": 50.0, "new_lambda": 2.5}
    assert pipe.lambda_value == 2.5
    assert c.post("/reset_stats").json() == {"message": "Statistics reset successfully"}
    assert c.get("/stats").json()["total_requests"] == 0
    models = c.get("/models").json()["models"]
    assert set(models) == {"8b", "13b", "34b", "70b"} and all("error" in v for v in models.values())  # fakes have no info
    assert "total_entries" in c.get("/cache_stats").json() or c.get("/cache_stats").status_code == 200

    class Boom(FakeManager):
        def get_stage(self, name):
            raise RuntimeError("stage exploded")
    bad = AdaptiveSpeculativePipeline(Boom(), SeqPredictor([0.5]), None, PipelineConfig(enable_caching=False))
    r = TestClient(create_app(bad)).post("/generate", json={"prompt": "x"})
    assert r.status_code == 500 and "stage exploded" in r.json()["detail"]
    cache.shutdown()


def test_batched_cascade_matches_the_sequential_decisions():
    """batch_process(batched=True): one generate call per stage for all live requests, same per-request decisions"""
    class ByPrompt:
        def predict(self, prompt, draft_output, draft_logprobs, stage_id, feature_extractor=None):
            base = {"easy": 0.97, "mid": 0.55, "hard": 0.05}[prompt.split()[0]]
            return min(0.99, base + 0.2 * stage_id)
    prompts = ["easy one", "hard two", "mid three", "hard four", "easy five"]
    cfg = dict(lambda_value=20.0, enable_caching=False, risk_adjustment=False)
    seq_mgr, bat_mgr = FakeManager(), FakeManager()
    seq = AdaptiveSpeculativePipeline(seq_mgr, ByPrompt(), None, PipelineConfig(**cfg)).batch_process(prompts, 8)
    bat_pipe = AdaptiveSpeculativePipeline(bat_mgr, ByPrompt(), None, PipelineConfig(**cfg))
    bat = bat_pipe.batch_process(prompts, 8, batched=True)
    assert [r.stopped_at_stage for r in bat] == [r.stopped_at_stage for r in seq]
    assert [r.output for r in bat] == [r.output for r in seq]
    assert [r.stage_probabilities for r in bat] == [r.stage_probabilities for r in seq]
    assert [r.stage_costs for r in bat] == [r.stage_costs for r in seq]
    assert len(set(r.stopped_at_stage for r in bat)) > 1            # the batch really splits across stages
    # one generate call per stage that still had live requests, instead of one per (request, stage)
    assert [bat_mgr.stages[n].calls for n in bat_mgr.stage_names()] == [1 if any(r.stopped_at_stage >= i for r in bat) else 0
                                                                        for i in range(4)]
    assert sum(s.calls for s in seq_mgr.stages.values()) == sum(r.stopped_at_stage + 1 for r in seq)
    assert bat_pipe.get_stats()["total_requests"] == 5 and not bat_pipe.active_requests


def test_predictor_training_on_saved_rows(tmp_path):
    """rows written by save_training_data train the reference's MLP recipe (K-fold, AdamW, MSE)"""
    import json
    import torch
    from asd_b200.training.generate_training_data import TrainingSample, extract_features, save_training_data
    from asd_b200.training.train_predictor import ResearchQualityPredictor, load_training_rows, train_quality_predictor
    rng = np.random.default_rng(0)
    samples = []
    for i in range(240):
        lps = (-rng.random(12) * (0.2 + 3.0 * rng.random())).tolist()
        prompt, out = "what is " + "x " * int(rng.integers(1, 9)), "tok " * int(rng.integers(2, 20))
        feats = extract_features(prompt, out, {"logprobs": lps, "generation_time": 0.1, "completion_tokens": 12}, i % 4)
        q = float(1.0 / (1.0 + np.exp(-(3.0 + 2.0 * np.mean(lps)))))            # quality follows the mean logprob
        samples.append(TrainingSample(prompt, i % 4, out, "", feats, float(q >= 0.7), q, 0.1, 3, 12))
    path = save_training_data(samples, str(tmp_path))
    X, y = load_training_rows(path)
    assert X.shape == (240, 64) and y.shape == (240,)
    cfg = {"predictor": {"model": {"input_dim": 64, "hidden_layers": [32, 16], "dropout": 0.0},
                         "training": {"batch_size": 32, "num_epochs": 40, "learning_rate": 0.01}, "data": {"cv_folds": 3}}}
    res = train_quality_predictor(X, y, cfg)
    cv = res["cross_validation"]
    assert set(cv) == {"mean_r2", "std_r2", "mean_mse", "std_mse", "fold_results"} and len(cv["fold_results"]) == 3
    assert set(cv["fold_results"][0]) == {"fold", "r2_score", "mse", "mae", "best_val_loss"}
    assert cv["mean_r2"] > 0.5, cv                                              # the relation is learnable
    assert res["training_config"]["num_samples"] == 240 and res["model_config"]["architecture"] == "mlp"
    m = ResearchQualityPredictor(64, [32, 16], 0.0)
    m.load_state_dict(res["state_dict"])
    m.eval()
    xs = torch.from_numpy(((X - np.array(res["scaler"]["mean"])) / np.array(res["scaler"]["std"])).astype(np.float32))
    with torch.no_grad():
        p = m(xs).numpy()
    assert float(((p - y) ** 2).mean()) < 0.02
    assert len(ResearchQualityPredictor().network) == 3 * 4 + 2                 # 128 -> 256 -> 128 -> 64 -> 1
    with pytest.raises(ValueError):
        train_quality_predictor(X, y, {"predictor": {"model": {"input_dim": 128}}})
    json.dumps({k: v for k, v in res.items() if k != "state_dict"})              # the report part is JSON-serialisable


def test_gpu_placement_rules_follow_the_reference():
    """src/config/model_config.py:136-150: len(gpu_ids) == tensor_parallel_size, no GPU shared between stages"""
    from asd_b200.models.stage import ModelLoadError, validate_gpu_assignment
    validate_gpu_assignment([("7b", 1, [0]), ("14b", 1, [1]), ("32b", 2, [2, 3]), ("72b", 4, [4, 5, 6, 7])])
    with pytest.raises(ModelLoadError, match="doesn't match tensor_parallel_size"):
        validate_gpu_assignment([("32b", 2, [2])])
    with pytest.raises(ModelLoadError, match="multiple stages"):
        validate_gpu_assignment([("7b", 1, [0]), ("32b", 2, [0, 1])])
    with pytest.raises(ModelLoadError, match="duplicate"):
        validate_gpu_assignment([("32b", 2, [1, 1])])


def test_safetensors_round_trip(tmp_path):
    import torch
    from asd_b200.models.qwen2 import load_safetensors_dir, random_hf_weights, save_safetensors, tiny_config
    w = random_hf_weights(tiny_config(num_hidden_layers=1), seed=4)
    w["extra.f32"] = torch.arange(6, dtype=torch.float32).reshape(2, 3)
    keys = sorted(w)
    save_safetensors({k: w[k] for k in keys[:5]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_safetensors({k: w[k] for k in keys[5:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    back = load_safetensors_dir(str(tmp_path))
    assert sorted(back) == keys
    for k in keys:
        assert back[k].dtype == w[k].dtype and torch.equal(back[k], w[k]), k
    with pytest.raises(FileNotFoundError):
        load_safetensors_dir(str(tmp_path / "nothing"))


def test_generations_sharing_a_gpu_take_the_same_device_lock_in_index_order():
    """Stage.generate serialises generations that touch the same GPU (kernels of different streams are outside the
    forward's TMEM hand-over invariant): one re-entrant lock per device, handed out sorted by
    device index so that two cascades with overlapping GPU sets can never wait for each other in a cycle."""
    import threading
    from asd_b200.models import stage
    a = stage._device_locks([3, 1, 3])
    b = stage._device_locks([1])
    c = stage._device_locks([5, 3])
    assert len(a) == 2 and a[0] is b[0] and a[1] is c[0] and len(c) == 2
    # re-entrant: the draft stage's devices may repeat the target's inside one generation
    a[0].acquire()
    assert a[0].acquire(blocking=False)
    a[0].release()
    a[0].release()
    got = []
    t = threading.Thread(target=lambda: (b[0].acquire(), got.append(1), b[0].release()))
    a[0].acquire()
    t.start()
    t.join(0.2)
    assert not got              # another thread waits while the device is busy
    a[0].release()
    t.join(2)
    assert got == [1]
