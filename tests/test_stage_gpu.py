"""End-to-end through the reference-facing API on the GPU: Stage.generate (plain and draft-then-verify),
the cascade pipeline, and the RealModelPipeline surface, on tiny Qwen2-shaped models."""
import numpy as np
import pytest

from asd_b200.models.qwen2 import tiny_config

pytestmark = pytest.mark.gpu


def make_stages():
    from asd_b200.models.stage import Stage
    small = Stage("tiny-draft", "7b", config=tiny_config(num_hidden_layers=1), seed=1, max_batch=4, max_model_len=256, k=3)
    big = Stage("tiny-target", "32b", config=tiny_config(), seed=2, draft=small, max_batch=4, max_model_len=256, k=3)
    return small, big


def test_stage_generate_contract_and_greedy_equivalence():
    from asd_b200.models.stage import Stage
    small, big = make_stages()
    texts, lps, stats = big.generate(["hello world", "hello there", "a much longer prompt here"], max_tokens=24,
                                     temperature=0.0)
    assert len(texts) == 3 and all(isinstance(t, str) for t in texts)
    assert all(lp.shape == (24, 5) and lp.fused.shape == (24, 6) for lp in lps)
    assert stats["generation_time_ms"] > 0 and stats["decode_steps"] > 0
    assert np.isfinite(np.asarray(lps[0])[:, :3]).all() and (np.asarray(lps[0])[1:, 0] <= 1e-6).all()
    # greedy speculative output == greedy output of the same target without a draft
    plain = Stage("tiny-target", "32b", config=tiny_config(), seed=2, max_batch=4, max_model_len=256)
    t2, _, _ = plain.generate(["hello world", "hello there", "a much longer prompt here"], max_tokens=24, temperature=0.0)
    assert t2 == texts
    info = big.get_model_info()
    assert info["draft"] == "7b" and info["cost_per_token"] == 4.5 and big.compute_kv_cache_size(1000) > 0


def test_pipeline_end_to_end_on_engines():
    from asd_b200.models.predictor import FeatureExtractor, QualityPredictor
    from asd_b200.serving.pipeline import AdaptiveSpeculativePipeline, PipelineConfig
    small, big = make_stages()

    class Mgr:
        def get_stage(self, name):
            return {"7b": small, "32b": big}[name]

        def stage_names(self):
            return ["7b", "32b"]

    pipe = AdaptiveSpeculativePipeline(Mgr(), QualityPredictor(256), FeatureExtractor(),
                                       PipelineConfig(lambda_value=100.0, risk_adjustment=True))
    r = pipe.process_request("What is the capital of France?", max_tokens=12, temperature=0.7)
    assert r.stopped_at_stage in (0, 1) and isinstance(r.output, str) and len(r.stage_probabilities) >= 1
    assert all(0.0 <= p <= 1.0 for p in r.stage_probabilities)
    pipe.update_lambda(0.0)          # free quality: always stop at the first stage
    assert pipe.process_request("hi", max_tokens=8).stopped_at_stage == 0
    pipe.shutdown()


def test_real_model_pipeline_surface():
    from asd_b200.serving.real_model_pipeline import InferenceRequest, RealModelPipeline, StageConfig
    scs = [StageConfig("qwen-7b", "none", 1, [0], 256, "bfloat16"), StageConfig("qwen-32b", "none", 1, [0], 256, "bfloat16")]
    pipe = RealModelPipeline(scs, stage_kwargs=dict(config=tiny_config(), max_batch=2, k=2))
    with pytest.raises(RuntimeError):
        pipe.infer_adaptive(InferenceRequest("x", 4))
    pipe.initialize()
    res = pipe.infer_adaptive(InferenceRequest("hello", max_tokens=8, temperature=0.0, request_id="r1"))
    assert res.request_id == "r1" and res.selected_stage in (0, 1) and len(res.stage_results) >= 1
    assert all(r.error is None for r in res.stage_results) and isinstance(res.output, str)
    assert pipe.get_statistics()["total_requests"] == 1
    pipe.cleanup()


def test_calibrate_costs_replaces_the_cost_table():
    """SURVEY 8 f3 (real_model_pipeline.py:313-362): measured per-token time, normalised to the first stage"""
    from asd_b200.models.stage import StageConfig, StageManager
    scs = [StageConfig("tiny-a", "7b", config=tiny_config(num_hidden_layers=1)),
           StageConfig("tiny-b", "32b", config=tiny_config(num_hidden_layers=2))]
    mgr = StageManager(scs, k=2, stage_kwargs=dict(max_batch=4, max_model_len=128))
    before = [mgr.get_stage(n).cost_per_token for n in mgr.stage_names()]
    assert before == [1.0, 4.5]
    table = mgr.calibrate_costs(max_tokens=8)
    assert list(table) == ["7b", "32b"] and table["7b"] == 1.0 and table["32b"] > 0
    assert [mgr.get_stage(n).cost_per_token for n in mgr.stage_names()] == [table["7b"], table["32b"]]
    # the pipeline reads the calibrated costs
    from asd_b200.models.predictor import FeatureExtractor, QualityPredictor
    from asd_b200.serving.pipeline import AdaptiveSpeculativePipeline, PipelineConfig
    pipe = AdaptiveSpeculativePipeline(mgr, QualityPredictor(256), FeatureExtractor(), PipelineConfig(lambda_value=1e9))
    r = pipe.process_request("x", max_tokens=4)
    assert r.stage_costs[0] == 1.0
    pipe.shutdown()


def test_two_stages_generate_concurrently_without_corrupting_the_shared_engine():
    """a stage's engine is also the next stage's draft: concurrent generate() calls on both (the pipeline's thread
    pool does that) must give exactly the outputs of the same calls made one after the other"""
    from concurrent.futures import ThreadPoolExecutor
    small, big = make_stages()
    prompts = ["alpha beta", "gamma delta epsilon", "zeta"]
    want_small = small.generate(prompts, max_tokens=20, temperature=0.0)[0]
    want_big = big.generate(prompts, max_tokens=20, temperature=0.0)[0]
    with ThreadPoolExecutor(4) as ex:
        futs = [ex.submit((small if i % 2 else big).generate, prompts, 20, 0.0) for i in range(8)]
        got = [f.result()[0] for f in futs]
    for i, g in enumerate(got):
        assert g == (want_small if i % 2 else want_big), i


def test_generation_refuses_to_run_past_max_model_len():
    from asd_b200._lib import AsdError
    from asd_b200.engine import QwenEngine, SpecDecoder
    import torch
    cfg = tiny_config()
    t = QwenEngine(cfg, max_seqs=2, max_seq_len=64, max_tokens=32).load_random(1)
    d = QwenEngine(cfg, max_seqs=2, max_seq_len=48, max_tokens=32).load_random(1)     # draft with a shorter limit
    dec = SpecDecoder(t, d, 2, 3, temperature=0.0)
    dec.prefill(torch.randint(0, cfg.vocab_size, (2, 30)))
    with pytest.raises(AsdError, match="max_seq_len"):
        for _ in range(10):
            dec.step()
    with pytest.raises(AsdError, match="does not fit"):
        SpecDecoder(t, d, 2, 3, temperature=0.0).prefill(torch.randint(0, cfg.vocab_size, (2, 60)))
    # ... while generate() freezes finished sequences and stays inside the limit
    toks, lps, feats, st = SpecDecoder(t, d, 2, 3, temperature=0.0).generate([[1, 2, 3], [4, 5, 6, 7, 8, 9]], 30)
    assert toks.shape == (2, 30) and (toks >= 0).all() and st["steps"] <= 31


def test_stage_loads_safetensors_checkpoint(tmp_path):
    from asd_b200.models.qwen2 import random_hf_weights, save_safetensors
    from asd_b200.models.stage import ModelLoadError, Stage
    cfg = tiny_config()
    w = random_hf_weights(cfg, seed=8)
    save_safetensors(w, str(tmp_path / "model.safetensors"))
    a = Stage(str(tmp_path), "7b", config=cfg, max_batch=2, max_model_len=128)
    b = Stage("in-memory", "7b", config=cfg, weights=w, max_batch=2, max_model_len=128)
    assert a.weights_source.startswith("safetensors:") and b.weights_source == "state_dict"
    assert a.generate(["same prompt"], 12, 0.0)[0] == b.generate(["same prompt"], 12, 0.0)[0]
    with pytest.raises(ModelLoadError, match="random weights are not allowed"):
        Stage("no/such/dir", "7b", config=cfg, allow_random_weights=False)
