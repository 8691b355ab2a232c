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
