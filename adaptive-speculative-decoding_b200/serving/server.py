"""HTTP front end of the cascade, mirroring the reference's FastAPI surface
(`src/serving/server.py:40-84` schemas, `:223-376` endpoints) on top of `asd_b200.serving.pipeline`:

    GET  /health         POST /generate        POST /batch_generate   GET /stats
    POST /update_lambda  POST /reset_stats     GET  /models           GET /cache_stats

Same field names, defaults and validation bounds, same status codes (503 before the pipeline exists, 500 with
the exception text when a request fails, 404 for /cache_stats without a cache manager).  Unlike the reference,
which builds one module-level app around globals filled by a lifespan hook, the app comes from a factory so that
tests (and embedders) hand it a ready pipeline; `create_app_from_config` is the lifespan path
(`server.py:92-200`: yaml -> KVCacheManager -> StageManager -> QualityPredictor -> pipeline -> warm-up).
"""
from __future__ import annotations

import logging
import os
import time
from typing import Any, Dict, List, Optional

from fastapi import FastAPI, HTTPException
from pydantic import BaseModel, Field

logger = logging.getLogger(__name__)


# ---------------------------------------------------------------------------------------------- schemas
class GenerationRequest(BaseModel):           # server.py:40-47
    prompt: str = Field(..., description="Input prompt")
    max_tokens: int = Field(512, ge=1, le=2048, description="Maximum tokens to generate")
    temperature: float = Field(0.7, ge=0.0, le=2.0, description="Generation temperature")
    top_p: float = Field(0.9, ge=0.0, le=1.0, description="Top-p sampling")
    stream: bool = Field(False, description="Whether to stream the response")
    request_id: Optional[str] = Field(None, description="Optional request ID")


class GenerationResponse(BaseModel):          # server.py:50-60
    request_id: str
    output: str
    stopped_at_stage: int
    latency_ms: float
    stage_probabilities: List[float]
    stage_costs: List[float]
    total_tokens: int
    tokens_per_second: float
    cache_hits: int


class BatchGenerationRequest(BaseModel):      # server.py:63-67
    prompts: List[str]
    max_tokens: int = Field(512, ge=1, le=2048)
    temperature: float = Field(0.7, ge=0.0, le=2.0)


class LambdaUpdateRequest(BaseModel):         # server.py:70-72
    lambda_value: float = Field(..., ge=0.01, le=100.0, description="New lambda value")


class StatsResponse(BaseModel):               # server.py:75-84
    total_requests: int
    stage_distribution: List[float]
    avg_latency: float
    avg_tokens_per_second: float
    avg_tokens_per_request: float
    active_requests: int
    error_count: int
    cache_stats: Optional[Dict[str, Any]] = None


_RESPONSE_FIELDS = tuple(GenerationResponse.model_fields)


def _as_response(result) -> GenerationResponse:
    return GenerationResponse(**{name: getattr(result, name) for name in _RESPONSE_FIELDS})


# ---------------------------------------------------------------------------------------------- app factory
def create_app(pipeline=None, cache_manager=None, model_sizes: Optional[List[str]] = None) -> FastAPI:
    """The reference's endpoints around `pipeline` (an AdaptiveSpeculativePipeline or anything with its methods).
    `pipeline=None` reproduces the not-yet-initialised server (every pipeline endpoint answers 503)."""
    app = FastAPI(title="Adaptive Speculative Decoding API",
                  description="Multi-stage draft-verify pipeline with the optimal-stopping rule (B200 engine)",
                  version="1.0.0")
    app.state.pipeline = pipeline
    app.state.cache_manager = cache_manager

    def live():
        if app.state.pipeline is None:
            raise HTTPException(status_code=503, detail="Pipeline not initialized")
        return app.state.pipeline

    @app.get("/health")
    async def health():
        import torch
        gpu = torch.cuda.is_available()
        return {"status": "healthy", "timestamp": time.time(), "gpu_available": gpu,
                "gpu_count": torch.cuda.device_count() if gpu else 0}

    @app.post("/generate", response_model=GenerationResponse)
    async def generate(request: GenerationRequest):
        p = live()
        try:
            result = await p.process_request_async(prompt=request.prompt, max_tokens=request.max_tokens,
                                                   temperature=request.temperature, request_id=request.request_id)
            return _as_response(result)
        except Exception as e:  # the reference maps every failure to a 500 carrying the message (server.py:261-263)
            logger.error(f"Generation failed: {e}")
            raise HTTPException(status_code=500, detail=str(e))

    @app.post("/batch_generate")
    async def batch_generate(request: BatchGenerationRequest):
        p = live()
        try:
            results = p.batch_process(prompts=request.prompts, max_tokens=request.max_tokens,
                                      temperature=request.temperature)
            return {"results": [_as_response(r) for r in results]}
        except Exception as e:
            logger.error(f"Batch generation failed: {e}")
            raise HTTPException(status_code=500, detail=str(e))

    @app.get("/stats", response_model=StatsResponse)
    async def stats():
        s = live().get_stats()
        return StatsResponse(cache_stats=s.get("cache_stats"),
                             **{k: s[k] for k in StatsResponse.model_fields if k != "cache_stats"})

    @app.post("/update_lambda")
    async def update_lambda(request: LambdaUpdateRequest):
        p = live()
        try:
            old = p.config.lambda_value
            p.update_lambda(request.lambda_value)
            return {"message": "Lambda updated successfully", "old_lambda": old, "new_lambda": request.lambda_value}
        except Exception as e:
            logger.error(f"Lambda update failed: {e}")
            raise HTTPException(status_code=500, detail=str(e))

    @app.post("/reset_stats")
    async def reset_stats():
        live().reset_stats()
        return {"message": "Statistics reset successfully"}

    @app.get("/models")
    async def models():
        p = live()
        sizes = model_sizes
        if sizes is None:       # the reference hard-codes its four size labels; ask the manager when it can tell
            names = getattr(p.stage_manager, "stage_names", None)
            sizes = list(names()) if callable(names) else ["8b", "13b", "34b", "70b"]
        info = {}
        for size in sizes:
            try:
                info[size] = p.stage_manager.get_stage(size).get_model_info()
            except Exception as e:
                info[size] = {"error": str(e)}
        return {"models": info}

    @app.get("/cache_stats")
    async def cache_stats():
        if app.state.cache_manager is None:
            raise HTTPException(status_code=404, detail="Cache manager not available")
        return app.state.cache_manager.get_stats()

    return app


def create_app_from_config(config: Dict[str, Any]) -> FastAPI:
    """Start-up path of the reference (`server.py:92-200`) from an already parsed serving config:
    `models.stages[*]` -> StageConfig, `pipeline.{lambda_value, risk_adjustment, cache}` -> PipelineConfig."""
    from ..models.predictor import FeatureExtractor, QualityPredictor
    from ..models.stage import StageConfig, StageManager
    from .cache_manager import KVCacheManager
    from .pipeline import AdaptiveSpeculativePipeline, PipelineConfig

    pcfg = config.get("pipeline", {})
    cache_cfg = pcfg.get("cache", {})
    cache = None
    if cache_cfg.get("enable_kv_cache", True):
        cache = KVCacheManager(max_cache_size_gb=cache_cfg.get("max_cache_size_gb", 40),
                               cleanup_interval=cache_cfg.get("cache_cleanup_interval", 300))
    stage_cfgs, alloc = [], {}
    for st in config.get("models", {}).get("stages", []):
        q = st.get("quantized", st.get("quantization", False))
        if isinstance(q, dict):
            q = bool(q.get("enabled", False))
        stage_cfgs.append(StageConfig(st.get("model_name", st.get("model_path", st.get("name", ""))),
                                      st.get("size", st.get("size_label")), int(st.get("tensor_parallel_size", 1)),
                                      st.get("gpu_memory_utilization", st.get("gpu_memory_fraction", 0.8)), bool(q),
                                      st.get("cost_per_token", 1.0), max_model_len=st.get("max_model_len")))
        if st.get("gpu_ids") is not None:
            alloc[stage_cfgs[-1].model_size] = list(st["gpu_ids"])
    manager = StageManager(stage_cfgs, alloc, stage_kwargs=config.get("engine", {}).get("stage_kwargs"))
    pred_cfg = config.get("predictor", {})
    predictor = QualityPredictor(feature_dim=pred_cfg.get("feature_dim", 256))
    ckpt = pred_cfg.get("checkpoint", "checkpoints/predictor.pt")         # server.py:170-176
    if ckpt and os.path.exists(ckpt):
        predictor.load_model(ckpt)
        logger.info(f"Loaded predictor from {ckpt}")
    else:
        logger.warning(f"Predictor weights not found at {ckpt}, using random weights")
    risk = pcfg.get("risk_adjustment", {})                                # server.py:183-193
    if isinstance(risk, dict):
        risk_on, risk_a, risk_b = bool(risk.get("enabled", True)), float(risk.get("alpha", 1.0)), float(risk.get("beta", 1.0))
    else:
        risk_on, risk_a, risk_b = bool(risk), float(pcfg.get("risk_alpha", 1.0)), float(pcfg.get("risk_beta", 1.0))
    pipe = AdaptiveSpeculativePipeline(
        manager, predictor, FeatureExtractor(),
        PipelineConfig(lambda_value=pcfg.get("lambda_value", 1.0), risk_adjustment=risk_on, risk_alpha=risk_a,
                       risk_beta=risk_b, enable_caching=cache_cfg.get("enable_kv_cache", True),
                       max_concurrent_requests=config.get("safety", {}).get("max_concurrent_requests", 100)),
        cache_manager=cache)
    if pcfg.get("warmup", True):
        pipe.warmup()
    return create_app(pipe, cache)


def main():   # server.py:379-404: uvicorn, single worker
    import uvicorn
    import yaml
    with open(os.getenv("CONFIG_PATH", "configs/serving.yaml")) as f:
        config = yaml.safe_load(f)
    server = config.get("server", {})
    uvicorn.run(create_app_from_config(config), host=server.get("host", "0.0.0.0"), port=server.get("port", 8000),
                workers=1, log_level="info")


if __name__ == "__main__":
    main()
