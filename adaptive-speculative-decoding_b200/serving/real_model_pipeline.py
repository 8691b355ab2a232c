"""Second API surface of the reference's engine seam: src/serving/real_model_pipeline.py
(``StageConfig`` :44-52, ``InferenceRequest`` :54-62, ``StageInferenceResult`` :64-76, ``RealModelStage``
:78-190, ``RealModelPipeline`` :192-520, ``create_real_pipeline`` :522-526).  Names and argument meaning are
kept; the body runs on the B200 engine instead of ``vllm.LLM`` (:98-108).  ``SamplingParams``,
``PipelineResult`` and ``DynamicProgrammingSolver`` are names the reference imports but never defines
(SURVEY.md Appendix A); minimal definitions live here / in algorithms.dp_solver."""
from __future__ import annotations

import logging
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

from ..algorithms.dp_solver import DynamicProgrammingSolver
from ..models.predictor import QualityPredictor
from ..models.stage import COST_PER_TOKEN, InferenceError, ModelLoadError, Stage


@dataclass
class SamplingParams:
    temperature: float = 0.7
    top_p: float = 0.9
    max_tokens: int = 512


@dataclass
class StageConfig:
    name: str
    model_path: str
    tensor_parallel_size: int = 1
    gpu_ids: List[int] = field(default_factory=lambda: [0])
    max_model_len: int = 4096
    dtype: str = "bfloat16"


@dataclass
class InferenceRequest:
    prompt: str
    max_tokens: int
    temperature: float = 0.7
    top_p: float = 0.9
    lambda_param: float = 1.0
    request_id: str = ""


@dataclass
class StageInferenceResult:
    stage_id: int
    stage_name: str
    output_text: str
    output_tokens: List[int]
    inference_time: float
    gpu_memory_used: float
    quality_score: float
    confidence_score: float
    should_continue: bool
    error: Optional[str] = None


@dataclass
class PipelineResult:
    request_id: str
    prompt: str
    output: str
    selected_stage: int
    stage_results: List[StageInferenceResult]
    total_inference_time: float
    total_pipeline_time: float
    lambda_param: float
    quality_score: float


def _size_from_name(name: str) -> str:
    n = name.lower()
    for s in ("0.5b", "1.5b", "72b", "70b", "34b", "32b", "14b", "13b", "8b", "7b"):
        if s in n:
            return s
    raise ModelLoadError(f"cannot infer a model size from stage name {name!r}")


class RealModelStage:
    def __init__(self, config: StageConfig, stage_id: int, draft: Optional["RealModelStage"] = None,
                 stage_kwargs: Optional[dict] = None):
        self.config, self.stage_id = config, stage_id
        self.model: Optional[Stage] = None
        self.is_loaded = False
        self._draft, self._kw = draft, dict(stage_kwargs or {})
        self.logger = logging.getLogger(f"stage_{stage_id}")

    def load_model(self):
        if self.is_loaded:
            return
        if self.config.dtype not in ("bfloat16", "bf16", "auto"):
            raise ModelLoadError(f"dtype {self.config.dtype!r} unsupported: the engine computes in bf16")
        draft = None
        if self._draft is not None:
            self._draft.load_model()
            draft = self._draft.model
        self.model = Stage(self.config.model_path, _size_from_name(self.config.name),
                           tensor_parallel_size=self.config.tensor_parallel_size, draft=draft,
                           max_model_len=self.config.max_model_len,
                           gpu_ids=list(self.config.gpu_ids) if self.config.gpu_ids else None, **self._kw)
        self.is_loaded = True

    def unload_model(self):
        if self.model is not None:
            self.model.engine.close()
        self.model, self.is_loaded = None, False

    def infer(self, prompt: str, sampling_params: SamplingParams) -> StageInferenceResult:
        """real_model_pipeline.py:117-181: runtime errors become ``result.error``."""
        if not self.is_loaded:
            raise InferenceError(f"Stage {self.stage_id} not loaded")
        import torch
        t0 = time.perf_counter()
        try:
            torch.cuda.synchronize()
            texts, lps, _ = self.model.generate([prompt], max_tokens=sampling_params.max_tokens,
                                                temperature=sampling_params.temperature, top_p=sampling_params.top_p)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            conf = float(np.exp(np.mean(np.asarray(lps[0])[:, 0]))) if len(lps[0]) else 0.0
            return StageInferenceResult(self.stage_id, self.config.name, texts[0],
                                        self.model._encode(texts[0]), dt,
                                        torch.cuda.memory_allocated() / 1e9, 0.0, conf, True)
        except Exception as e:
            return StageInferenceResult(self.stage_id, self.config.name, "", [], time.perf_counter() - t0, 0.0, 0.0,
                                        0.0, True, error=str(e))


class RealModelPipeline:
    def __init__(self, stage_configs: List[StageConfig], lambda_param: float = 1.0, speculative: bool = True,
                 predictor: Optional[QualityPredictor] = None, stage_kwargs: Optional[dict] = None):
        self.stages: List[RealModelStage] = []
        if any(sc.tensor_parallel_size > 1 for sc in stage_configs):
            # sharded stages own their GPUs: the reference's placement rules apply (src/config/model_config.py:136-150)
            from ..models.stage import validate_gpu_assignment
            validate_gpu_assignment([(sc.name, sc.tensor_parallel_size, list(sc.gpu_ids)) for sc in stage_configs])
        for i, sc in enumerate(stage_configs):
            self.stages.append(RealModelStage(sc, i, self.stages[-1] if (speculative and self.stages) else None,
                                              stage_kwargs))
        self.lambda_param = lambda_param
        self.quality_predictor = predictor or QualityPredictor({"feature_dim": 128})
        self.dp_solver: Optional[DynamicProgrammingSolver] = None
        self.is_initialized = False
        self.total_requests = 0
        self.total_inference_time = 0.0
        self.stage_usage = [0] * len(self.stages)
        self.logger = logging.getLogger("real_pipeline")

    def initialize(self):
        for s in self.stages:
            s.load_model()
        costs = [COST_PER_TOKEN.get(_size_from_name(s.config.name), 1.0) for s in self.stages]
        self.dp_solver = DynamicProgrammingSolver(len(self.stages), costs)
        self.is_initialized = True

    def _extract_features(self, prompt: str, result: StageInferenceResult) -> np.ndarray:
        f = [len(prompt), len(result.output_text), result.inference_time, result.stage_id]   # :445-460
        return np.array(f + [0.0] * (128 - len(f)), dtype=np.float32)

    def infer_adaptive(self, request: InferenceRequest) -> PipelineResult:
        if not self.is_initialized:
            raise RuntimeError("Pipeline not initialized. Call initialize() first.")
        t0 = time.perf_counter()
        self.total_requests += 1
        sp = SamplingParams(request.temperature, request.top_p, request.max_tokens)
        results: List[StageInferenceResult] = []
        selected = len(self.stages) - 1
        for sid, stage in enumerate(self.stages):
            res = stage.infer(request.prompt, sp)                       # every stage sees the original prompt (:395)
            results.append(res)
            if res.error:
                continue                                                # :399-401
            res.quality_score = self.quality_predictor.predict(self._extract_features(request.prompt, res))
            stop = self.dp_solver.should_stop(stage_id=sid, current_cost=sum(r.inference_time for r in results),
                                              quality_estimate=res.quality_score, lambda_param=request.lambda_param)
            res.should_continue = not stop
            if stop:
                selected = sid
                break
        good = [r for r in results if not r.error]
        final = results[selected] if not results[selected].error else (good[-1] if good else results[-1])
        self.stage_usage[final.stage_id] += 1
        infer_t = sum(r.inference_time for r in results)
        self.total_inference_time += infer_t
        return PipelineResult(request.request_id, request.prompt, final.output_text, final.stage_id, results, infer_t,
                              time.perf_counter() - t0, request.lambda_param, final.quality_score)

    def get_statistics(self) -> Dict[str, Any]:
        if self.total_requests == 0:
            return {"error": "No requests processed"}
        return {"total_requests": self.total_requests,
                "average_inference_time": self.total_inference_time / self.total_requests,
                "stage_usage": list(self.stage_usage)}

    def cleanup(self):
        for s in self.stages:
            s.unload_model()


def create_real_pipeline(config_path: str = "configs/qwen3_models.yaml", **kw) -> RealModelPipeline:
    """real_model_pipeline.py:522-526: stage list from the reference's YAML layout
    (configs/qwen3_models.yaml: ``models.stages[].{name, model_path, tensor_parallel_size, gpu_ids,
    max_model_len, dtype}``)."""
    import yaml
    with open(config_path) as f:
        cfg = yaml.safe_load(f)
    stages = cfg.get("models", cfg).get("stages", [])
    scs = [StageConfig(name=s.get("name", s.get("size_label", f"stage{i}")), model_path=s.get("model_path", ""),
                       tensor_parallel_size=int(s.get("tensor_parallel_size", 1)), gpu_ids=list(s.get("gpu_ids", [0])),
                       max_model_len=int(s.get("max_model_len", 4096)), dtype=s.get("dtype", "bfloat16"))
           for i, s in enumerate(stages)]
    pipe = RealModelPipeline(scs, **kw)
    pipe.initialize()
    return pipe
