"""Cascade orchestrator: the Python API surface of the reference's src/serving/pipeline.py
(``PipelineConfig`` :22-31, ``RequestResult`` :34-45, ``AdaptiveSpeculativePipeline`` :48-423) with the
B200 engine behind the stage seam.  Per stage: generate -> predict acceptance probability -> Bayesian
shrinkage -> DP stop rule -> stop or escalate (pipeline.py:184-266).

Reference quirks are decided explicitly, not inherited silently:
  * ``start_time`` is an undefined global in the reference (:269): here it is passed to ``_process_stages``.
  * The reference evaluates the DP on the PREFIX of probabilities seen so far (:248-256); with the
    last-index rule of dp_solver.py:69 that always stops at stage 0.  ``reference_compat=True``
    reproduces that loop exactly (checked against the unmodified reference in tests); the default
    evaluates the rule on the FULL stage vector, unseen stages taking probability 1.0 (the pipeline's
    own "last stage always accepts" convention, :241-242), which is the function the reference documents
    and precomputes (dp_solver.py:165-168, docs/guides/GETTING_STARTED.md:61-65).
  * Stage labels are configurable (the reference hard-codes "8b","13b","34b","70b", :175).
  * ``RequestResult`` also answers ``result["latency_ms"]`` etc. and the pipeline exposes ``lambda_value``
    because the reference's optimizer uses both (src/algorithms/optimizer.py:231,237-241)."""
from __future__ import annotations

import asyncio
import logging
import time
import uuid
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

from ..algorithms.dp_solver import bayesian_adjustment, optimal_stopping_rule
from ..models.predictor import FeatureExtractor, QualityPredictor
from ..models.stage import Stage, StageManager
from .cache_manager import KVCacheManager

logger = logging.getLogger(__name__)

REFERENCE_STAGE_NAMES = ["8b", "13b", "34b", "70b"]      # pipeline.py:175


@dataclass
class PipelineConfig:
    lambda_value: float = 1.0
    risk_adjustment: bool = True
    risk_alpha: float = 1.0
    risk_beta: float = 1.0
    enable_caching: bool = True
    max_concurrent_requests: int = 100
    batch_timeout_ms: float = 50.0
    # additions (not in the reference)
    stage_names: Optional[List[str]] = None
    reference_compat: bool = False
    # scorer -> stop decision on the GPU (asd_cascade_decide) whenever the stage returned device-resident features
    device_policy: bool = True


@dataclass
class RequestResult:
    request_id: str
    output: str
    stopped_at_stage: int
    latency_ms: float
    stage_probabilities: List[float]
    stage_costs: List[float]
    cache_hits: int
    total_tokens: int
    tokens_per_second: float

    def __getitem__(self, key):                      # optimizer.py:237-241 reads the result as a dict
        if key == "costs":
            return self.stage_costs
        return getattr(self, key)

    def get(self, key, default=None):
        try:
            return self[key]
        except AttributeError:
            return default


class AdaptiveSpeculativePipeline:
    def __init__(self, stage_manager: StageManager, predictor: QualityPredictor,
                 feature_extractor: FeatureExtractor, config: PipelineConfig,
                 cache_manager: Optional[KVCacheManager] = None):
        self.stage_manager, self.predictor, self.feature_extractor, self.config = (stage_manager, predictor,
                                                                                  feature_extractor, config)
        if cache_manager is None and config.enable_caching:
            cache_manager = KVCacheManager()
        self.cache_manager = cache_manager
        self.stats = self._fresh_stats()
        self.executor = ThreadPoolExecutor(max_workers=config.max_concurrent_requests)
        self.active_requests: Dict[str, Dict[str, Any]] = {}
        logger.info("AdaptiveSpeculativePipeline initialized")

    # -- lambda_value attribute used by optimizer.py:231
    @property
    def lambda_value(self) -> float:
        return self.config.lambda_value

    @lambda_value.setter
    def lambda_value(self, v: float):
        self.config.lambda_value = v

    def _stage_names(self) -> List[str]:
        if self.config.stage_names:
            return list(self.config.stage_names)
        if self.config.reference_compat:
            return list(REFERENCE_STAGE_NAMES)
        names = getattr(self.stage_manager, "stage_names", None)
        return list(names()) if callable(names) else list(REFERENCE_STAGE_NAMES)

    def _fresh_stats(self):
        n = max(4, len(self._stage_names()))
        return {"total_requests": 0, "stage_stops": [0] * n, "avg_latency": 0.0, "avg_tokens_per_second": 0.0,
                "total_tokens": 0, "avg_stage_probabilities": [0.0] * n, "error_count": 0}

    # -- pipeline.py:90-142
    def process_request(self, prompt: str, max_tokens: int = 512, temperature: float = 0.7,
                        request_id: Optional[str] = None) -> RequestResult:
        if request_id is None:
            request_id = str(uuid.uuid4())
        start_time = time.time()
        try:
            self.active_requests[request_id] = {"start_time": start_time,
                                                "prompt": prompt[:100] + "..." if len(prompt) > 100 else prompt}
            result = self._process_stages(request_id=request_id, prompt=prompt, max_tokens=max_tokens,
                                          temperature=temperature, start_time=start_time)
            self._update_stats(result)
            return result
        except Exception as e:
            logger.error(f"Request {request_id} failed: {e}")
            self.stats["error_count"] += 1
            raise
        finally:
            self.active_requests.pop(request_id, None)
            if self.cache_manager:
                self.cache_manager.cleanup_request(request_id)

    # -- pipeline.py:144-163
    async def process_request_async(self, prompt: str, max_tokens: int = 512, temperature: float = 0.7,
                                    request_id: Optional[str] = None) -> RequestResult:
        loop = asyncio.get_event_loop()
        return await loop.run_in_executor(self.executor, self.process_request, prompt, max_tokens, temperature,
                                          request_id)

    def _decide(self, probabilities: List[float], costs: List[float], stage_idx: int, n_stages: int,
                all_costs: Optional[List[float]]):
        """stop / escalate after stage `stage_idx`; returns (stop, k_star)."""
        if self.config.reference_compat:   # the reference's loop: DP on the prefix seen so far (:248-256)
            k_star, _ = optimal_stopping_rule(p=probabilities[:stage_idx + 1], C=costs[:stage_idx + 1],
                                              lam=self.config.lambda_value, risk_adjustment=False)
            return k_star == stage_idx, k_star
        # the documented rule: DP over ALL stages, unseen stages at probability 1.0
        full_p = probabilities + [1.0] * (n_stages - len(probabilities))
        k_full, _ = optimal_stopping_rule(p=full_p, C=all_costs, lam=self.config.lambda_value, risk_adjustment=False)
        return k_full <= stage_idx, stage_idx

    def _device_decide(self, stage, prompts, outputs, logprobs, prev_probs, costs, all_costs, stage_idx, n_stages, n_obs):
        """(prob, stop, k_star) arrays from ``Stage.decide`` (one ``asd_cascade_decide`` launch for all requests), or
        None when the stage is not a B200 stage / returned no device features / the predictor is not the MLP."""
        if not self.config.device_policy or not hasattr(stage, "decide"):
            return None
        if any(getattr(lp, "fused_device", None) is None for lp in logprobs):
            return None
        mlp = getattr(self.predictor, "mlp", None)
        if mlp is None or len(mlp) < 4 or type(self.feature_extractor).__name__ != "FeatureExtractor":
            return None
        compat = self.config.reference_compat
        C = list(costs[:stage_idx + 1]) if compat else list(all_costs)
        if compat:
            C = C + [0.0] * (n_stages - len(C))       # only the first stage_idx + 1 entries are read in prefix mode
        return stage.decide(self.predictor, prompts, outputs, logprobs, prev_probs, C, stage_idx,
                            self.config.lambda_value, prefix_mode=compat, risk_adjustment=self.config.risk_adjustment,
                            n_obs=n_obs, alpha=self.config.risk_alpha, beta=self.config.risk_beta)

    # -- pipeline.py:165-286
    def _process_stages(self, request_id: str, prompt: str, max_tokens: int, temperature: float,
                        start_time: float) -> RequestResult:
        stage_names = self._stage_names()
        compat = self.config.reference_compat
        all_costs = None if compat else [self.stage_manager.get_stage(n).cost_per_token for n in stage_names]
        current_prompt = prompt
        probabilities: List[float] = []
        costs: List[float] = []
        outputs: List[str] = []
        cache_hits = total_tokens = 0
        k_star = 0
        for stage_idx, stage_name in enumerate(stage_names):
            stage = self.stage_manager.get_stage(stage_name)
            cached_output = None
            if self.cache_manager:
                cached = self.cache_manager.get_cache(request_id, stage_idx)
                if cached:
                    cached_output = cached.get("output")
                    cache_hits += 1
            if cached_output:
                stage_output, stage_logprobs, stage_stats = [cached_output], [np.array([])], {"generation_time_ms": 0}
            else:
                stage_output, stage_logprobs, stage_stats = stage.generate(
                    prompts=[current_prompt], max_tokens=max_tokens, temperature=temperature, return_logprobs=True)
                if self.cache_manager:
                    self.cache_manager.allocate(request_id, stage_idx, {
                        "output": stage_output[0],
                        "logprobs": stage_logprobs[0] if len(stage_logprobs) else np.array([])})
            outputs.append(stage_output[0])
            costs.append(stage.cost_per_token)
            total_tokens += len(stage_output[0].split())
            dev = self._device_decide(stage, [current_prompt], [stage_output[0]],
                                      [stage_logprobs[0]] if len(stage_logprobs) else [None], [probabilities], costs,
                                      all_costs, stage_idx, len(stage_names), max(100, self.stats["total_requests"]))
            if dev is not None:      # features -> MLP -> shrinkage -> DP in one launch on the GPU
                probabilities.append(float(dev[0][0]))
                stop, k_star = bool(dev[1][0]), int(dev[2][0])
            else:
                if stage_idx < len(stage_names) - 1:
                    prob = self.predictor.predict(prompt=current_prompt, draft_output=stage_output[0],
                                                  draft_logprobs=stage_logprobs[0] if len(stage_logprobs) else None,
                                                  stage_id=stage_idx, feature_extractor=self.feature_extractor)
                    if self.config.risk_adjustment:
                        n_obs = max(100, self.stats["total_requests"])                       # :235
                        prob = bayesian_adjustment(prob, n_obs, self.config.risk_alpha, self.config.risk_beta)
                    probabilities.append(prob)
                else:
                    probabilities.append(1.0)                                                # :241-242
                stop, k_star = self._decide(probabilities, costs, stage_idx, len(stage_names), all_costs)
            if stop:
                break
            if stage_idx < len(stage_names) - 1:
                current_prompt = prompt + " " + stage_output[0]                          # :266
        total_time_ms = (time.time() - start_time) * 1000
        tokens_per_second = total_tokens / (total_time_ms / 1000) if total_time_ms > 0 else 0
        if self.cache_manager:
            self.cache_manager.truncate_at_stage(request_id, k_star)
        return RequestResult(request_id=request_id, output=outputs[k_star], stopped_at_stage=k_star,
                             latency_ms=total_time_ms, stage_probabilities=probabilities,
                             stage_costs=costs[:k_star + 1], cache_hits=cache_hits, total_tokens=total_tokens,
                             tokens_per_second=tokens_per_second)

    # -- pipeline.py:288-312
    def _update_stats(self, result: RequestResult):
        st = self.stats
        st["total_requests"] += 1
        st["stage_stops"][result.stopped_at_stage] += 1
        st["total_tokens"] += result.total_tokens
        alpha = 0.01
        st["avg_latency"] = (1 - alpha) * st["avg_latency"] + alpha * result.latency_ms
        st["avg_tokens_per_second"] = (1 - alpha) * st["avg_tokens_per_second"] + alpha * result.tokens_per_second
        for i, prob in enumerate(result.stage_probabilities):
            if i < len(st["avg_stage_probabilities"]):
                st["avg_stage_probabilities"][i] = (1 - alpha) * st["avg_stage_probabilities"][i] + alpha * prob

    # -- pipeline.py:314-338 (sequential, as in the reference) + the batching its TODO at :331 asks for
    def batch_process(self, prompts: List[str], max_tokens: int = 512, temperature: float = 0.7,
                      batched: bool = False) -> List[RequestResult]:
        """`batched=False`: one request after the other, exactly like the reference.  `batched=True`: the requests
        walk the cascade together - every stage makes ONE `generate` call for all requests still alive, which is
        what fills the engine's batch dimension; each request keeps its own probabilities, costs and stop decision
        (same rule as `_process_stages`).  Differences to the sequential loop: no per-(request, stage) blob cache,
        and the Bayesian n_obs is the request count before the batch for all of its members."""
        if not batched:
            return [self.process_request(p, max_tokens, temperature) for p in prompts]
        start = time.time()
        names = self._stage_names()
        all_costs = None if self.config.reference_compat else [self.stage_manager.get_stage(n).cost_per_token for n in names]
        n_obs = max(100, self.stats["total_requests"])
        st = [dict(rid=str(uuid.uuid4()), prompt=p, cur=p, probs=[], costs=[], outs=[], tokens=0, k=None)
              for p in prompts]
        for s_ in st:
            self.active_requests[s_["rid"]] = {"start_time": start, "prompt": s_["prompt"][:100]}
        try:
            live = list(range(len(st)))
            for stage_idx, name in enumerate(names):
                if not live:
                    break
                stage = self.stage_manager.get_stage(name)
                texts, logprobs, _ = stage.generate(prompts=[st[i]["cur"] for i in live], max_tokens=max_tokens,
                                                    temperature=temperature, return_logprobs=True)
                nxt = []
                dev = self._device_decide(stage, [st[i]["cur"] for i in live], texts,
                                          [logprobs[j] if len(logprobs) > j else None for j in range(len(live))],
                                          [st[i]["probs"] for i in live],
                                          st[live[0]]["costs"] + [stage.cost_per_token], all_costs, stage_idx,
                                          len(names), n_obs)
                for j, i in enumerate(live):
                    r = st[i]
                    r["outs"].append(texts[j])
                    r["costs"].append(stage.cost_per_token)
                    r["tokens"] += len(texts[j].split())
                    if dev is not None:
                        r["probs"].append(float(dev[0][j]))
                        stop, k_star = bool(dev[1][j]), int(dev[2][j])
                        if stop or stage_idx == len(names) - 1:
                            r["k"] = k_star
                        else:
                            r["cur"] = r["prompt"] + " " + texts[j]
                            nxt.append(i)
                        continue
                    if stage_idx < len(names) - 1:
                        lp = logprobs[j] if len(logprobs) > j else None
                        prob = self.predictor.predict(prompt=r["cur"], draft_output=texts[j], draft_logprobs=lp,
                                                      stage_id=stage_idx, feature_extractor=self.feature_extractor)
                        if self.config.risk_adjustment:
                            prob = bayesian_adjustment(prob, n_obs, self.config.risk_alpha, self.config.risk_beta)
                        r["probs"].append(prob)
                    else:
                        r["probs"].append(1.0)
                    stop, k_star = self._decide(r["probs"], r["costs"], stage_idx, len(names), all_costs)
                    if stop or stage_idx == len(names) - 1:
                        r["k"] = k_star
                    else:
                        r["cur"] = r["prompt"] + " " + texts[j]
                        nxt.append(i)
                live = nxt
            results = []
            for r in st:
                ms = (time.time() - start) * 1000
                res = RequestResult(request_id=r["rid"], output=r["outs"][r["k"]], stopped_at_stage=r["k"], latency_ms=ms,
                                    stage_probabilities=r["probs"], stage_costs=r["costs"][:r["k"] + 1], cache_hits=0,
                                    total_tokens=r["tokens"], tokens_per_second=r["tokens"] / (ms / 1000) if ms > 0 else 0)
                self._update_stats(res)
                results.append(res)
            return results
        except Exception as e:
            logger.error(f"Batch of {len(prompts)} requests failed: {e}")
            self.stats["error_count"] += 1
            raise
        finally:
            for s_ in st:
                self.active_requests.pop(s_["rid"], None)

    def update_lambda(self, new_lambda: float):
        old = self.config.lambda_value
        self.config.lambda_value = new_lambda
        logger.info(f"Updated lambda: {old:.3f} -> {new_lambda:.3f}")

    def get_stats(self) -> Dict[str, Any]:
        stats = dict(self.stats)
        n = len(stats["stage_stops"])
        if stats["total_requests"] > 0:
            stats["stage_distribution"] = [c / stats["total_requests"] for c in stats["stage_stops"]]
            stats["avg_tokens_per_request"] = stats["total_tokens"] / stats["total_requests"]
        else:
            stats["stage_distribution"] = [0.0] * n
            stats["avg_tokens_per_request"] = 0.0
        if self.cache_manager:
            stats["cache_stats"] = self.cache_manager.get_stats()
        stats["active_requests"] = len(self.active_requests)
        return stats

    def reset_stats(self):
        self.stats = self._fresh_stats()
        logger.info("Pipeline statistics reset")

    def warmup(self, num_requests: int = 5):
        prompts = ["Hello, how are you today?", "What is the capital of France?",
                   "Explain machine learning in simple terms.", "Write a short poem about nature.",
                   "What are the benefits of renewable energy?"]
        for i in range(num_requests):
            try:
                self.process_request(prompt=prompts[i % len(prompts)], max_tokens=50, temperature=0.7)
            except Exception as e:
                logger.warning(f"Warmup request {i + 1} failed: {e}")

    def shutdown(self):
        self.executor.shutdown(wait=True)
        if self.cache_manager and hasattr(self.cache_manager, "shutdown"):
            self.cache_manager.shutdown()


AdaptiveSpeculativeDecodingPipeline = AdaptiveSpeculativePipeline   # name used by experiments/scripts/run_comprehensive_evaluation.py:21
