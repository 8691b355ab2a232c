"""Engine seam.  ``Stage`` / ``StageManager`` / ``StageConfig`` are imported by the reference
(src/serving/pipeline.py:14, src/serving/server.py:23,151-164) from ``src/models/stage.py``, a file the
reference does not ship; the signatures below are reconstructed from those call sites and from the
listing in docs/guides/RESEARCH_PROTOCOL.md:233-304.  Where the reference's
Stage wraps ``vllm.LLM`` this one wraps the B200 engine: a ``QwenEngine`` target, optionally
accelerated by a smaller draft engine through chain draft-then-verify (``SpecDecoder``)."""
from __future__ import annotations

import logging
import threading
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from .._lib import AsdError
from .qwen2 import QWEN25, SIZE_ALIASES, Qwen2Config, get_config

logger = logging.getLogger(__name__)


class ModelLoadError(AsdError):
    """name used by the reference at the stage boundary (real_model_pipeline.py:92,115)"""


class InferenceError(AsdError):
    """name used by the reference at the stage boundary (real_model_pipeline.py:120,166)"""


COST_PER_TOKEN = {"8b": 1.0, "13b": 1.6, "34b": 4.2, "70b": 8.8,       # RESEARCH_PROTOCOL.md:255-260
                  "7b": 1.0, "14b": 2.0, "32b": 4.5, "72b": 10.0,      # docs/papers/FINAL_PAPER.md:130-132
                  "0.5b": 0.1, "1.5b": 0.3}
KV_GB_PER_TOKEN = {"8b": 0.002, "13b": 0.003, "34b": 0.008, "70b": 0.016}   # RESEARCH_PROTOCOL.md:296-301
PAD_LOGPROB = -1.0e4   # finite stand-in for "not in the top-k": exp(PAD) * PAD == -0.0, never NaN


class FusedLogprobs(np.ndarray):
    """ndarray [T, 5] shaped like vLLM's top-5 logprobs (col 0: emitted token, col 1: arg-max token,
    col 2: second best, cols 3-4 padding) that also carries ``fused``: float32 [T, 6] =
    (lse, p_max, margin, entropy, ln p(draft tok), ln p(emitted tok)) computed by the sampling kernel over
    the FULL vocabulary (asd_b200.ops.FEATURE_NAMES)."""
    fused: Optional[np.ndarray] = None
    fused_device = None     # the same [T, 6] rows as a CUDA tensor (input of asd_cascade_decide; never copied back)

    def __array_finalize__(self, obj):
        self.fused = getattr(obj, "fused", None)
        self.fused_device = getattr(obj, "fused_device", None)


def make_logprobs(emitted_lp: np.ndarray, fused: np.ndarray, fused_device=None) -> FusedLogprobs:
    T = len(emitted_lp)
    a = np.full((T, 5), PAD_LOGPROB, dtype=np.float64)
    if T:
        pmax = np.clip(fused[:, 1].astype(np.float64), 1e-300, 1.0)
        p2 = np.clip(pmax - fused[:, 2].astype(np.float64), 1e-300, 1.0)
        a[:, 0], a[:, 1], a[:, 2] = emitted_lp, np.log(pmax), np.log(p2)
    out = a.view(FusedLogprobs)
    out.fused = np.asarray(fused, dtype=np.float32)
    out.fused_device = fused_device
    return out


class ByteTokenizer:
    """Offline stand-in for the HF tokenizer (no tokenizer files can be fetched here): UTF-8 bytes."""

    def __init__(self, vocab_size: int):
        self.vocab_size = vocab_size

    def encode(self, text: str) -> List[int]:
        ids = list(text.encode("utf-8")) or [0]
        return [i % self.vocab_size for i in ids]

    def decode(self, ids: List[int]) -> str:
        return bytes(int(i) & 0xFF for i in ids).decode("utf-8", errors="replace")


def load_tokenizer(model_name: str, vocab_size: int):
    import os
    if os.path.isdir(model_name):
        try:
            from transformers import AutoTokenizer
            tok = AutoTokenizer.from_pretrained(model_name, local_files_only=True)
            tok.encode_ids = lambda t: tok.encode(t, add_special_tokens=False)
            return tok
        except Exception as e:  # pragma: no cover - depends on local files
            logger.warning("tokenizer at %s unusable (%s); falling back to bytes", model_name, e)
    return ByteTokenizer(vocab_size)


@dataclass
class StageConfig:
    """server.py:151-158 (+ ``gpu_ids`` / ``max_model_len`` of configs/qwen3_models.yaml:10-51)"""
    model_name: str
    model_size: str
    tensor_parallel_size: int = 1
    gpu_memory_utilization: float = 0.8
    quantized: bool = False
    cost_per_token: Optional[float] = None
    gpu_ids: Optional[List[int]] = None
    max_model_len: Optional[int] = None
    config: Optional[Qwen2Config] = None      # architecture override (e.g. a layer-truncated shape for tests)


def validate_gpu_assignment(stages) -> None:
    """The reference's placement rules (src/config/model_config.py:136-150): a stage needs exactly
    ``tensor_parallel_size`` GPUs and no GPU may serve two stages.  ``stages``: iterable of
    (label, tensor_parallel_size, gpu_ids)."""
    seen = {}
    for label, tp, gpus in stages:
        if len(gpus) != tp:
            raise ModelLoadError(f"Stage {label}: GPU count ({len(gpus)}) doesn't match tensor_parallel_size ({tp})")
        if len(set(gpus)) != len(gpus):
            raise ModelLoadError(f"Stage {label}: duplicate GPU ids {gpus}")
        for g in gpus:
            if g in seen:
                raise ModelLoadError(f"GPU IDs cannot be assigned to multiple stages (GPU {g}: {seen[g]} and {label})")
            seen[g] = label


_ORDER = threading.Lock()
_NEXT_ORDER = [0]
# one lock per GPU: generations that touch the same device run one after the other.  Kernels of different streams on
# one SM are not covered by the forward's TMEM hand-over invariant: a 512-column tcgen05.alloc
# of one stream can hold the allocation permit while a clustered GEMM of the other waits for a sibling that needs it.
_DEVICE_LOCKS = {}


def _device_locks(gpu_ids):
    with _ORDER:
        return [_DEVICE_LOCKS.setdefault(int(g), threading.RLock()) for g in sorted(set(int(g) for g in gpu_ids))]


class Stage:
    """One model of the cascade.  ``generate`` keeps the reference's contract:
    ``(texts, logprobs, stats)`` with ``stats["generation_time_ms"]`` (pipeline.py:204-209,245).

    ``tensor_parallel_size = t`` shards the model Megatron-style over ``gpu_ids`` (t GPUs of this process,
    exchanging partial sums over NVLink peer memory) exactly where the reference passes ``tensor_parallel_size``
    to ``vllm.LLM`` (RESEARCH_PROTOCOL.md:262-268, real_model_pipeline.py:98-108).  ``draft`` is the previous
    (smaller) stage whose engine proposes k tokens per step; it may sit on another GPU."""

    def __init__(self, model_name: str, model_size: str, tensor_parallel_size: int = 1,
                 gpu_memory_utilization: float = 0.8, quantized: bool = False, *,
                 cost_per_token: Optional[float] = None, config: Optional[Qwen2Config] = None,
                 draft: Optional["Stage"] = None, k: int = 5, weights: Optional[dict] = None, seed: int = 0,
                 max_batch: int = 16, max_model_len: int = 4096, device=None, gpu_ids: Optional[List[int]] = None,
                 allow_random_weights: bool = True):
        self.model_name, self.model_size = model_name, model_size.lower()
        self.tensor_parallel_size = int(tensor_parallel_size)
        self.gpu_memory_utilization, self.quantized = gpu_memory_utilization, quantized
        if quantized:
            logger.warning("quantized=True is ignored: the B200 engine runs bf16 weights")
        if self.tensor_parallel_size < 1:
            raise ModelLoadError("tensor_parallel_size must be >= 1")
        if gpu_ids is None:
            if device is not None:
                import torch
                d = torch.device(device)
                first = d.index if d.index is not None else 0
            else:
                first = 0
            gpu_ids = list(range(first, first + self.tensor_parallel_size))
        self.gpu_ids = [int(g) for g in gpu_ids]
        validate_gpu_assignment([(self.model_size, self.tensor_parallel_size, self.gpu_ids)])
        try:
            self.cfg = config or get_config(self.model_size)
        except KeyError as e:
            raise ModelLoadError(f"unknown model size {model_size!r}") from e
        if self.cfg.num_key_value_heads % self.tensor_parallel_size or self.cfg.intermediate_size % self.tensor_parallel_size:
            raise ModelLoadError(f"{self.model_size}: {self.cfg.num_key_value_heads} KV heads / ffn "
                                 f"{self.cfg.intermediate_size} do not split over {self.tensor_parallel_size} ranks")
        self.cost_per_token = cost_per_token if cost_per_token is not None else COST_PER_TOKEN.get(self.model_size, 1.0)
        self.draft, self.k = draft, k
        self.max_batch, self.max_model_len = max_batch, max_model_len
        self.tokenizer = load_tokenizer(model_name, self.cfg.vocab_size)
        self._lock = threading.RLock()       # the pipeline calls generate() from up to 100 threads
        with _ORDER:
            self._order = _NEXT_ORDER[0]
            _NEXT_ORDER[0] += 1
        try:
            import torch
            if not torch.cuda.is_available():
                raise ModelLoadError("Stage needs CUDA devices (the engine has no CPU fallback)")
            if max(self.gpu_ids) >= torch.cuda.device_count():
                raise ModelLoadError(f"Stage {self.model_size}: gpu_ids {self.gpu_ids} but only "
                                     f"{torch.cuda.device_count()} CUDA devices are visible")
            from ..engine import QwenEngine, TPQwenEngine
            mt = max(256, max_batch * (k + 1))
            if self.tensor_parallel_size == 1:
                self.engine = QwenEngine(self.cfg, max_seqs=max_batch, max_seq_len=max_model_len, max_tokens=mt,
                                         device=f"cuda:{self.gpu_ids[0]}")
            else:
                self.engine = TPQwenEngine(self.cfg, self.gpu_ids, max_seqs=max_batch, max_seq_len=max_model_len,
                                           max_tokens=mt)
            self.weights_source = self._load_weights(weights, seed, allow_random_weights)
        except AsdError:
            raise
        except Exception as e:
            raise ModelLoadError(f"failed to load {model_name}: {e}") from e

    def _load_weights(self, weights, seed, allow_random):
        import os
        if weights is not None:
            self.engine.load_hf_weights(weights)
            return "state_dict"
        if os.path.isdir(self.model_name):
            from .qwen2 import load_safetensors_dir
            try:
                w = load_safetensors_dir(self.model_name)
            except FileNotFoundError:
                w = None
            if w is not None:
                self.engine.load_hf_weights(w)
                return "safetensors:" + self.model_name
        if not allow_random:
            raise ModelLoadError(f"no checkpoint found at {self.model_name!r} and random weights are not allowed")
        logger.warning("Stage %s: no checkpoint at %r - running RANDOM-INIT weights of the %s shape (seed %d); "
                       "outputs are meaningless text, timings are real", self.model_size, self.model_name,
                       self.cfg.name, seed)
        self.engine.load_random(seed)
        return f"random(seed={seed})"

    # ------------------------------------------------------------------ reference API
    def generate(self, prompts: List[str], max_tokens: int = 512, temperature: float = 0.7, top_p: float = 0.9,
                 return_logprobs: bool = True) -> Tuple[List[str], List[np.ndarray], Dict[str, float]]:
        """Prompts of any lengths are decoded together (ragged prefill, one batch of up to ``max_batch``).
        ``top_p`` is accepted for signature compatibility (RESEARCH_PROTOCOL.md:272-279); sampling is over
        the full softmax(z / T) - values below 1.0 are logged once and not applied."""
        if top_p is not None and top_p < 1.0 and not getattr(self, "_top_p_warned", False):
            self._top_p_warned = True
            logger.info("top_p=%.2f is not applied: the fused sampler draws from the full softmax(z/T)", top_p)
        t0 = time.time()
        ids = [self._encode(p) for p in prompts]
        texts: List[Optional[str]] = [None] * len(prompts)
        lps: List[Optional[np.ndarray]] = [None] * len(prompts)
        acc_tok = steps = 0
        # a stage's engine also serves as the NEXT stage's draft: whole generations are serialised per engine,
        # locks taken in creation order (smaller stage first) so two stages can never deadlock
        chain = sorted([st for st in (self.draft, self) if st is not None], key=lambda st: st._order)
        devs = _device_locks([g for st in chain for g in st.gpu_ids])      # device locks first (by index), then stages
        for d in devs:
            d.acquire()
        for st in chain:
            st._lock.acquire()
        try:
            order = sorted(range(len(ids)), key=lambda i: len(ids[i]))      # similar lengths share a batch
            for s in range(0, len(order), self.max_batch):
                grp = order[s:s + self.max_batch]
                toks, lp, fused, st = self._generate_ids([ids[i] for i in grp], max_tokens, temperature)
                acc_tok += st["accepted"]
                steps += st["steps"]
                fdev = st.get("features_device")
                for j, i in enumerate(grp):
                    texts[i] = self.tokenizer.decode(toks[j])
                    lps[i] = (make_logprobs(lp[j], fused[j], None if fdev is None else fdev[j])
                              if return_logprobs else np.array([]))
        finally:
            for st in reversed(chain):
                st._lock.release()
            for d in reversed(devs):
                d.release()
        stats = {"generation_time_ms": (time.time() - t0) * 1000.0, "draft_tokens_accepted": acc_tok,
                 "decode_steps": steps}
        return texts, lps, stats

    def decide(self, predictor, prompts: List[str], outputs: List[str], logprobs, prev_probs, costs, stage_idx: int,
               lam: float, prefix_mode: bool = False, risk_adjustment: bool = False, n_obs: float = 100.0,
               alpha: float = 1.0, beta: float = 1.0):
        """Scorer -> stop decision for requests that just ran this stage, on the device (``asd_cascade_decide``):
        the per-token features of ``logprobs[i].fused_device`` never leave the GPU; returns
        (prob [n], stop [n], k_star [n]) as numpy.  Same chain as pipeline.py:225-256 of the reference."""
        import torch
        from ..ops import cascade_decide
        rows = [lp.fused_device for lp in logprobs]
        T = max(int(r.shape[0]) for r in rows)
        dev = rows[0].device
        if len(rows) == 1 and rows[0].is_contiguous():
            feats = rows[0][None]
        else:
            feats = torch.zeros(len(rows), T, rows[0].shape[1], dtype=torch.float32, device=dev)
            for i, r in enumerate(rows):
                feats[i, :r.shape[0]] = r
        ntok = torch.tensor([int(r.shape[0]) for r in rows], dtype=torch.int32).to(dev, non_blocking=True)
        scalars = [[len(p.split()) / 2048, len(o.split()) / 512, stage_idx / 4.0] for p, o in zip(prompts, outputs)]
        return cascade_decide(feats, ntok, scalars, predictor, prev_probs, costs, stage_idx, lam, prefix_mode,
                              risk_adjustment, n_obs, alpha, beta)

    def get_model_info(self) -> Dict[str, object]:
        return {"model_name": self.model_name, "model_size": self.model_size, "parameters": self.cfg.params,
                "tensor_parallel_size": self.tensor_parallel_size, "gpu_ids": list(self.gpu_ids),
                "cost_per_token": self.cost_per_token, "weights": self.weights_source,
                "dtype": "bfloat16", "hidden_size": self.cfg.hidden_size, "num_layers": self.cfg.num_hidden_layers,
                "draft": None if self.draft is None else self.draft.model_size, "engine": "asd_b200 (sm_100a)"}

    def compute_kv_cache_size(self, seq_len: int) -> float:
        """GB of KV cache for one sequence (exact for the Qwen2.5 shape; the reference's listing uses a
        per-size constant, RESEARCH_PROTOCOL.md:292-303)."""
        if self.model_size in KV_GB_PER_TOKEN and self.model_size not in QWEN25:
            return seq_len * KV_GB_PER_TOKEN[self.model_size]
        return seq_len * self.cfg.kv_bytes_per_token() / 1e9

    # ------------------------------------------------------------------ internals
    def _encode(self, text: str) -> List[int]:
        enc = getattr(self.tokenizer, "encode_ids", None) or self.tokenizer.encode
        ids = list(enc(text))
        return ids[-(self.max_model_len // 2):] or [0]

    def _draft_engine(self):
        d = self.draft
        if d is None:
            return None
        if d.cfg.vocab_size != self.cfg.vocab_size:
            if not getattr(self, "_vocab_warned", False):
                self._vocab_warned = True
                logger.warning("stage %s cannot draft for %s (vocabularies differ): plain decoding", d.model_size,
                               self.model_size)
            return None
        return d.engine

    def _generate_ids(self, prompts: List[List[int]], max_tokens: int, temperature: float):
        from ..engine import SpecDecoder
        B, P = len(prompts), max(len(p) for p in prompts)
        draft = self._draft_engine()
        max_len = self.max_model_len if draft is None else min(self.max_model_len, self.draft.max_model_len)
        max_new = max(1, min(max_tokens, max_len - P - self.k - 2))
        try:
            dec = SpecDecoder(self.engine, draft, B, self.k, temperature)
            toks, lp, fused, st = dec.generate(prompts, max_new)
        except InferenceError:
            raise
        except AsdError as e:
            raise InferenceError(str(e)) from e
        toks, lp, fused = toks.numpy(), lp.numpy().astype(np.float64), fused.numpy()
        return ([[int(x) for x in toks[b]] for b in range(B)], [lp[b] for b in range(B)],
                [fused[b] for b in range(B)], st)


class StageManager:
    """server.py:163-164, pipeline.py:185: ``StageManager(stage_configs, gpu_allocation).get_stage(name)``.
    Stages are keyed by size label; the reference's Llama-era labels (8b/13b/34b/70b, pipeline.py:175)
    resolve to the Qwen2.5 cascade it describes (7b/14b/32b/72b).  Each stage drafts with the previous
    (smaller) stage when ``speculative=True``."""

    def __init__(self, stage_configs: List[StageConfig], gpu_allocation: Optional[Dict[str, List[int]]] = None, *,
                 speculative: bool = True, k: int = 5, stage_kwargs: Optional[dict] = None):
        self.gpu_allocation = {str(k_).lower(): list(v) for k_, v in (gpu_allocation or {}).items()}
        self.stages: Dict[str, Stage] = {}
        self.order: List[str] = []
        prev: Optional[Stage] = None
        placement, nxt = [], 0
        for sc in stage_configs:
            gpus = self.gpu_allocation.get(sc.model_size.lower(), sc.gpu_ids)
            if gpus is None:
                # no explicit placement: single-GPU stages share GPU 0 (one box, one GPU), sharded stages take
                # the next free block
                gpus = [0] if sc.tensor_parallel_size == 1 else list(range(nxt, nxt + sc.tensor_parallel_size))
            nxt = max(nxt, max(gpus) + 1)
            placement.append(list(gpus))
        explicit = [(sc.model_size, sc.tensor_parallel_size, g) for sc, g in zip(stage_configs, placement)
                    if self.gpu_allocation.get(sc.model_size.lower(), sc.gpu_ids) is not None]
        validate_gpu_assignment(explicit)            # model_config.py:136-150
        for i, (sc, gpus) in enumerate(zip(stage_configs, placement)):
            kw = dict(stage_kwargs or {})
            if sc.max_model_len is not None:
                kw.setdefault("max_model_len", sc.max_model_len)
            if sc.config is not None:
                kw["config"] = sc.config
            st = Stage(sc.model_name, sc.model_size, sc.tensor_parallel_size, sc.gpu_memory_utilization, sc.quantized,
                       cost_per_token=sc.cost_per_token, draft=prev if speculative else None, k=k, seed=i,
                       gpu_ids=gpus, **kw)
            key = sc.model_size.lower()
            self.stages[key] = st
            self.order.append(key)
            prev = st

    def get_stage(self, name: str) -> Stage:
        key = name.lower()
        if key in self.stages:
            return self.stages[key]
        alias = SIZE_ALIASES.get(key)
        if alias in self.stages:
            return self.stages[alias]
        rev = {v: k for k, v in SIZE_ALIASES.items()}
        if rev.get(key) in self.stages:
            return self.stages[rev[key]]
        raise KeyError(f"no stage named {name!r}; have {list(self.stages)}")

    def stage_names(self) -> List[str]:
        return list(self.order)

    def calibrate_costs(self, prompts: Optional[List[str]] = None, max_tokens: int = 32) -> Dict[str, float]:
        """SURVEY.md 8 f3 (RealModelPipeline._calibrate_costs, real_model_pipeline.py:313-362): replace the
        hard-coded cost table by measured per-token generation time on this machine, normalised so that the
        first stage costs 1.0; updates ``stage.cost_per_token`` in place and returns the new table."""
        prompts = prompts or ["Hello, how are you today?", "What is the capital of France?",
                              "Explain machine learning in simple terms."]
        per_tok = {}
        for key in self.order:
            st = self.stages[key]
            st.generate(prompts[:1], max_tokens=4, temperature=0.0)          # warm-up
            _, _, stats = st.generate(prompts, max_tokens=max_tokens, temperature=0.0)
            per_tok[key] = stats["generation_time_ms"] / (len(prompts) * max_tokens)
        base = per_tok[self.order[0]]
        for key in self.order:
            self.stages[key].cost_per_token = per_tok[key] / base
        return {k: self.stages[k].cost_per_token for k in self.order}

    def warmup_all(self):
        for st in self.stages.values():
            st.generate(["warmup"], max_tokens=4, temperature=0.0)
