"""ORACLE-side CPU baseline for BASELINE configs[0]: Qwen2.5-0.5B draft -> Qwen2.5-1.5B target, random-init,
chain k = 4, batch 1, greedy, on the host cores (``bench.py --impl reference`` only).

The reference's Stage is ``vllm.LLM`` / HF ``generate`` (src/serving/real_model_pipeline.py:98-108,
src/training/generate_training_data.py:110-119): the CPU path it can actually run is HF ``Qwen2ForCausalLM``
in fp32, so that is what is timed - the installed ``transformers`` model with its own KV cache, driven by a plain
draft-then-verify loop (k draft forwards of 1 token, one (k+1)-token verify forward, greedy accept rule of
SURVEY App. C).  A bounded number of steps; nothing here is extrapolated."""
from __future__ import annotations

import os
import time

import torch


def _hf_model(cfg, seed):
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(seed)
    hc = Qwen2Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                     num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                     num_key_value_heads=cfg.num_key_value_heads, max_position_embeddings=4096,
                     rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta, tie_word_embeddings=cfg.tie_word_embeddings)
    m = Qwen2ForCausalLM(hc).eval()
    return m


@torch.no_grad()
def run(draft_cfg, target_cfg, k=4, prompt_len=64, steps=8, threads=None, seed=1234):
    from transformers import DynamicCache
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    t_build = time.perf_counter()
    d, t = _hf_model(draft_cfg, 0), _hf_model(target_cfg, 1)
    t_build = time.perf_counter() - t_build
    ids = torch.randint(0, target_cfg.vocab_size, (1, prompt_len), generator=torch.Generator().manual_seed(seed))
    dc, tc = DynamicCache(), DynamicCache()
    last = t(ids, past_key_values=tc, use_cache=True).logits[:, -1].argmax(-1)
    d(ids, past_key_values=dc, use_cache=True)
    pos, emitted = prompt_len, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        toks, x = [last], last
        for _i in range(k):
            x = d(x[:, None], past_key_values=dc, use_cache=True).logits[:, -1].argmax(-1)
            toks.append(x)
        block = torch.stack(toks, 1)                                   # [1, k+1]
        tl = t(block, past_key_values=tc, use_cache=True).logits[0]    # [k+1, V]
        am = tl.argmax(-1)
        n = 0
        while n < k and int(block[0, n + 1]) == int(am[n]):
            n += 1
        last = am[n:n + 1]
        emitted += n + 1
        pos += n + 1
        tc.crop(pos)                                                   # roll the rejected tail back
        if n == k:                                                     # draft cache lacks the last accepted token
            d(block[:, k:k + 1], past_key_values=dc, use_cache=True)
        else:
            dc.crop(pos)
    dt = time.perf_counter() - t0
    return {"value": emitted / dt, "unit": "tok/s", "steps": steps, "seconds": dt, "tokens": emitted, "cores": threads,
            "model_build_seconds": round(t_build, 1),
            "what": f"HF Qwen2ForCausalLM fp32 on CPU, {draft_cfg.name} -> {target_cfg.name}, chain k={k}, batch 1, "
                    f"greedy, {prompt_len}-token prompt"}
