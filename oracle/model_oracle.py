"""ORACLE (test infrastructure only): CPU fp32 restatement of the Qwen2 decoder forward.

The reference never implements the model forward: Stage.generate delegates to vLLM
(/root/reference/src/serving/real_model_pipeline.py:98-108,135; docs/guides/RESEARCH_PROTOCOL.md:233-304)
and data generation to HF ``model.generate`` (src/training/generate_training_data.py:110-119).
The arithmetic therefore lives in third-party code absent from /root/reference:
``transformers>=4.40,<5`` (requirements.txt:5-6; docs pin transformers==4.40.0), class
``Qwen2ForCausalLM``.  This file restates its published algorithm (RMSNorm with eps inside rsqrt,
rotate_half RoPE with theta = 1e6, GQA attention with QKV bias and 1/sqrt(head_dim) scale, SwiGLU MLP)
in plain torch fp32 loops over layers, with full causal attention and no KV cache.  It is pinned by
tests/golden/qwen2_tiny_golden.npz, produced by oracle/gen_model_golden.py from the installed
``transformers`` Qwen2ForCausalLM (version recorded in the fixture).
"""
from __future__ import annotations

import math

import torch


def rms_norm(x, w, eps):
    v = x.float().pow(2).mean(-1, keepdim=True)
    return x.float() * torch.rsqrt(v + eps) * w.float()


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def inv_freq(head_dim: int, theta: float) -> torch.Tensor:
    return 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))


@torch.no_grad()
def qwen2_forward(w: dict, cfg, input_ids: torch.Tensor, positions: torch.Tensor | None = None,
                  last_n: int | None = None) -> torch.Tensor:
    """w: HF-named state dict (any float dtype, used in fp32); cfg has hidden_size, num_hidden_layers,
    num_attention_heads, num_key_value_heads, head_dim, rms_norm_eps, rope_theta.
    input_ids [B, T] -> logits fp32 [B, T, V] (``last_n``: only the last n positions, [B, n, V] - the lm_head
    of a 152K vocabulary over every prefix position is most of the oracle's time at the BASELINE shapes)."""
    B, T = input_ids.shape
    nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    f = lambda name: w[name].float()
    x = f("model.embed_tokens.weight")[input_ids]
    pos = torch.arange(T).expand(B, T) if positions is None else positions
    fr = pos[..., None].float() * inv_freq(hd, cfg.rope_theta)
    emb = torch.cat((fr, fr), -1)
    cos, sin = emb.cos()[:, None], emb.sin()[:, None]            # [B, 1, T, hd]
    mask = torch.full((T, T), float("-inf")).triu(1)
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        h = rms_norm(x, f(p + "input_layernorm.weight"), cfg.rms_norm_eps)
        q = (h @ f(p + "self_attn.q_proj.weight").T + f(p + "self_attn.q_proj.bias")).view(B, T, nh, hd).transpose(1, 2)
        k = (h @ f(p + "self_attn.k_proj.weight").T + f(p + "self_attn.k_proj.bias")).view(B, T, nkv, hd).transpose(1, 2)
        v = (h @ f(p + "self_attn.v_proj.weight").T + f(p + "self_attn.v_proj.bias")).view(B, T, nkv, hd).transpose(1, 2)
        q = q * cos + rotate_half(q) * sin
        k = k * cos + rotate_half(k) * sin
        k = k.repeat_interleave(nh // nkv, dim=1)
        v = v.repeat_interleave(nh // nkv, dim=1)
        s = q @ k.transpose(-1, -2) / math.sqrt(hd) + mask
        a = torch.softmax(s, -1) @ v
        x = x + a.transpose(1, 2).reshape(B, T, nh * hd) @ f(p + "self_attn.o_proj.weight").T
        h = rms_norm(x, f(p + "post_attention_layernorm.weight"), cfg.rms_norm_eps)
        g = torch.nn.functional.silu(h @ f(p + "mlp.gate_proj.weight").T) * (h @ f(p + "mlp.up_proj.weight").T)
        x = x + g @ f(p + "mlp.down_proj.weight").T
    if last_n is not None:
        x = x[:, -last_n:]
    x = rms_norm(x, f("model.norm.weight"), cfg.rms_norm_eps)
    head = w.get("lm_head.weight", w["model.embed_tokens.weight"]).float()
    return x @ head.T
