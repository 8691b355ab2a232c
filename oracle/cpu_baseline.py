"""ORACLE-side CPU baseline (used only by bench.py's ``cpu_baseline`` leg and ``--impl reference``).

Times the CPU restatement of the draft-then-verify step on the host cores on a BOUNDED sample of the
workload: ``sample_layers`` decoder layers of the target shape at M = B*(k+1) tokens and of the draft
shape at M = B tokens (fp32 math on bf16-valued weights, torch CPU with all host threads), both
lm_heads, and the C sampler oracle on real-sized logits; the per-layer time is then multiplied out to
the full depth.  The reference itself has no runnable CPU implementation of this path (its Stage is
vLLM on GPU), so the kind reported is "port"."""
from __future__ import annotations

import math
import os
import time

import numpy as np
import torch

from .model_oracle import inv_freq, rms_norm, rotate_half


def _layer(cfg, M, prefix, B, g):
    h, nh, nkv, hd, ff = (cfg.hidden_size, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim,
                          cfg.intermediate_size)
    rn = lambda *s: torch.randn(*s, generator=g) * 0.02
    w = dict(q=rn(nh * hd, h), k=rn(nkv * hd, h), v=rn(nkv * hd, h), o=rn(h, nh * hd), g=rn(ff, h), u=rn(ff, h),
             d=rn(h, ff), ln1=torch.ones(h), ln2=torch.ones(h))
    q_len = M // B
    kc = torch.randn(B, nkv, prefix + q_len, hd, generator=g)
    vc = torch.randn(B, nkv, prefix + q_len, hd, generator=g)
    return w, kc, vc, q_len


def _run_layer(cfg, w, x, kc, vc, B, q_len, prefix):
    nh, nkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    hcur = rms_norm(x, w["ln1"], cfg.rms_norm_eps)
    q = (hcur @ w["q"].T).view(B, q_len, nh, hd).transpose(1, 2)
    k = (hcur @ w["k"].T).view(B, q_len, nkv, hd).transpose(1, 2)
    v = (hcur @ w["v"].T).view(B, q_len, nkv, hd).transpose(1, 2)
    pos = torch.arange(prefix, prefix + q_len)
    fr = pos[:, None].float() * inv_freq(hd, cfg.rope_theta)
    emb = torch.cat((fr, fr), -1)
    cos, sin = emb.cos(), emb.sin()
    q = q * cos + rotate_half(q) * sin
    k = k * cos + rotate_half(k) * sin
    kc[:, :, prefix:] = k
    vc[:, :, prefix:] = v
    kk = kc.repeat_interleave(nh // nkv, dim=1)
    vv = vc.repeat_interleave(nh // nkv, dim=1)
    s = q @ kk.transpose(-1, -2) / math.sqrt(hd)
    mask = torch.full((q_len, prefix + q_len), float("-inf")).triu(prefix + 1)
    a = torch.softmax(s + mask, -1) @ vv
    x = x + a.transpose(1, 2).reshape(B * q_len, nh * hd) @ w["o"].T
    hcur = rms_norm(x, w["ln2"], cfg.rms_norm_eps)
    return x + (torch.nn.functional.silu(hcur @ w["g"].T) * (hcur @ w["u"].T)) @ w["d"].T


@torch.no_grad()
def time_model(cfg, B, q_len, prefix, sample_layers=2, repeats=2, seed=0):
    """seconds per decoder layer and for the lm_head at M = B*q_len tokens"""
    g = torch.Generator().manual_seed(seed)
    M = B * q_len
    w, kc, vc, _ = _layer(cfg, M, prefix, B, g)
    x = torch.randn(M, cfg.hidden_size, generator=g)
    _run_layer(cfg, w, x, kc, vc, B, q_len, prefix)          # warm-up
    t0 = time.perf_counter()
    for _ in range(sample_layers * repeats):
        x2 = _run_layer(cfg, w, x, kc, vc, B, q_len, prefix)
    t_layer = (time.perf_counter() - t0) / (sample_layers * repeats)
    head = torch.randn(cfg.vocab_size, cfg.hidden_size, generator=g) * 0.02
    t0 = time.perf_counter()
    logits = rms_norm(x2, w["ln1"], cfg.rms_norm_eps) @ head.T
    t_head = time.perf_counter() - t0
    return t_layer, t_head, logits


def spec_step_baseline(target_cfg, draft_cfg, B, k, prefix, temperature=0.7, sample_layers=2, threads=None):
    """Estimated CPU seconds for ONE draft-then-verify step and the tokens it emits."""
    from . import reject_sample
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    tl_t, th_t, logits_t = time_model(target_cfg, B, k + 1, prefix, sample_layers, seed=1)
    tl_d, th_d, logits_d = time_model(draft_cfg, B, 1, prefix, sample_layers, seed=0)
    V = target_cfg.vocab_size
    tlg = logits_t.view(B, k + 1, V).numpy()
    rng = np.random.default_rng(4321)
    dlg = (tlg[:, :k] + rng.standard_normal((B, k, V)).astype(np.float32) * float(tlg.std())).astype(np.float32)
    dt = np.argmax(dlg / max(temperature, 1e-3) + rng.gumbel(size=dlg.shape), -1).astype(np.int32)
    t0 = time.perf_counter()
    out = reject_sample(tlg, dlg, dt, rng.random((B, k)), rng.random(B), temperature)
    t_samp = time.perf_counter() - t0
    step = k * (tl_d * draft_cfg.num_hidden_layers + th_d) + tl_t * target_cfg.num_hidden_layers + th_t + t_samp
    return dict(step_seconds=step, tokens_per_step=float((out["accepted_len"] + 1).sum()),
                target_layer_s=tl_t, target_head_s=th_t, draft_layer_s=tl_d, draft_head_s=th_d, sampler_s=t_samp,
                threads=threads,
                sample=f"{sample_layers} of {target_cfg.num_hidden_layers} target layers at M={B*(k+1)} + "
                       f"{sample_layers} of {draft_cfg.num_hidden_layers} draft layers at M={B}, both lm_heads, "
                       f"C sampler oracle on B={B},k={k},V={V}; per-layer time scaled to full depth")
