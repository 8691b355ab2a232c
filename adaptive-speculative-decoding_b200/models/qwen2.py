"""Qwen2.5 shapes (public HF config.json values), random initialisation and the
packing of HF-named weights into the layouts ``asd_engine_set_layer`` expects (include/asd_b200.h)."""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Dict

import torch


@dataclass(frozen=True)
class Qwen2Config:
    hidden_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    intermediate_size: int
    vocab_size: int
    head_dim: int = 128
    tie_word_embeddings: bool = False
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e6
    name: str = "custom"

    @property
    def params(self) -> int:
        h, L = self.hidden_size, self.num_hidden_layers
        qkv = (self.num_attention_heads + 2 * self.num_key_value_heads) * self.head_dim
        per = qkv * h + qkv + self.num_attention_heads * self.head_dim * h + 3 * h * self.intermediate_size + 2 * h
        return L * per + h + self.vocab_size * h * (1 if self.tie_word_embeddings else 2)

    def streamed_bytes(self) -> int:
        """bf16 weight bytes one forward step must read (embedding gather excluded)."""
        h, L = self.hidden_size, self.num_hidden_layers
        qkv = (self.num_attention_heads + 2 * self.num_key_value_heads) * self.head_dim
        per = qkv * h + self.num_attention_heads * self.head_dim * h + 3 * h * self.intermediate_size
        return 2 * (L * per + self.vocab_size * h)

    def kv_bytes_per_token(self) -> int:
        return 2 * self.num_key_value_heads * self.head_dim * 2 * self.num_hidden_layers


QWEN25 = {
    "0.5b": Qwen2Config(896, 24, 14, 2, 4864, 151936, head_dim=64, tie_word_embeddings=True, name="Qwen2.5-0.5B"),
    "1.5b": Qwen2Config(1536, 28, 12, 2, 8960, 151936, tie_word_embeddings=True, name="Qwen2.5-1.5B"),
    "7b": Qwen2Config(3584, 28, 28, 4, 18944, 152064, name="Qwen2.5-7B"),
    "14b": Qwen2Config(5120, 48, 40, 8, 13824, 152064, name="Qwen2.5-14B"),
    "32b": Qwen2Config(5120, 64, 40, 8, 27648, 152064, name="Qwen2.5-32B"),
    "72b": Qwen2Config(8192, 80, 64, 8, 29568, 152064, name="Qwen2.5-72B"),
}
# the reference's Llama-era size labels (pipeline.py:175) map onto the Qwen2.5 cascade it describes
SIZE_ALIASES = {"8b": "7b", "13b": "14b", "34b": "32b", "70b": "72b"}


def get_config(size: str) -> Qwen2Config:
    s = size.lower()
    return QWEN25[SIZE_ALIASES.get(s, s)]


def tiny_config(**kw) -> Qwen2Config:
    base = Qwen2Config(128, 2, 4, 2, 256, 512, head_dim=64, name="tiny")
    return replace(base, **kw)


def random_hf_weights(cfg: Qwen2Config, seed: int, device="cpu", std: float = 0.02,
                      dtype=torch.bfloat16, logit_std: float = None) -> Dict[str, torch.Tensor]:
    """HF-named state dict with N(0, std^2) weights (norm weights 1 + N(0, std^2)), one seed per model
    (SURVEY.md 8d synthetic inputs).  ``logit_std`` rescales the output embedding so that the logits have that
    standard deviation (the final RMSNorm leaves unit-RMS activations, so logit std = head std * sqrt(hidden)):
    the north-star tolerance "max-abs <= 2e-2" is an absolute bound stated for logits of ordinary magnitude,
    and logit_std = 0.4 keeps |logit| < ~2 over a 152K vocabulary."""
    g = torch.Generator(device=device).manual_seed(seed)
    h, nh, nkv, hd, ff = (cfg.hidden_size, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim,
                          cfg.intermediate_size)

    def rn(*shape, mean=0.0):
        return (torch.randn(*shape, generator=g, device=device) * std + mean).to(dtype)

    w = {"model.embed_tokens.weight": rn(cfg.vocab_size, h)}
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        w[p + "self_attn.q_proj.weight"] = rn(nh * hd, h)
        w[p + "self_attn.q_proj.bias"] = rn(nh * hd)
        w[p + "self_attn.k_proj.weight"] = rn(nkv * hd, h)
        w[p + "self_attn.k_proj.bias"] = rn(nkv * hd)
        w[p + "self_attn.v_proj.weight"] = rn(nkv * hd, h)
        w[p + "self_attn.v_proj.bias"] = rn(nkv * hd)
        w[p + "self_attn.o_proj.weight"] = rn(h, nh * hd)
        w[p + "mlp.gate_proj.weight"] = rn(ff, h)
        w[p + "mlp.up_proj.weight"] = rn(ff, h)
        w[p + "mlp.down_proj.weight"] = rn(h, ff)
        w[p + "input_layernorm.weight"] = rn(h, mean=1.0)
        w[p + "post_attention_layernorm.weight"] = rn(h, mean=1.0)
    w["model.norm.weight"] = rn(h, mean=1.0)
    if not cfg.tie_word_embeddings:
        w["lm_head.weight"] = rn(cfg.vocab_size, h)
    if logit_std is not None:
        key = "model.embed_tokens.weight" if cfg.tie_word_embeddings else "lm_head.weight"
        w[key] = (w[key].float() * (logit_std / (std * h ** 0.5))).to(dtype)
    return w


def interleave_gate_up(gate_w: torch.Tensor, up_w: torch.Tensor) -> torch.Tensor:
    ff, K = gate_w.shape
    ffp = (ff + 63) // 64 * 64
    g = torch.zeros(ffp, K, dtype=gate_w.dtype, device=gate_w.device)
    u = torch.zeros(ffp, K, dtype=up_w.dtype, device=up_w.device)
    g[:ff], u[:ff] = gate_w, up_w
    return torch.stack([g.view(ffp // 64, 64, K), u.view(ffp // 64, 64, K)], dim=1).reshape(2 * ffp, K).contiguous()


def pack_layer(w: Dict[str, torch.Tensor], cfg: Qwen2Config, l: int, tp_rank: int = 0, tp_size: int = 1,
               device="cuda", fold_norm: bool = False) -> Dict[str, torch.Tensor]:
    """One layer of HF-named weights -> engine layout for one tensor-parallel rank (Megatron split:
    heads and ffn columns are sharded, O / down are sharded along their input dimension).
    ``fold_norm``: multiply the RMSNorm weights into the columns of the projections that consume the
    normalised activations (engine option fuse_norm: the GEMM then reads the raw bf16 residual and its
    epilogue applies the per-token rstd)."""
    p = f"model.layers.{l}."
    nh, nkv, hd, ff = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.intermediate_size
    assert nh % tp_size == 0 and nkv % tp_size == 0 and ff % tp_size == 0
    nhl, nkvl, ffl = nh // tp_size, nkv // tp_size, ff // tp_size
    r = tp_rank
    dev = lambda t: t.to(device=device, dtype=torch.bfloat16).contiguous()
    sl = lambda t, n: t[r * n:(r + 1) * n]
    wq, wk, wv = (w[p + f"self_attn.{n}_proj.weight"] for n in "qkv")
    bq, bk, bv = (w[p + f"self_attn.{n}_proj.bias"] for n in "qkv")
    wg, wu = w[p + "mlp.gate_proj.weight"], w[p + "mlp.up_proj.weight"]
    if fold_norm:
        ln1 = w[p + "input_layernorm.weight"].float()[None, :]
        ln2 = w[p + "post_attention_layernorm.weight"].float()[None, :]
        wq, wk, wv = ((t.float() * ln1).to(torch.bfloat16) for t in (wq, wk, wv))
        wg, wu = ((t.float() * ln2).to(torch.bfloat16) for t in (wg, wu))
    out = {
        "wqkv": dev(torch.cat([sl(wq, nhl * hd), sl(wk, nkvl * hd), sl(wv, nkvl * hd)], 0)),
        "bqkv": dev(torch.cat([sl(bq, nhl * hd), sl(bk, nkvl * hd), sl(bv, nkvl * hd)], 0)),
        "wo": dev(w[p + "self_attn.o_proj.weight"][:, r * nhl * hd:(r + 1) * nhl * hd]),
        "wgateup": dev(interleave_gate_up(sl(wg, ffl), sl(wu, ffl))),
        "wdown": dev(w[p + "mlp.down_proj.weight"][:, r * ffl:(r + 1) * ffl]),
        "ln1": dev(w[p + "input_layernorm.weight"]),
        "ln2": dev(w[p + "post_attention_layernorm.weight"]),
    }
    if tp_size > 1 and r != 0:
        out["bqkv"] = out["bqkv"]  # bias is column-parallel: every rank keeps its own slice
    return out


def random_packed_layer(cfg: Qwen2Config, gen: torch.Generator, tp_size: int = 1, device="cuda", std: float = 0.02,
                        gen_replicated: torch.Generator = None):
    """Random weights generated directly in the engine layout (for the full-size benchmarks, where a
    second HF-layout copy of 64-143 GB would not fit).  ``gen`` seeds the sharded matrices (one stream per
    rank), ``gen_replicated`` the RMSNorm weights, which every rank must hold identically."""
    h, hd = cfg.hidden_size, cfg.head_dim
    nhl, nkvl, ffl = cfg.num_attention_heads // tp_size, cfg.num_key_value_heads // tp_size, cfg.intermediate_size // tp_size
    nqkv, ffp = (nhl + 2 * nkvl) * hd, (ffl + 63) // 64 * 64

    def rn(*shape, mean=0.0):
        return (torch.randn(*shape, generator=gen, device=device) * std + mean).to(torch.bfloat16)

    wgu = rn(2 * ffp, h)
    if ffp != ffl:
        wgu.view(ffp // 64, 2, 64, h)[-1, :, ffl - (ffp - 64):, :] = 0
    out = {"wqkv": rn(nqkv, h), "bqkv": rn(nqkv), "wo": rn(h, nhl * hd), "wgateup": wgu, "wdown": rn(h, ffl)}
    gr = gen_replicated if gen_replicated is not None else gen
    for name in ("ln1", "ln2"):
        out[name] = (torch.randn(h, generator=gr, device=device) * std + 1.0).to(torch.bfloat16)
    return out


def load_safetensors_dir(path: str) -> Dict[str, torch.Tensor]:
    """HF checkpoint directory (``*.safetensors`` shards, as ``scripts/download_qwen3_models.py`` of the reference
    stores them) -> HF-named state dict on the CPU.  Parsed directly (8-byte header length, JSON header, raw
    little-endian tensors): the ``safetensors`` package is not a dependency."""
    import glob
    import json
    import os
    import struct

    import numpy as np
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {path}")
    dtypes = {"BF16": (torch.bfloat16, 2), "F16": (torch.float16, 2), "F32": (torch.float32, 4)}
    out: Dict[str, torch.Tensor] = {}
    for fn in files:
        with open(fn, "rb") as f:
            (n,) = struct.unpack("<Q", f.read(8))
            header = json.loads(f.read(n))
        base = 8 + n
        mm = np.memmap(fn, dtype=np.uint8, mode="r")
        for name, meta in header.items():
            if name == "__metadata__":
                continue
            if meta["dtype"] not in dtypes:
                raise ValueError(f"{fn}: tensor {name} has unsupported dtype {meta['dtype']}")
            dt, _ = dtypes[meta["dtype"]]
            b, e = meta["data_offsets"]
            raw = torch.from_numpy(np.array(mm[base + b:base + e]))       # copy out of the mapping
            out[name] = raw.view(dt).reshape(meta["shape"])
    return out


def save_safetensors(w: Dict[str, torch.Tensor], filename: str) -> None:
    """inverse of ``load_safetensors_dir`` for one shard (tests, tools)"""
    import json
    import struct
    names = {torch.bfloat16: "BF16", torch.float16: "F16", torch.float32: "F32"}
    header, blobs, off = {}, [], 0
    for k, t in w.items():
        t = t.contiguous().cpu()
        raw = t.view(torch.uint8).numpy().tobytes() if t.numel() else b""
        header[k] = {"dtype": names[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    hb = json.dumps(header).encode()
    hb += b" " * (-len(hb) % 8)
    with open(filename, "wb") as f:
        f.write(struct.pack("<Q", len(hb)))
        f.write(hb)
        for b in blobs:
            f.write(b)
