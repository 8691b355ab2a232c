"""N > 1 path on CPU (gloo, world_size 2): the Megatron split made by ``pack_layer`` (column-parallel
QKV / gate|up in the engine's interleaved layout, row-parallel O / down, all-reduce after each
row-parallel product) reproduces the unsharded oracle forward.  The arithmetic here is a torch fp32
emulation of the engine's data flow over the PACKED tensors, so the packing, the slicing and the
all-reduce placement are what is being tested."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from asd_b200.models.qwen2 import pack_layer, random_hf_weights, tiny_config
from asd_b200.parallel import shard_ranges
from oracle.model_oracle import inv_freq, qwen2_forward, rms_norm, rotate_half


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def packed_forward(cfg, w, ids, rank, world, allreduce):
    """full-sequence forward from the per-rank packed tensors (fp32 math)"""
    B, T = ids.shape
    nh, nkv, hd, ff = cfg.num_attention_heads // world, cfg.num_key_value_heads // world, cfg.head_dim, \
        cfg.intermediate_size // world
    x = w["model.embed_tokens.weight"].float()[ids]
    fr = torch.arange(T)[:, None].float() * inv_freq(hd, cfg.rope_theta)
    emb = torch.cat((fr, fr), -1)
    cos, sin = emb.cos(), emb.sin()
    mask = torch.full((T, T), float("-inf")).triu(1)
    for l in range(cfg.num_hidden_layers):
        pk = {k: v.float() for k, v in pack_layer(w, cfg, l, rank, world, device="cpu").items()}
        h = rms_norm(x, pk["ln1"], cfg.rms_norm_eps)
        qkv = h @ pk["wqkv"].T + pk["bqkv"]
        q, k, v = qkv.split([nh * hd, nkv * hd, nkv * hd], -1)
        q = q.view(B, T, nh, hd).transpose(1, 2)
        k = k.view(B, T, nkv, hd).transpose(1, 2)
        v = v.view(B, T, nkv, hd).transpose(1, 2)
        q, k = q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin
        k, v = k.repeat_interleave(nh // nkv, 1), v.repeat_interleave(nh // nkv, 1)
        a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd) + mask, -1) @ v
        part = a.transpose(1, 2).reshape(B, T, nh * hd) @ pk["wo"].T          # row-parallel: partial sum
        x = x + allreduce(part)
        h = rms_norm(x, pk["ln2"], cfg.rms_norm_eps)
        gu = (h @ pk["wgateup"].T).view(B, T, -1, 2, 64)                         # tiles of 64 gate | 64 up rows
        act = (torch.nn.functional.silu(gu[..., 0, :]) * gu[..., 1, :]).reshape(B, T, -1)[..., :ff]
        x = x + allreduce(act @ pk["wdown"].T)
    x = rms_norm(x, w["model.norm.weight"].float(), cfg.rms_norm_eps)
    return x @ w.get("lm_head.weight", w["model.embed_tokens.weight"]).float().T


def _worker(rank, world, port, out, heads=(4, 2, 384)):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1 if world > 2 else 2)
    cfg = tiny_config(num_attention_heads=heads[0], num_key_value_heads=heads[1], intermediate_size=heads[2])
    w = random_hf_weights(cfg, seed=3)
    ids = torch.randint(0, cfg.vocab_size, (2, 11), generator=torch.Generator().manual_seed(0))

    def allreduce(t):
        t = t.contiguous()
        dist.all_reduce(t)
        return t
    got = packed_forward(cfg, w, ids, rank, world, allreduce)
    ref = qwen2_forward(w, cfg, ids)
    out[rank] = float((got - ref).abs().max())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tensor_parallel_matches_unsharded():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world and all(v < 1e-4 for v in out.values()), dict(out)


@pytest.mark.parametrize("heads", [(40, 8, 1024), (64, 8, 1408)], ids=["32b-ratio", "72b-ratio"])
def test_eight_rank_split_at_the_target_head_ratios_matches_unsharded(heads):
    """Over 8 ranks Qwen2.5-32B keeps 5 query heads and ONE kv head per rank (group size 5) and an ffn shard that is a
    whole number of 64-row gate|up tiles; Qwen2.5-72B keeps 8 query heads, one kv head and an ffn shard of 3696 =
    57.75 tiles (the last gate|up tile of every rank is padded).  Same ratios on a small model: 40 / 64 heads, 8 kv
    heads, ffn 8 x 128 / 8 x 176."""
    world = 8
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, heads), nprocs=world, join=True)
    assert len(out) == world and all(v < 2e-4 for v in out.values()), dict(out)


def test_single_rank_packing_is_identity():
    cfg = tiny_config()
    w = random_hf_weights(cfg, seed=4)
    ids = torch.randint(0, cfg.vocab_size, (1, 9), generator=torch.Generator().manual_seed(1))
    got = packed_forward(cfg, w, ids, 0, 1, lambda t: t)
    assert (got - qwen2_forward(w, cfg, ids)).abs().max().item() < 1e-4


def test_shard_ranges():
    assert shard_ranges(8, 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    with pytest.raises(AssertionError):
        shard_ranges(7, 2)
