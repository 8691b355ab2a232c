"""Attention kernel comparison on the three shapes the benches time (CUDA events per launch via the engine's
`profile` option, 2-layer models of the right width so the weights are small).  Not the bench contract.
  python tools/bench_attn.py            -> one JSON line per (shape, attn_impl)"""
import json
import os
import sys
from dataclasses import replace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asd_b200.engine import QwenEngine
from asd_b200.models.qwen2 import QWEN25

PEAK = 6548.8


PAGE = int(os.environ.get("ASD_PAGE", "16"))


def run(name, cfg, B, q, prefix, tp, impls=(1, 2), opts=()):
    cfg2 = replace(cfg, num_hidden_layers=2)
    M = B * q
    for impl in impls:
        eng = QwenEngine(cfg2, max_seqs=B, max_seq_len=prefix + 64, max_tokens=max(M, 256), page_size=PAGE,
                         tp_rank=0, tp_size=tp, device="cuda:0")
        eng.load_random(seed=3)
        if tp > 1:   # a lone rank of a tp-way split: boundaries through a no-op "all-reduce" is not available; use p2p off + tp 1 dims
            raise SystemExit("use local dims instead of tp")
        eng.kv_pool.normal_()
        eng.set_option("attn_impl", impl)
        for o in opts:
            n, v = o.split("=")
            eng.set_option(n, int(v))
        toks = torch.randint(0, cfg.vocab_size, (B, q), device="cuda", dtype=torch.int32)
        slots = torch.arange(B, dtype=torch.int32, device="cuda")
        start = torch.full((B,), prefix, dtype=torch.int32, device="cuda")
        f = lambda: eng.forward_uniform(toks, start, slots, prefix + q, want_logits=False)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        eng.set_option("profile", 1)
        for _ in range(5):
            f()
        prof = eng.profile_read()
        eng.set_option("profile", 0)
        ms, n = prof["attention"]
        us = ms / n * 1e3
        nkv, hd = cfg2.num_key_value_heads, cfg2.head_dim
        nbytes = B * (prefix + q) * 2 * nkv * hd * 2
        print(json.dumps(dict(shape=name, impl=impl, B=B, q=q, prefix=prefix, nh=cfg2.num_attention_heads, nkv=nkv,
                              us=round(us, 2), GBs=round(nbytes / us / 1e3, 1), frac=round(nbytes / us / 1e3 / PEAK, 3),
                              opts=list(opts), page=PAGE)), flush=True)
        eng.close()
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    c32, c7, c72 = QWEN25["32b"], QWEN25["7b"], QWEN25["72b"]
    shard = lambda c, t: replace(c, num_attention_heads=c.num_attention_heads // t,
                                 num_key_value_heads=c.num_key_value_heads // t, intermediate_size=c.intermediate_size // t // 64 * 64)
    run("32b-verify", c32, 16, 6, 512, 1)
    run("7b-draft", c7, 16, 1, 512, 1)
    run("7b-draft2", c7, 16, 2, 512, 1)
    run("72b-tp4-verify", shard(c72, 4), 64, 9, 4096, 1)
    run("72b-tp8-verify", shard(c72, 8), 64, 9, 4096, 1)
    run("72b-tp2-verify", shard(c72, 2), 64, 9, 4096, 1)
    run("32b-verify-4k", c32, 16, 6, 4096, 1)
