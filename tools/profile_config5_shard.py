"""One rank's share of BASELINE configs[4] on ONE GPU (72B sharded t ways as local dims, 2 layers, no exchange):
for ncu launch lists of the per-rank GEMM / attention shapes."""
import os, sys
from dataclasses import replace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asd_b200.engine import QwenEngine
from asd_b200.models.qwen2 import QWEN25
t = int(sys.argv[1]) if len(sys.argv) > 1 else 4
opts = dict(a.split("=") for a in sys.argv[2:])
c = QWEN25["72b"]
cfg = replace(c, num_attention_heads=c.num_attention_heads // t, num_key_value_heads=c.num_key_value_heads // t,
              intermediate_size=c.intermediate_size // t // 64 * 64, num_hidden_layers=2)
B, q, prefix = 64, 9, 4096
M = B * q
eng = QwenEngine(cfg, max_seqs=B, max_seq_len=prefix + 64, max_tokens=M, device="cuda:0")
eng.load_random(seed=3)
eng.kv_pool.normal_()
for n, v in opts.items():
    eng.set_option(n, int(v))
toks = torch.randint(0, cfg.vocab_size, (B, q), device="cuda", dtype=torch.int32)
slots = torch.arange(B, dtype=torch.int32, device="cuda")
start = torch.full((B,), prefix, dtype=torch.int32, device="cuda")
for _ in range(4):
    eng.forward_uniform(toks, start, slots, prefix + q, want_logits=False)
torch.cuda.synchronize()
print("done")
