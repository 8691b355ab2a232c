"""Per-CTA timeline of the attention kernels (asd_debug_attn_trace) on the shapes the benches time.
  python tools/trace_attn.py [shape] [impl] [name=value ...]
Prints, per stamp slot, the median / min / max over CTAs of (stamp - first CTA entry) in microseconds, and the
span of the launch.  Slots of the tcgen05 kernel (impl 2): 0 entry, 1 Q staged, 2 S(0) ready, 3 P(0) written,
4 S(1) ready, 15 row maxima of tile 1 done, 10 P buffer free, 5 P(1) written, 6 O(0) folded, 7 S(last) ready,
8 loop end, 9 last fold, 11 loader: tile 0 issued, 12 loader done, 13 MMA: QK(0) issued, 14 MMA: PV(0) issued.
Slots of the mma.sync kernel (impl 1): see attention.cu (0 entry .. 9 merged)."""
import os
import sys
from dataclasses import replace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from asd_b200 import _lib
from asd_b200.engine import QwenEngine
from asd_b200.models.qwen2 import QWEN25

shape = sys.argv[1] if len(sys.argv) > 1 else "72b-tp4"
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
opts = dict(a.split("=") for a in sys.argv[3:])
c32, c7, c72 = QWEN25["32b"], QWEN25["7b"], QWEN25["72b"]
shard = lambda c, t: replace(c, num_attention_heads=c.num_attention_heads // t,
                             num_key_value_heads=c.num_key_value_heads // t,
                             intermediate_size=c.intermediate_size // t // 64 * 64)
cfg, B, q, prefix = {"72b-tp4": (shard(c72, 4), 64, 9, 4096), "72b-tp2": (shard(c72, 2), 64, 9, 4096),
                     "72b-tp8": (shard(c72, 8), 64, 9, 4096), "32b": (c32, 16, 6, 512), "7b": (c7, 16, 1, 512),
                     "32b-4k": (c32, 16, 6, 4096)}[shape]
cfg = replace(cfg, num_hidden_layers=2)
M = B * q
page = int(opts.pop("page", 16))
eng = QwenEngine(cfg, max_seqs=B, max_seq_len=prefix + 64, max_tokens=max(M, 256), page_size=page, device="cuda:0")
eng.load_random(seed=3)
eng.kv_pool.normal_()
eng.set_option("attn_impl", impl)
for n, v in opts.items():
    eng.set_option(n, int(v))
toks = torch.randint(0, cfg.vocab_size, (B, q), device="cuda", dtype=torch.int32)
slots = torch.arange(B, dtype=torch.int32, device="cuda")
start = torch.full((B,), prefix, dtype=torch.int32, device="cuda")
f = lambda: eng.forward_uniform(toks, start, slots, prefix + q, want_logits=False)
for _ in range(3):
    f()
torch.cuda.synchronize()
L = _lib.lib()
MAXL = 8
abuf = torch.zeros(MAXL * 1024 * 16, dtype=torch.int64, device="cuda")
L.asd_debug_attn_trace(abuf.data_ptr(), MAXL)
f()
torch.cuda.synchronize()
L.asd_debug_attn_trace(None, 0)
tr = abuf.cpu().numpy().reshape(MAXL, 1024, 16)
for li in range(MAXL):
    x = tr[li]
    live = x[:, 0] != 0
    if not live.any():
        continue
    x = x[live].astype(np.float64)
    t0 = x[:, 0].min()
    tmax = x[x > 0].max()
    print(f"launch {li}: {int(live.sum())} CTAs, span {(tmax - t0) / 1e3:.2f} us")
    for s in range(16):
        col = x[:, s]
        ok = col > 0
        if not ok.any():
            continue
        rel = (col[ok] - t0) / 1e3
        own = (col[ok] - x[ok, 0]) / 1e3
        print(f"  slot {s:2d}: n={int(ok.sum()):4d}  since launch med {np.median(rel):8.2f} min {rel.min():8.2f} "
              f"max {rel.max():8.2f} | since own entry med {np.median(own):8.2f} max {own.max():8.2f}")
