"""Short workload for ncu: config-3 engines (7B draft -> 32B target, B=16, k=5, prefix 512) with a
synthetic KV prefix (no prefill), N draft-then-verify steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asd_b200.engine import QwenEngine, SpecDecoder
from asd_b200.models.qwen2 import QWEN25

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, k, prefix = 16, 5, 512
t = QwenEngine(QWEN25["32b"], max_seqs=B, max_seq_len=prefix + 64, max_tokens=256).load_random(1)
d = QwenEngine(QWEN25["7b"], max_seqs=B, max_seq_len=prefix + 64, max_tokens=256).load_random(0)
for e in (t, d):
    e.kv_pool.normal_(0, 0.5)
dec = SpecDecoder(t, d, B, k, 0.7)
tok = torch.randint(0, 152064, (B,), device="cuda", dtype=torch.int32)
dec.seed_state(prefix, tok, tok)
for i in range(steps):
    if i == steps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()      # ncu --profile-from-start off captures only the last step
    out = dec.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", out["accepted_len"].tolist())
