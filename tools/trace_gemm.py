"""Per-CTA timeline of every weight-streaming GEMM of one draft-then-verify step (asd_debug_gemm_trace).

Prints, for each distinct GEMM of the verify forward (by position inside a layer), the median over the
middle layers of: launch-to-launch period, kernel span, and the phase stamps relative to the first CTA's
entry.  The raw stamps go to gpurun_out/gemm_trace.npy for offline reading."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from asd_b200 import _lib
from asd_b200.engine import QwenEngine, SpecDecoder
from asd_b200.models.qwen2 import QWEN25

B, k, prefix = 16, 5, 512
opts = dict(a.split("=") for a in sys.argv[1:])
t = QwenEngine(QWEN25["32b"], max_seqs=B, max_seq_len=prefix + 64, max_tokens=256).load_random(1)
d = QwenEngine(QWEN25["7b"], max_seqs=B, max_seq_len=prefix + 64, max_tokens=256).load_random(0)
for e in (t, d):
    e.kv_pool.normal_(0, 0.5)
    for name, v in opts.items():
        e.set_option(name, int(v))
dec = SpecDecoder(t, d, B, k, 0.7)
tok = torch.randint(0, 152064, (B,), device="cuda", dtype=torch.int32)
dec.seed_state(prefix, tok, tok)
for _ in range(3):
    dec.step()
torch.cuda.synchronize()
L = _lib.lib()
MAXL = 1000
stride = L.asd_debug_gemm_trace(None, 0)
buf = torch.zeros(MAXL * stride, dtype=torch.int64, device="cuda")
L.asd_debug_gemm_trace(buf.data_ptr(), MAXL)
dec.step()
torch.cuda.synchronize()
L.asd_debug_gemm_trace(None, 0)
tr = buf.cpu().numpy().reshape(MAXL, stride // 16, 16)
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/gemm_trace.npy", tr[:, :512].copy())

used = [i for i in range(MAXL) if tr[i, 0, 0] != 0]
print("launches traced:", len(used))
info = []
for i in used:
    x = tr[i]
    n = int((x[:, 0] != 0).sum())
    x = x[:n]
    info.append(dict(i=i, n=n, t0=x[:, 0].min(), t1=x[:, 8].max(), x=x))
# the verify forward is the tail: 64 layers x 4 GEMMs + lm_head
ver = info[-(64 * 4 + 2):]   # (slot 0 of a layer group = the previous layer's down projection)
names = ["down(l-1)", "qkv", "o", "gate|up"]
lab = ["entry", "setup", "upstream", "tile0", "mainloop", "s5", "s6", "s7", "exit", "smid", "s10", "s11", "s12"]
NS = 13
for which, part in (("verify", ver), ("draft", info[: 28 * 4 + 1])):
    nl = (len(part) - 1) // 4
    print(f"== {which}: {nl} layers, forward span {(part[-1]['t1'] - part[0]['t0']) / 1e3:.1f} us")
    for g in range(4):
        rows = []
        for l in range(nl // 4, 3 * nl // 4):
            cur = part[l * 4 + g]
            prev = part[l * 4 + g - 1]
            nxt = part[l * 4 + g + 1]
            x = cur["x"].astype(np.int64)
            rel = x[:, :NS] - cur["t0"]
            rel[rel < 0] = 0
            sms = len(set(x[:, 9].tolist()))
            x[:, 9] = cur["t0"]
            rows.append([cur["n"], sms, (nxt["t0"] - cur["t0"]) / 1e3, (cur["t1"] - cur["t0"]) / 1e3,
                         (cur["t0"] - prev["t1"]) / 1e3]
                        + [np.median(rel[:, j]) / 1e3 for j in range(NS)] + [rel[:, j].max() / 1e3 for j in range(NS)])
        r = np.median(np.array(rows), axis=0)
        print(f"{names[g]:8s} ctas {int(r[0]):4d} sms {int(r[1]):3d} period {r[2]:6.1f} span {r[3]:6.1f} gap_prev {r[4]:6.1f}")
        print("   median " + " ".join(f"{lab[j]} {r[5 + j]:5.1f}" for j in range(NS) if j != 9))
        print("   max    " + " ".join(f"{lab[j]} {r[5 + NS + j]:5.1f}" for j in range(NS) if j != 9))
