"""Per-CTA timeline of every weight-streaming GEMM of one draft-then-verify step (asd_debug_gemm_trace).
Stamps: entry, setup (barriers + TMEM), upstream (PDL wait returned), tile0 (first k-block landed), mainloop (last
MMA committed), bar1 / scatter / bar2 (cluster split-K reduction; for the SwiGLU GEMM: transposed / - / -), owner
(owner loop done), exit.  A layer launches QKV (56 tiles x 4 splits = 224 CTAs for 32B), O (320), gate|up (432), down (320).

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
abuf = torch.zeros(MAXL * 1024 * 16, dtype=torch.int64, device="cuda")
L.asd_debug_gemm_trace(buf.data_ptr(), MAXL)
L.asd_debug_attn_trace(abuf.data_ptr(), MAXL)
dec.step()
torch.cuda.synchronize()
L.asd_debug_gemm_trace(None, 0)
L.asd_debug_attn_trace(None, 0)
atr = abuf.cpu().numpy().reshape(MAXL, 1024, 16)
np.save("gpurun_out/attn_trace.npy", atr[:, :512].copy())
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
ver = info[-(64 * 4 + 1):]
names = ["g0", "g1", "g2", "g3"]   # launch order inside a layer: QKV, O, gate|up, down (identify by CTA count)
lab = ["entry", "setup", "upstream", "tile0", "mainloop", "bar1", "scatter", "bar2", "exit", "smid", "owner", "qkv-ep", "s12"]
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


# ---- absolute timeline of one mid layer (GEMMs + attention), times in us from the layer's first CTA entry
def timeline(gl, al, title):
    print(f"== {title}")
    ev = []
    for i in gl:
        x = tr[i]
        x = x[x[:, 0] != 0].astype(np.int64)
        ev.append(("gemm%4d" % len(x), x, [0, 2, 3, 4, 8]))
    x = atr[al]
    x = x[x[:, 0] != 0].astype(np.int64)
    ev.append(("attn%4d" % len(x), x, [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]))
    ev.sort(key=lambda e: e[1][:, 0].min())
    T0 = ev[0][1][:, 0].min()
    for name, x, cols in ev:
        out = []
        for c in cols:
            v = x[:, c][x[:, c] > 0]
            out.append(f"s{c}: " + ("-" if len(v) == 0 else f"{np.percentile(v - T0, 10) / 1e3:6.1f}/{np.median(v - T0) / 1e3:6.1f}/{(v - T0).max() / 1e3:6.1f}"))
        print(name, " ".join(out))


aused = [i for i in range(MAXL) if atr[i, 0, 0] != 0]
nd = 28
# draft forward #2 (0-based), layer 14; verify layer 30.  Per forward: 4 GEMMs per layer (+ lm_head), 1 attention.
ndraft_fw = (len(used) - (64 * 4 + 1)) // (nd * 4 + 1)
g0 = used[2 * (nd * 4 + 1) + 14 * 4]
timeline([g0 + j for j in range(5)], aused[2 * nd + 14], f"draft forward 2 layer 14 ({ndraft_fw} draft forwards)")
vb = len(used) - (64 * 4 + 1)
g0 = used[vb + 30 * 4]
timeline([g0 + j for j in range(5)], aused[ndraft_fw * nd + 30], "verify layer 30")
