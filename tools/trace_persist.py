"""Per-phase timeline of the persistent forward kernel (asd_engine_persist_trace): for every phase type the time from
the previous phase's last CTA to this phase's last CTA (critical path) and the skew between the first and the last
CTA to finish, averaged over layers; plus the forward time with and without the persistent kernel (CUDA events).
  python tools/trace_persist.py [32b|7b] [layers]"""
import ctypes
import os
import sys
from dataclasses import replace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from asd_b200 import lib
from asd_b200.engine import QwenEngine
from asd_b200.models.qwen2 import QWEN25

STRIDE = 643


def main(size, layers, B=16, prefix=512, q=None, opts=()):
    cfg = replace(QWEN25[size], num_hidden_layers=layers)
    q = q or (6 if size != "7b" else 1)
    eng = QwenEngine(cfg, max_seqs=B, max_seq_len=prefix + 64, max_tokens=256).load_random(seed=1)
    for o in opts:
        n, v = o.split("=")
        eng.set_option(n, int(v))
    eng.kv_pool.normal_()
    toks = torch.randint(0, cfg.vocab_size, (B, q), device="cuda", dtype=torch.int32)
    slots = torch.arange(B, dtype=torch.int32, device="cuda")
    start = torch.full((B,), prefix, dtype=torch.int32, device="cuda")
    logits = torch.empty(B * q, cfg.vocab_size, dtype=torch.float32, device="cuda")
    f = lambda: eng.forward_uniform(toks, start, slots, prefix + q, logits_out=logits)
    res = {}
    for persist in (0, 1):
        eng.set_option("persist", persist)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(10):
            f()
        ev[1].record()
        torch.cuda.synchronize()
        res[persist] = ev[0].elapsed_time(ev[1]) / 10
    wbytes = cfg.streamed_bytes()
    print(f"{cfg.name} x{layers} layers, M={B * q}: per-kernel path {res[0] * 1e3:.1f} us, persistent {res[1] * 1e3:.1f} us; "
          f"weights {wbytes / 1e6:.0f} MB -> {wbytes / res[1] / 1e6:.0f} GB/s persistent, {wbytes / res[0] / 1e6:.0f} GB/s per-kernel; "
          f"device_error={eng.tp_error()}")
    L = lib()
    L.asd_engine_persist_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    ncta = L.asd_engine_persist_trace(eng.h, None)
    buf = torch.zeros(8 * ncta * STRIDE, dtype=torch.int64, device="cuda")
    L.asd_engine_persist_trace(eng.h, ctypes.c_void_p(buf.data_ptr()))
    f()
    torch.cuda.synchronize()
    L.asd_engine_persist_trace(eng.h, None)
    planes = buf.view(8, ncta, STRIDE).cpu().numpy().astype(np.float64)
    t = planes[0]
    nph = 3 + 5 * layers
    t0 = t[:, 0].min()
    done = t[:, 1:nph + 1] - t0          # [cta, phase]
    last, first = done.max(0), done.min(0)
    names = ["qkv", "attn", "o", "gate_up", "down"]
    print(f"entry skew {(t[:, 0].max() - t0) / 1e3:.2f} us; embed done at {last[0] / 1e3:.2f} us")
    for w, nm in enumerate(names):
        crit, skew = [], []
        for l in range(1 if layers > 1 else 0, layers):     # skip layer 0 (cold)
            p = 1 + 5 * l + w
            crit.append(last[p] - last[p - 1])
            skew.append(last[p] - first[p])
        extra = ""
        if nm != "attn":
            f1, f2, f3 = [], [], []
            more = {4: [], 5: [], 6: [], 7: []}
            for l in range(1 if layers > 1 else 0, layers):
                p = 1 + 5 * l + w
                prev = last[p - 1]
                a1 = planes[1][:, p + 1] - t0
                a2 = planes[2][:, p + 1] - t0
                a3 = planes[3][:, p + 1] - t0
                ok = planes[1][:, p + 1] > 0
                f1.append((a1[ok] - prev).mean())
                f2.append((a2[ok] - prev).mean())
                ok3 = planes[3][:, p + 1] > 0
                f3.append((a3[ok3] - prev).mean() if ok3.any() else float("nan"))
                for pl_ in (4, 5, 6):
                    okp = planes[pl_][:, p + 1] > 0
                    more[pl_].append((planes[pl_][:, p + 1][okp] - t0 - prev).mean() if okp.any() else float("nan"))
                more[7].append((done[:, p] - prev).mean())
            extra = (f"   after prev phase: first acc {np.mean(f1) / 1e3:6.2f}, last acc {np.mean(f2) / 1e3:6.2f}, "
                     f"partial out {np.mean(more[4]) / 1e3:6.2f}, flags seen {np.mean(f3) / 1e3:6.2f}, reduced {np.mean(more[5]) / 1e3:6.2f}, "
                     f"epilogue {np.mean(more[6]) / 1e3:6.2f}, signalled {np.mean(more[7]) / 1e3:6.2f} us (CTA means)")
        print(f"  {nm:8s} critical path {np.mean(crit) / 1e3:7.2f} us   skew {np.mean(skew) / 1e3:6.2f} us{extra}")
    pl = 2 + 5 * layers
    print(f"  gather   {(last[pl - 1] - last[pl - 2]) / 1e3:7.2f} us; lm_head {(last[pl] - last[pl - 1]) / 1e3:7.2f} us; total {last[pl] / 1e3:.1f} us")
    per_layer = (last[5 * layers] - last[5]) / max(layers - 1, 1)
    lay_bytes = (wbytes - 2 * cfg.vocab_size * cfg.hidden_size) / layers
    print(f"  per layer {per_layer / 1e3:.2f} us = {lay_bytes / per_layer:.0f} GB/s of weights")
    eng.close()


if __name__ == "__main__":
    size = sys.argv[1] if len(sys.argv) > 1 else "32b"
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    main(size, layers, opts=sys.argv[3:])
