"""Development check of the persistent forward kernel: same inputs through the per-kernel path (persist = 0) and
the persistent kernel (persist = 1); prints the largest logit difference and the engine's device error word."""
import os
import sys
from dataclasses import replace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asd_b200.engine import QwenEngine
from asd_b200.models.qwen2 import QWEN25, Qwen2Config, random_hf_weights, tiny_config


def run(name, cfg, B, P, q, rows=None, max_tokens=256):
    w = random_hf_weights(cfg, seed=3, device="cuda", logit_std=0.25)
    ids = torch.randint(0, cfg.vocab_size, (B, P + q), generator=torch.Generator().manual_seed(1)).cuda().to(torch.int32)
    outs = []
    for persist in (0, 1):
        eng = QwenEngine(cfg, max_seqs=B, max_seq_len=P + q + 16, max_tokens=max_tokens).load_hf_weights(w)
        eng.set_option("persist", persist)
        slots = torch.arange(B, dtype=torch.int32, device="cuda")
        if P > 0:
            eng.prefill(ids[:, :P], slots, want_logits=False)
        start = torch.full((B,), P, dtype=torch.int32, device="cuda")
        out = eng.forward_uniform(ids[:, P:].contiguous(), start, slots, P + q, last_only=bool(rows))
        torch.cuda.synchronize()
        err = eng.tp_error()
        outs.append(out.float().cpu())
        print(f"  {name} persist={persist} device_error={err} logits mean|x|={out.abs().mean().item():.4f}", flush=True)
        eng.close()
    d = (outs[0] - outs[1]).abs().max().item()
    print(f"{name}: B={B} P={P} q={q} rows={'last' if rows else 'all'} max|persist - legacy| = {d:.5f}, "
          f"argmax agree {(outs[0].argmax(-1) == outs[1].argmax(-1)).float().mean().item():.4f}", flush=True)
    return d


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run("tiny-hd64", tiny_config(), 3, 32, 5)
        run("tiny-hd64-noprefix", tiny_config(), 2, 0, 7)
        run("g5-hd128", Qwen2Config(512, 2, 10, 2, 1024, 4096, head_dim=128, name="g5"), 4, 144, 6)
        run("g7-hd128", Qwen2Config(896, 2, 7, 1, 1280, 2048, head_dim=128, name="g7"), 2, 69, 1)
        run("g7-hd128-last", Qwen2Config(896, 2, 7, 1, 1280, 2048, head_dim=128, name="g7"), 4, 69, 2, rows=True)
    else:
        run("32b-x2", replace(QWEN25["32b"], num_hidden_layers=2), 16, 512, 6)
        run("7b-x2", replace(QWEN25["7b"], num_hidden_layers=2), 16, 512, 1)
        run("7b-x2-q2-last", replace(QWEN25["7b"], num_hidden_layers=2), 16, 512, 2, rows=True)
