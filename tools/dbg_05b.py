import os, sys
sys.path.insert(0, os.getcwd())
from dataclasses import replace
import torch
from asd_b200.engine import QwenEngine
from asd_b200.models.qwen2 import QWEN25, random_hf_weights
from oracle.model_oracle import qwen2_forward
cfg = replace(QWEN25["0.5b"], num_hidden_layers=2)
w = random_hf_weights(cfg, seed=0)
ids = torch.randint(0, cfg.vocab_size, (1, 69), generator=torch.Generator().manual_seed(1234))
ref = qwen2_forward(w, cfg, ids)[:, -5:]
for fuse, opts in [(True, {}), (False, {}), (False, {"fuse_rope": 0}), (False, {"reduce": 0, "fuse_rope": 0})]:
    eng = QwenEngine(cfg, max_seqs=2, max_seq_len=96, max_tokens=64, fuse_norm=fuse).load_hf_weights(w)
    for k, v in opts.items():
        eng.set_option(k, v)
    slots = torch.ones(1, dtype=torch.int32, device="cuda")
    idc = ids.cuda().to(torch.int32)
    eng.prefill(idc[:, :64], slots)
    ver = eng.forward_uniform(idc[:, 64:].contiguous(), torch.full((1,), 64, dtype=torch.int32, device="cuda"), slots, 69)
    got = ver.view(1, 5, -1).cpu()
    d = (got - ref).abs()
    print(fuse, opts, "max", d.max().item(), "mean", d.mean().item(), "ref std", ref.std().item(), "agree", (got.argmax(-1) == ref.argmax(-1)).float().mean().item())
    eng.close()
