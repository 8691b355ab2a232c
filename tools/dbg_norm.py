import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, os.getcwd())
import numpy as np, torch
from asd_b200.engine import QwenEngine
from asd_b200.models.qwen2 import tiny_config
from oracle.model_oracle import qwen2_forward
z = np.load("tests/golden/qwen2_tiny_golden.npz")
w = {k[3:]: torch.from_numpy(z[k]).view(torch.bfloat16) for k in z.files if k.startswith("w::")}
ids = torch.from_numpy(z["input_ids"]).long(); ref = torch.from_numpy(z["logits"])
cfg = tiny_config()
def run(fuse, chunk):
    B, T = ids.shape
    eng = QwenEngine(cfg, max_seqs=B+1, max_seq_len=T+16, max_tokens=64, fuse_norm=fuse).load_hf_weights(w)
    slots = torch.arange(1, B+1, dtype=torch.int32, device="cuda")
    idc = ids.cuda().to(torch.int32)
    last = eng.prefill(idc[:, :T-4], slots, chunk=chunk)
    ver = eng.forward_uniform(idc[:, T-4:].contiguous(), torch.full((B,), T-4, dtype=torch.int32, device="cuda"), slots, T)
    torch.cuda.synchronize()
    return ver.view(B, 4, -1).cpu(), last.cpu()
for chunk in (7, 0):
    a, al = run(True, chunk); b, bl = run(False, chunk)
    r = ref[:, -4:]
    print("chunk", chunk, "fused-vs-ref", (a-r).abs().max().item(), "unfused-vs-ref", (b-r).abs().max().item(), "fused-vs-unfused", (a-b).abs().max().item(),
          "argmax fused", (a.argmax(-1)==r.argmax(-1)).float().mean().item(), "unfused", (b.argmax(-1)==r.argmax(-1)).float().mean().item(), "ref std", r.std().item())
    print("   last: fused", (al-ref[:, -5]).abs().max().item(), "unfused", (bl-ref[:, -5]).abs().max().item(), (al.argmax(-1)==ref[:, -5].argmax(-1)).tolist(), (bl.argmax(-1)==ref[:, -5].argmax(-1)).tolist())
