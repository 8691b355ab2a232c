"""One sampler call per shape for ncu (tools/microbench.py times it; this only launches it)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asd_b200.ops import RejectionSampler

B, k = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 8)
V, T = 152064, 0.7
tl = torch.randn(B, k + 1, V, device="cuda") * 2
dl = (tl[:, :k] + torch.randn(B, k, V, device="cuda")).contiguous()
dt = dl.argmax(-1).int()
ua = torch.rand(B, k, dtype=torch.float64, device="cuda")
ur = torch.rand(B, dtype=torch.float64, device="cuda")
s = RejectionSampler(B, k)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    s(tl, dl, dt, ua, ur, T)
torch.cuda.synchronize()
