"""Short workload for ncu: the fused rejection-sampling kernel at BASELINE config 2 (B=64, k=8)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asd_b200.ops import RejectionSampler

B, k, V = 64, 8, 152064
tl = torch.randn(B, k + 1, V, device="cuda") * 2
dl = tl[:, :k].contiguous() + torch.randn(B, k, V, device="cuda")
dt = torch.randint(0, V, (B, k), device="cuda", dtype=torch.int32)
s = RejectionSampler(B, k)
for _ in range(3):
    s(tl, dl, dt, torch.rand(B, k, dtype=torch.float64, device="cuda"), torch.rand(B, dtype=torch.float64, device="cuda"), 0.7)
torch.cuda.synchronize()
print("ok")
