"""Per-CTA GEMM timeline of one tensor-parallel verify forward (run under torchrun, one rank per GPU).
Prints rank 0's view of a middle layer: stamps as in trace_gemm.py plus s13 = partial pushed to the peers,
s14 = every peer's partial arrived."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from asd_b200 import _lib
from asd_b200.engine import QwenEngine
from asd_b200.models.qwen2 import QWEN25
from asd_b200.parallel import NcclComm, init_distributed

rank, local, world = init_distributed("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
opts = dict(a.split("=") for a in sys.argv[1:])
B, k, prefix = 16, 5, 512
t = QwenEngine(QWEN25["32b"], max_seqs=B, max_seq_len=prefix + 64, max_tokens=256, tp_rank=rank, tp_size=world,
               device=dev).load_random(1)
comm = NcclComm(rank, world)
t.set_allreduce(comm.comm_ptr, comm.allreduce_fn_ptr)
t.enable_p2p()
for name, v in opts.items():
    t.set_option(name, int(v))
t.kv_pool.normal_(0, 0.5)
tok = torch.randint(0, 152064, (B, k + 1), device=dev, dtype=torch.int32)
start = torch.full((B,), prefix, dtype=torch.int32, device=dev)
slots = torch.arange(B, dtype=torch.int32, device=dev)
for _ in range(3):
    t.forward_uniform(tok, start, slots, prefix + k + 1)
torch.cuda.synchronize()
dist.barrier()
L = _lib.lib()
MAXL = 300
stride = L.asd_debug_gemm_trace(None, 0)
buf = torch.zeros(MAXL * stride, dtype=torch.int64, device=dev)
L.asd_debug_gemm_trace(buf.data_ptr(), MAXL)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
t.forward_uniform(tok, start, slots, prefix + k + 1)
ev[1].record()
torch.cuda.synchronize()
L.asd_debug_gemm_trace(None, 0)
dist.barrier()
for rr in range(world):
    dist.barrier()
    if rank != rr:
        continue
    print("---- rank", rank)
    tr = buf.cpu().numpy().reshape(MAXL, stride // 16, 16)
    used = [i for i in range(MAXL) if tr[i, 0, 0] != 0]
    print("forward ms", ev[0].elapsed_time(ev[1]), "launches", len(used))
    base = used[30 * 4]
    T0 = None
    for j in range(5):
        x = tr[base + j]
        x = x[x[:, 0] != 0].astype(np.int64)
        if T0 is None:
            T0 = x[:, 0].min()
        out = [f"ctas {len(x):4d}"]
        for c, nm in ((0, "entry"), (2, "upstream"), (3, "tile0"), (4, "mainloop"), (5, "bar1"), (7, "bar2"), (13, "pushed"),
                      (15, "flagged"), (14, "arrived"), (10, "owner"), (8, "exit")):
            v = x[:, c][x[:, c] > 0]
            out.append(f"{nm} " + ("-" if len(v) == 0 else f"{np.median(v - T0) / 1e3:6.1f}/{(v - T0).max() / 1e3:6.1f}"))
        print(" ".join(out), flush=True)
dist.barrier()
comm.destroy()
dist.destroy_process_group()
