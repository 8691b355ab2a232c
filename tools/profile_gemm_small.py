"""ncu target: small weight-streaming GEMMs (draft o_proj shape and a near-empty K) to study the fixed cost."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asd_b200 import lib
L = lib()
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
used = ctypes.c_int(0)
for (M, N, K) in [(16, 3584, 256), (16, 3584, 3584), (96, 5120, 5120)]:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    for _ in range(3):
        rc = L.asd_linear_bf16(x.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, 3, 0, 0, ctypes.byref(used), s)
        assert rc == 0
torch.cuda.synchronize()
print("ok")
