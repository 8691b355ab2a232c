#!/usr/bin/env python
"""Times the 72B shard GEMM shapes of BASELINE configs[4] (576 tokens) on both kernels:
the weight-streaming kernel (asd_linear_bf16) and the tensor-bound CTA-pair kernel (asd_linear_bf16_tc)."""
import ctypes
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from asd_b200 import lib
from asd_b200._lib import check


def time_call(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 576
    h, ff, V = 8192, 29568, 152064
    L = lib()
    stream = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rows = []
    for tp in (2, 4, 8):
        ffl = ff // tp
        ffp = (ffl + 63) // 64 * 64
        shapes = [("qkv", (64 + 16) // tp * 128, h, 0), ("o", h, h // tp, 0), ("gateup", 2 * ffp, h, 2), ("down", h, ffl, 0)]
        if tp == 2:
            shapes.append(("lm_head", V, h, 0))
        for name, N, K, mode in shapes:
            x = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
            w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
            flops = 2.0 * M * N * K
            used = ctypes.c_int(0)
            out_tc = torch.empty(8 * M * N if mode == 0 else M * N // 2, dtype=torch.float32 if mode == 0 else torch.bfloat16, device="cuda")
            def tc():
                check(L.asd_linear_bf16_tc(x.data_ptr(), w.data_ptr(), out_tc.data_ptr(), M, N, K, mode, 0, 0, ctypes.byref(used), stream()), "tc")
            us_tc = time_call(tc)
            ks_tc = used.value
            out_ws = torch.empty(M * N, dtype=torch.float32, device="cuda")
            ws_mode = 3 if mode == 0 else 2
            def ws():
                check(L.asd_linear_bf16(x.data_ptr(), w.data_ptr(), out_ws.data_ptr(), M, N, K, ws_mode, 0, 0, ctypes.byref(used), stream()), "ws")
            us_ws = time_call(ws)
            rows.append(dict(tp=tp, gemm=name, M=M, N=N, K=K, tc_us=round(us_tc, 1), tc_tflops=round(flops / us_tc / 1e6, 1),
                             tc_ksplit=ks_tc, ws_us=round(us_ws, 1), ws_tflops=round(flops / us_ws / 1e6, 1)))
            print(json.dumps(rows[-1]), flush=True)
            del x, w, out_tc, out_ws
    return rows


if __name__ == "__main__":
    main()
