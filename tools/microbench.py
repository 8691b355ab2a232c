"""Kernel microbenchmarks (CUDA events, rotating buffers larger than L2).  Not the bench contract."""
import ctypes
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asd_b200 import lib
from asd_b200.ops import RejectionSampler, interleave_gate_up

PEAK = 6548.8


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2], ts[0]


def bench_gemm(M, N, K, mode=0, ksplit=0, stages=0, copies=None):
    L = lib()
    nb = N * K * 2
    copies = copies or max(2, int(300e6 // nb) + 1)
    ws = [(torch.randn(N, K, device="cuda") * 0.02).bfloat16() for _ in range(copies)]
    x = torch.randn(M, K, device="cuda").bfloat16()
    ks, st, tt = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    L.asd_linear_plan(M, N, K, mode, ctypes.byref(ks), ctypes.byref(st), ctypes.byref(tt))
    nsl = ksplit or ks.value
    out = torch.empty(max(nsl, 1) * M * N, dtype=torch.float32, device="cuda")
    used = ctypes.c_int(0)
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def fn(i):
        rc = L.asd_linear_bf16(x.data_ptr(), ws[i % copies].data_ptr(), out.data_ptr(), M, N, K, mode, ksplit, stages,
                               ctypes.byref(used), s)
        assert rc == 0, L.asd_last_error()
    med, best = timeit(fn)
    gbs = nb / med / 1e6
    print(json.dumps(dict(op="gemm", M=M, N=N, K=K, mode=mode, ksplit=used.value, stages=stages or st.value,
                          us=round(med * 1e3, 2), best_us=round(best * 1e3, 2), GBs=round(gbs, 1),
                          frac=round(gbs / PEAK, 3))), flush=True)


def bench_sampler(B, k, V=152064, T=0.7, copies=None):
    rows = B * (k + 1)
    nbytes = (B * k * 2 + B) * V * 4
    copies = copies or max(1, int(300e6 // nbytes) + 1)
    tls = [torch.randn(B, k + 1, V, device="cuda") * 2 for _ in range(copies)]
    dls = [t[:, :k].contiguous() + torch.randn(B, k, V, device="cuda") for t in tls]
    dt = torch.randint(0, V, (B, k), device="cuda", dtype=torch.int32)
    ua = torch.rand(B, k, dtype=torch.float64, device="cuda")
    ur = torch.rand(B, dtype=torch.float64, device="cuda")
    s = RejectionSampler(B, k)

    def fn(i):
        s(tls[i % copies], dls[i % copies], dt, ua, ur, T)
    med, best = timeit(fn)
    gbs = nbytes / med / 1e6
    print(json.dumps(dict(op="sampler", B=B, k=k, V=V, us=round(med * 1e3, 2), best_us=round(best * 1e3, 2),
                          GBs=round(gbs, 1), frac=round(gbs / PEAK, 3))), flush=True)


if __name__ == "__main__" and (len(sys.argv) < 2 or sys.argv[1] in ("all", "gemm", "sampler")):
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "gemm"):
        # Qwen2.5-32B verify shapes (M = 96) and 7B draft shapes (M = 16)
        for M, N, K, mode in [(96, 7168, 5120, 0), (96, 5120, 5120, 0), (96, 55296, 5120, 2), (96, 5120, 27648, 0),
                              (96, 152064, 5120, 0), (16, 4608, 3584, 0), (16, 3584, 3584, 0), (16, 37888, 3584, 2),
                              (16, 3584, 18944, 0), (16, 152064, 3584, 0)]:
            bench_gemm(M, N, K, mode)
        for ks in (1, 2, 3, 4, 6, 8):
            bench_gemm(96, 5120, 5120, 0, ksplit=ks)
        for st in (2, 3, 4, 6):
            bench_gemm(96, 55296, 5120, 2, stages=st)
    if what in ("all", "sampler"):
        for B, k in [(1, 8), (16, 5), (64, 8), (256, 8), (256, 1)]:
            bench_sampler(B, k)


def sweep():
    shapes = [("o32", 96, 5120, 5120), ("qkv32", 96, 7168, 5120), ("down32", 96, 5120, 27648),
              ("o7", 16, 3584, 3584), ("qkv7", 16, 4608, 3584), ("down7", 16, 3584, 18944)]
    for name, M, N, K in shapes:
        for ks in (1, 2, 3, 4, 6, 8, 12):
            for st in (2, 4, 6, 7, 10):
                if M == 96 and st > 7:
                    continue
                try:
                    bench_gemm(M, N, K, 0, ksplit=ks, stages=st)
                except AssertionError as e:
                    print("skip", name, ks, st, e)


if len(sys.argv) > 1 and sys.argv[1] == "sweep":
    sweep()


def bench_reduce():
    for M, N, K in [(96, 5120, 5120), (96, 7168, 5120), (96, 5120, 27648), (16, 3584, 3584), (16, 4608, 3584),
                    (16, 3584, 18944)]:
        bench_gemm(M, N, K, 0)
        for ks in (0, 2, 4, 8):
            bench_gemm(M, N, K, 3, ksplit=ks)


if len(sys.argv) > 1 and sys.argv[1] == "reduce":
    bench_reduce()


def bench_auto():
    for M, N, K, mode in [(96, 7168, 5120, 3), (96, 5120, 5120, 3), (96, 55296, 5120, 2), (96, 5120, 27648, 3),
                          (16, 4608, 3584, 3), (16, 3584, 3584, 3), (16, 37888, 3584, 2), (16, 3584, 18944, 3),
                          (16, 3584, 256, 3), (96, 5120, 256, 3)]:
        bench_gemm(M, N, K, mode)


if len(sys.argv) > 1 and sys.argv[1] == "auto":
    bench_auto()
