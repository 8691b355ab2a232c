#!/usr/bin/env python
"""bench.py - accepted tok/s of the draft-then-verify step (BASELINE.json metric).

Workload (BASELINE.json configs[2]): Qwen2.5-7B draft -> Qwen2.5-32B target, bf16, chain k = 5,
batch 16, 512-token prefix, temperature 0.7, random-init weights, synthetic prompts.
  --gpus 1 : both models on one B200.
  --gpus N : the 32B target tensor-parallel over N ranks (NCCL all-reduce on the two row-parallel
             boundaries of each layer), the 7B draft replicated per rank (it "stays single-GPU");
             total work is fixed -> "scaling": "strong".  --workload 72b selects the 72B target.
One "step" = k draft forwards + one (k+1)-token verify forward + one fused rejection-sampling launch
for the whole batch.  value = tokens emitted by all sequences / device time (CUDA events, max over ranks).

  --impl reference : the CPU restatement of the same step (oracle/cpu_baseline.py) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "accepted_tokens_per_second"
UNIT = "tok/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.03)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


class Watchdog(threading.Thread):
    """N > 1 only: the tensor-parallel kernels spin for their peers on the device and a stuck collective cannot be
    cancelled from the host, so a stage that overruns its budget would otherwise hang the whole job until the caller's
    own limit.  Every rank runs one; when a stage overruns, rank 0 prints what it has - the finished bench line with a
    "watchdog" entry if the timed region is done, else an error record - and every rank leaves with os._exit."""

    def __init__(self, rank, world):
        super().__init__(daemon=True)
        self.rank, self.world = rank, world
        self.name_, self.deadline, self.line, self.lock = "start", None, None, threading.Lock()
        self.printed_rc = None      # set once rank 0 has printed the line: a stuck teardown then just leaves with it
        self.marks = []             # (stage, start time): the line reports how long each stage took

    def stage(self, name, budget_s):
        with self.lock:
            self.name_, self.deadline = name, (time.time() + budget_s if budget_s else None)
            self.marks.append((name, time.time()))
        sys.stderr.write(f"[bench rank {self.rank}] {time.strftime('%H:%M:%S')} stage {name}\n")
        sys.stderr.flush()

    def seconds(self):
        m = self.marks + [("end", time.time())]
        return {m[i][0]: round(m[i + 1][1] - m[i][1], 2) for i in range(len(m) - 1)}

    def run(self):
        while True:
            time.sleep(0.5)
            with self.lock:
                name, deadline, line = self.name_, self.deadline, self.line
            if deadline is None or time.time() < deadline:
                continue
            msg = f"stage '{name}' did not finish within its budget on rank {self.rank}"
            sys.stderr.write(f"[bench rank {self.rank}] watchdog: {msg}\n")
            sys.stderr.flush()
            if self.printed_rc is not None:
                os._exit(self.printed_rc)
            code = 5
            if line is not None:
                code = 0 if name == "config5" else 3
                if self.rank == 0:
                    line = dict(line)
                    line["watchdog"] = msg
                    if name == "config5":
                        line["config5"] = {"error": msg}
                    else:
                        line["tp_check"] = {"ok": False, "error": msg}
                    print(json.dumps(line), flush=True)
            elif self.rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": self.world,
                                  "error": "watchdog: " + msg}), flush=True)
            os._exit(code)


def workload(args):
    from asd_b200.models.qwen2 import QWEN25
    target = QWEN25["72b" if args.workload == "72b" else "32b"]
    return dict(target=target, draft=QWEN25["7b"], B=args.batch, k=args.k, prefix=args.prefix, T=args.temperature)


def config_dict(wl, n_gpus, extra=None):
    c = {"workload": f"{wl['draft'].name} draft -> {wl['target'].name} target, bf16, chain k={wl['k']}, batch {wl['B']}, "
                     f"prefix {wl['prefix']}, temperature {wl['T']}, random-init weights, synthetic prompts",
         "batch": wl["B"], "k": wl["k"], "prefix": wl["prefix"], "temperature": wl["T"],
         "parallelism": "single-gpu" if n_gpus == 1 else (f"target tp{n_gpus} (" + ("all-reduce fused into the row-parallel GEMM epilogues over NVLink peer memory" if (n_gpus == 2 or (n_gpus - 1) * wl["B"] * (wl["k"] + 1) * wl["target"].hidden_size * 8 <= 5 << 20) else "one-kernel peer-memory all-reduce + residual + norm over NVLink") + "), draft replicated"),
         "l2": "inputs larger than L2 (64 GB of weights streamed per verify step; no flush needed)"}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """CPU arm.  The line's ``value`` is the same metric on the same config as our arm (configs[2]); a full
    7B -> 32B step on CPU takes minutes, so it is a bounded sample multiplied out (``"extrapolated": true``).
    Next to it ``cpu_baselines`` holds three MEASURED, un-extrapolated baselines: the stop
    rule over 10^5 triples in CPython, the C sampler oracle on two config-2 grid points, and configs[0] end to end
    (HF Qwen2ForCausalLM fp32, 0.5B -> 1.5B, k = 4, greedy) - our arm reports the same three on the GPU."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle  # noqa: F401  (the only place outside tests/smoke where bench executes oracle/)
    from oracle.cpu_baseline import spec_step_baseline
    wl = workload(args)
    oracle.build()
    cores = os.cpu_count()
    t_start = time.time()
    r = spec_step_baseline(wl["target"], wl["draft"], wl["B"], wl["k"], wl["prefix"], wl["T"], sample_layers=1,
                           threads=cores)
    value = r["tokens_per_step"] / r["step_seconds"]
    measured = {}
    if not args.no_cpu_baseline:
        measured = measured_cpu_baselines(cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["step_seconds"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32 (bf16-valued weights)",
            "data": "synthetic", "config": config_dict(wl, args.gpus), "extrapolated": True,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": r["sample"],
                             "extrapolated": True},
            "cpu_baselines": measured, "wall_seconds": round(time.time() - t_start, 1),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measured_cpu_baselines(cores):
    """the three measured CPU baselines (each runs to completion inside this call; nothing multiplied out)"""
    import numpy as np
    import oracle
    from oracle import config1_cpu, stop_rule_py
    from asd_b200.models.qwen2 import QWEN25
    out = {"stop_rule": stop_rule_py.time_triples(100_000, seed=7)}
    grid = []
    for B, k in ((16, 5), (64, 8)):
        rng = np.random.default_rng(4321)
        V, T = 152064, 0.7
        tl = (rng.standard_normal((B, k + 1, V), dtype=np.float32) * 2)
        dl = (tl[:, :k] + rng.standard_normal((B, k, V), dtype=np.float32))
        dt = dl.argmax(-1).astype(np.int32)
        t0 = time.perf_counter()
        oracle.reject_sample(tl, dl, dt, rng.random((B, k)), rng.random(B), T)
        dtm = time.perf_counter() - t0
        grid.append({"B": B, "k": k, "V": V, "seconds": dtm, "rows_per_second": B * (k + 1) / dtm,
                     "GB_per_s": (B * k * 2 * V * 4 + B * V * 4) / dtm / 1e9})
    out["sampler"] = {"kind": "port", "cores": 1, "what": "oracle/sampler_oracle.c (scalar C), fixed uniforms, T = 0.7",
                      "grid": grid}
    out["config1"] = dict(config1_cpu.run(QWEN25["0.5b"], QWEN25["1.5b"], k=4, prompt_len=64, steps=8, threads=cores),
                          kind="reference-stack (HF transformers, the model code the reference's data generation calls)")
    return out


def load_traffic():
    """DRAM bytes actually moved per algorithmic byte for the dominant kernel, from the committed ncu capture
    (profiles/traffic.json, written by tools/ncu_summary.py from `ncu --set full`); None if there is no capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def tp_check(torch, dist, rank, world, dev, dec, target, tcfg):
    """Driver-visible tensor-parallel parity (runs after the timed region, N > 1):
    1. every rank's verify logits of the last timed step must be bit-identical (ranks stay in lock step without
       exchanging tokens only if they are), 2. no rank may have timed out waiting for a peer
       (asd_engine_tp_error), 3. a 2-layer model of the target's width is run on M = 96 tokens by the N-way sharded
       engine and by a single-GPU engine on rank 0; their logits must agree within the north-star tolerance
       (max-abs <= 2e-2 at logit std 0.25; they differ by fp32 summation order, which flips bf16 roundings)."""
    from dataclasses import replace
    from asd_b200.engine import QwenEngine
    from asd_b200.models.qwen2 import random_hf_weights
    lg = dec.target_logits
    chk = torch.stack([lg.double().sum(), lg.double().abs().sum(), lg.view(-1)[::4099].double().square().sum(),
                       lg.argmax(-1).double().sum()])
    allc = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    identical = all(torch.equal(allc[0], c) for c in allc)
    err = torch.tensor([float(target.tp_error())], device=dev)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    cfg2 = replace(tcfg, num_hidden_layers=2)
    w = random_hf_weights(cfg2, seed=3, device=dev, logit_std=0.25)      # same seed, same device type: same weights on all ranks
    B, q = 16, 6
    ids = torch.randint(0, cfg2.vocab_size, (B, q), generator=torch.Generator().manual_seed(7)).to(dev).to(torch.int32)
    slots = torch.arange(B, dtype=torch.int32, device=dev)
    zero = torch.zeros(B, dtype=torch.int32, device=dev)
    sh = QwenEngine(cfg2, max_seqs=B, max_seq_len=64, max_tokens=256, tp_rank=rank, tp_size=world, device=dev)
    sh.load_hf_weights(w)
    sh.enable_p2p()
    got = sh.forward_uniform(ids, zero, slots, q)
    torch.cuda.synchronize()
    dist.barrier()
    err2 = torch.tensor([float(sh.tp_error())], device=dev)
    dist.all_reduce(err2, op=dist.ReduceOp.MAX)
    max_abs = None
    if rank == 0:
        one = QwenEngine(cfg2, max_seqs=B, max_seq_len=64, max_tokens=256, device=dev).load_hf_weights(w)
        ref = one.forward_uniform(ids, zero, slots, q)
        torch.cuda.synchronize()
        max_abs = float((got - ref).abs().max())
        top2 = ref.topk(2, -1).values
        dec = (top2[..., 0] - top2[..., 1]) > 4e-2          # rows an implementation within 2e-2 can decide
        agree = float((got.argmax(-1) == ref.argmax(-1))[dec].float().mean()) if bool(dec.any()) else 1.0
        one.close()
    sh.close()
    del w
    torch.cuda.empty_cache()
    out = {"ranks_identical": bool(identical), "tp_error": int(max(err.item(), err2.item()))}
    if rank == 0:
        out.update(max_abs_vs_tp1_sample=max_abs, argmax_agree_vs_tp1_sample=agree,
                   sample=f"2 layers of {tcfg.name} at M={B * q}, sharded x{world} vs one GPU")
        out["ok"] = bool(identical and out["tp_error"] == 0 and max_abs <= 2e-2 and agree >= 0.999)
    return out


def config5(torch, dist, rank, world, dev, args, peaks):
    """BASELINE configs[4]: Qwen2.5-72B verify-only at TP = N, batch 64, 4096-token prefix, k = 8 (576 tokens per
    forward), paged KV.  The KV prefix is N(0,1) bf16 written straight into the paged pool (SURVEY 8d allows it
    for the verify-only sweep: a real 64 x 4096-token prefill is 38 PFLOP); everything timed is the real forward."""
    from asd_b200.engine import QwenEngine
    from asd_b200.models.qwen2 import QWEN25
    cfg = QWEN25["72b"]
    B, k, prefix = 64, 8, 4096
    M = B * (k + 1)
    eng = QwenEngine(cfg, max_seqs=B, max_seq_len=prefix + 64, max_tokens=M, tp_rank=rank, tp_size=world, device=dev)
    eng.load_random(seed=2)
    if args.nccl_only:                # development aid: the exchange through ncclAllReduce instead of the peer kernels
        from asd_b200.parallel import NcclComm
        comm5 = NcclComm(rank, world)
        eng.set_allreduce(comm5.comm_ptr, comm5.allreduce_fn_ptr)
    else:
        eng.enable_p2p()
    for o in args.opt or []:          # development aid: engine options also reach this record
        name, val = o.split("=")
        eng.set_option(name, int(val))
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    pool = eng.kv_pool
    step = 1 << 28
    for o in range(0, pool.numel(), step):
        pool[o:o + step].normal_(generator=g)
    toks = torch.randint(0, cfg.vocab_size, (B, k + 1), generator=torch.Generator().manual_seed(5)).to(dev).to(torch.int32)
    slots = torch.arange(B, dtype=torch.int32, device=dev)
    start = torch.full((B,), prefix, dtype=torch.int32, device=dev)
    logits = torch.empty(M, cfg.vocab_size, dtype=torch.float32, device=dev)
    run = lambda: eng.forward_uniform(toks, start, slots, prefix + k + 1, logits_out=logits)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    dist.barrier()
    n = 8
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        run()
    ev[1].record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev[0].elapsed_time(ev[1]) / n], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    eng.set_option("profile", 1)
    run()
    prof = eng.profile_read()
    eng.set_option("profile", 0)
    tp_err = eng.tp_error()
    # per-rank view of the profiled classes: a rank that waits for a slower peer books the wait under "allreduce",
    # so the MINIMUM over ranks is the cost of the exchange itself and the spread is load imbalance between GPUs
    mine = torch.tensor([prof["gemm"][0], prof["attention"][0], prof["allreduce"][0]], dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    per_rank = {"gemm": [round(float(t[0]), 3) for t in allr], "attention": [round(float(t[1]), 3) for t in allr],
                "allreduce": [round(float(t[2]), 3) for t in allr]}
    dist.barrier()
    eng.close()
    del eng, pool, logits
    torch.cuda.empty_cache()
    ms = float(ms.item())
    params = cfg.streamed_bytes() / 2
    gemm_flops = 2.0 * M * params / world
    attn_flops = 4.0 * M * (prefix + k + 1) * cfg.num_attention_heads * cfg.head_dim * cfg.num_hidden_layers / world
    tensor_floor_ms = (gemm_flops + attn_flops) / (peaks["bf16_tflops_sustained"] * 1e12) * 1e3
    kv_bytes = B * prefix * cfg.kv_bytes_per_token() / world
    hbm_floor_ms = (cfg.streamed_bytes() / world + kv_bytes) / (peaks["hbm_gbs"] * 1e9) * 1e3
    return {"workload": f"{cfg.name} verify only, TP={world}, batch {B}, prefix {prefix} (N(0,1) KV), k={k}, {M} tokens/forward",
            "verify_us": ms * 1e3, "tensor_floor_ms": tensor_floor_ms, "hbm_floor_ms": hbm_floor_ms,
            "frac_of_tensor_floor": tensor_floor_ms / ms, "tflops_per_gpu": (gemm_flops + attn_flops) / (ms * 1e-3) / 1e12,
            "verified_tokens_per_second": M / (ms * 1e-3), "tp_error": int(tp_err),
            "breakdown_ms_profiled": {c: round(v[0], 3) for c, v in prof.items()},
            "breakdown_ms_per_rank": per_rank,
            "allreduce_bytes_per_boundary": M * cfg.hidden_size * 4, "boundaries": 2 * cfg.num_hidden_layers}


def companions(torch, dev, peak):
    """Single-GPU companions of the measured CPU baselines of the reference arm (same inputs, same units)."""
    import numpy as np
    from asd_b200.algorithms import dp_solver
    from asd_b200.engine import QwenEngine, SpecDecoder
    from asd_b200.models.qwen2 import QWEN25
    from asd_b200.ops import RejectionSampler
    out = {}
    # stop rule: the same 10^5 triples (L = 4 rows; L = 3 rows carry p = 1, C = 0 in the unused stage... kept simple:
    # two launches, one per L), Bayesian shrinkage on the device path via risk_adjustment
    rng = np.random.default_rng(7)
    n = 100_000
    lam = torch.from_numpy(rng.choice([0.1, 0.5, 1.0, 2.0, 5.0, 10.0], n)).to(dev)
    p = torch.from_numpy(rng.random((n, 4))).to(dev)
    C = torch.tensor([1.0, 2.0, 4.5, 10.0], dtype=torch.float64, device=dev).expand(n, 4).contiguous()
    p[:, 3] = 1.0
    for _ in range(3):
        dp_solver.stop_rule_rows(p, C, lam, True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        k_star, _ = dp_solver.stop_rule_rows(p, C, lam, True)
    ev[1].record()
    torch.cuda.synchronize()
    out["stop_rule"] = {"decisions_per_second": n / (ev[0].elapsed_time(ev[1]) / 10 * 1e-3), "n": n,
                        "what": "asd_stop_rule_rows on the device, one launch for all rows"}
    grid = []
    for B, k in ((1, 8), (16, 5), (64, 8), (256, 8)):
        V, T = 152064, 0.7
        g = torch.Generator(device=dev).manual_seed(4321)
        tl = torch.randn(B, k + 1, V, device=dev, generator=g) * 2
        dl = (tl[:, :k] + torch.randn(B, k, V, device=dev, generator=g)).contiguous()
        dt = dl.argmax(-1).int()
        ua = torch.rand(B, k, dtype=torch.float64, device=dev, generator=g)
        ur = torch.rand(B, dtype=torch.float64, device=dev, generator=g)
        smp = RejectionSampler(B, k, dev)
        flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
        times = []
        for it in range(8):
            flush.zero_()                                 # L2 flush between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            smp(tl, dl, dt, ua, ur, T)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        us = sorted(times)[len(times) // 2] * 1e3
        nbytes = B * k * 2 * V * 4 + B * V * 4
        grid.append({"B": B, "k": k, "V": V, "us": us, "GB_per_s": nbytes / us / 1e3, "hbm_frac": nbytes / us / 1e3 / peak})
        del tl, dl, flush
    out["sampler"] = {"grid": grid, "what": "asd_reject_sample, L2 flushed between launches, median of 5"}
    torch.cuda.empty_cache()
    # configs[0]: 0.5B -> 1.5B, k = 4, batch 1, greedy
    t = QwenEngine(QWEN25["1.5b"], max_seqs=1, max_seq_len=1024, max_tokens=64, device=dev).load_random(seed=1)
    d = QwenEngine(QWEN25["0.5b"], max_seqs=1, max_seq_len=1024, max_tokens=64, device=dev).load_random(seed=0)
    dec = SpecDecoder(t, d, 1, 4, 0.0)
    dec.prefill(torch.randint(0, QWEN25["1.5b"].vocab_size, (1, 64), generator=torch.Generator().manual_seed(1234)))
    for _ in range(3):
        dec.step()
    em = torch.zeros((), dtype=torch.int64, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(32):
        em += (dec.step()["accepted_len"].to(torch.int64) + 1).sum()
    ev[1].record()
    torch.cuda.synchronize()
    out["config1"] = {"value": int(em.item()) / (ev[0].elapsed_time(ev[1]) * 1e-3), "unit": "tok/s", "steps": 32,
                      "what": "Qwen2.5-0.5B -> Qwen2.5-1.5B, chain k=4, batch 1, greedy, 64-token prompt, on the GPU"}
    t.close()
    d.close()
    torch.cuda.empty_cache()
    return out


def stage_generate_record(torch, wl, dev, max_tokens=96):
    """the reference-facing seam itself: Stage.generate(prompts: list[str]) -> texts, on the bench's workload
    (strings in, strings out; prefill, detokenisation and the per-request bookkeeping included)"""
    from asd_b200.models.stage import Stage
    B, prefix, k, T = wl["B"], wl["prefix"], wl["k"], wl["T"]
    gpu = dev.index or 0
    small = Stage("bench-draft", "7b", config=wl["draft"], seed=0, max_batch=B, max_model_len=prefix * 2 + 256, k=k, gpu_ids=[gpu])
    big = Stage("bench-target", "32b", config=wl["target"], seed=1, draft=small, max_batch=B, max_model_len=prefix * 2 + 256,
                k=k, gpu_ids=[gpu])
    import random
    rnd = random.Random(1234)
    prompts = ["".join(chr(rnd.randrange(97, 123)) for _ in range(prefix)) for _ in range(B)]   # byte tokenizer: 1 token / char
    big.generate(prompts[:B], max_tokens=8, temperature=T)                                          # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    texts, lps, st = big.generate(prompts, max_tokens=max_tokens, temperature=T)
    dt = time.perf_counter() - t0
    n = sum(len(lp) for lp in lps)
    rec = {"value": n / dt, "unit": "tok/s", "wall_seconds": dt, "tokens": n, "decode_steps": st["decode_steps"],
           "api": "Stage.generate(prompts=[str]*16, max_tokens, temperature) -> (texts, logprobs, stats)",
           "includes": f"tokenisation, ragged prefill of {B} x {prefix} tokens on both models, decode, detokenisation"}
    big.engine.close()
    small.engine.close()
    return rec



# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import asd_b200
    from asd_b200.engine import QwenEngine, SpecDecoder
    from asd_b200.parallel import NcclComm, init_distributed

    rank, local, world = init_distributed("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = workload(args)
    wd = Watchdog(rank, world)
    if world > 1:
        wd.start()
    wd.stage("setup + prefill", 180 if world > 1 else 0)
    if args.config5_only:       # development aid: only the 72B verify-only record
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        c5 = config5(torch, dist, rank, world, dev, args, peaks)
        if rank == 0:
            print(json.dumps({"config5": c5}), flush=True)
        dist.destroy_process_group()
        return
    B, k, prefix, T = wl["B"], wl["k"], wl["prefix"], wl["T"]
    n_steps_total = args.warmup + 2 * args.steps + args.profile_steps + 4
    max_len = prefix + n_steps_total * (k + 1) + 32
    L = asd_b200.lib()

    t0 = time.time()
    target = QwenEngine(wl["target"], max_seqs=B, max_seq_len=max_len, max_tokens=max(256, B * (k + 1)),
                        tp_rank=rank if world > 1 else 0, tp_size=world, device=dev,
                        fuse_norm=not args.no_fuse_norm).load_random(seed=1)
    draft = QwenEngine(wl["draft"], max_seqs=B, max_seq_len=max_len, max_tokens=256, device=dev,
                       fuse_norm=not args.no_fuse_norm).load_random(seed=0)
    comm = None
    if world > 1:
        comm = NcclComm(rank, world)
        target.set_allreduce(comm.comm_ptr, comm.allreduce_fn_ptr)     # NCCL path (fallback / --opt p2p=0)
        if not args.nccl_only:
            target.enable_p2p()                                        # fused peer-memory all-reduce kernel
    for opt in args.opt or []:
        name, val = opt.split("=")
        target.set_option(name, int(val))
        draft.set_option(name, int(val))
    dec = SpecDecoder(target, draft, B, k, T, seed=4321)
    prompts = torch.randint(0, wl["target"].vocab_size, (B, prefix), generator=torch.Generator().manual_seed(1234))
    dec.prefill(prompts)
    torch.cuda.synchronize()
    setup_s = time.time() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wd.stage("warm-up + timed steps + e2e + profile", 180 if world > 1 else 0)
    for _ in range(args.warmup):
        dec.step()
    barrier()

    # ---- timed region: K steps, device-resident state, CUDA events; verify forward bracketed separately
    clocks = ClockSampler(local)
    clocks.start()
    L.asd_reset_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    emitted = torch.zeros((), dtype=torch.int64, device=dev)
    barrier()
    dec.time_verify, dec.verify_events = True, []
    ncu_range = os.environ.get("ASD_NCU_RANGE") == "1"   # `ncu --profile-from-start off`: capture the timed steps only
    if ncu_range:
        torch.cuda.profiler.start()
    ev[0].record()
    for _ in range(args.steps):
        out = dec.step()
        emitted += (out["accepted_len"].to(torch.int64) + 1).sum()
    ev[1].record()
    if ncu_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    dec.time_verify = False
    verify_ms_timed = sorted(a.elapsed_time(b) for a, b in dec.verify_events)
    verify_ms_timed = verify_ms_timed[len(verify_ms_timed) // 2]
    launches = int(L.asd_launch_count())
    ms_total = ev[0].elapsed_time(ev[1])
    tokens = int(emitted.item())

    # ---- e2e: same steps through the host-buffer entry (pinned H2D of the step inputs, D2H of the result)
    host_state = torch.empty(3, B, dtype=torch.int32).pin_memory()
    host_tokens = torch.empty(B, k + 1, dtype=torch.int32).pin_memory()
    host_acc = torch.empty(B, dtype=torch.int32).pin_memory()
    host_state.copy_(torch.stack([dec.last_tok, dec.prev_tok, dec.pos]).cpu())
    barrier()
    e2e_tokens = 0
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev2[0].record()
    for _ in range(args.steps):
        toks, acc = dec.step_host(host_state, host_tokens, host_acc)
        e2e_tokens += int((acc.long() + 1).sum())
    ev2[1].record()
    barrier()
    e2e_ms = ev2[0].elapsed_time(ev2[1])
    clocks.stop_flag.set()
    clocks.join(timeout=2)

    # ---- roofline of the dominant kernel (weight-streaming GEMMs of the verify forward), CUDA events
    for e in (target, draft):
        e.set_option("profile", 1)
    prof_t = {c: [0.0, 0] for c in target.PROFILE_CLASSES}
    prof_d = {c: [0.0, 0] for c in target.PROFILE_CLASSES}
    for _ in range(args.profile_steps):
        dec.step()
        for acc_d, e in ((prof_t, target), (prof_d, draft)):
            for c, (ms, n) in e.profile_read().items():
                acc_d[c][0] += ms
                acc_d[c][1] += n
    for e in (target, draft):
        e.set_option("profile", 0)
    ps = max(args.profile_steps, 1)

    # max over ranks
    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = t.tolist()

    line = None
    peak, peak_src = load_peaks()
    if rank == 0:
        tcfg, dcfg = wl["target"], wl["draft"]
        traffic = load_traffic()
        gemm_ms, gemm_n = prof_t["gemm"][0] / ps, prof_t["gemm"][1] / ps
        head_ms = prof_t["lm_head"][0] / ps
        layer_bytes = (tcfg.streamed_bytes() - 2 * tcfg.vocab_size * tcfg.hidden_size) / world
        head_bytes = 2 * tcfg.vocab_size * tcfg.hidden_size
        achieved = (layer_bytes + head_bytes) / ((gemm_ms + head_ms) * 1e-3) / 1e9 if gemm_ms > 0 else 0.0
        roof_kernel = ("gemm_ws_kernel (all weight-streaming GEMM launches of one "
                       f"{tcfg.name} verify forward, {int(gemm_n) + 1} launches)")
        roof_ms = gemm_ms + head_ms
        if world > 1:
            # tensor parallel: the row-parallel GEMMs finish their all-reduce inside the epilogue, so a launch
            # bracketed by events on one rank mostly measures how far the ranks drifted apart under profiling.
            # Use the in-stream verify forward (GEMMs + attention + exchange) instead: conservative.
            roof_ms = verify_ms_timed
            achieved = (layer_bytes + head_bytes) / (roof_ms * 1e-3) / 1e9
            roof_kernel = (f"gemm_ws_kernel with fused NVLink all-reduce: whole in-stream {tcfg.name} verify forward of "
                           f"one rank (weight shard {int((layer_bytes + head_bytes) / 1e6)} MB, attention included)")
        verify_ms = sum(v[0] for v in prof_t.values()) / ps
        draft_ms = sum(v[0] for v in prof_d.values()) / ps / max(k, 1)
        kv_bytes = B * (prefix + args.warmup * (k + 1)) * tcfg.kv_bytes_per_token() / world
        step_bytes = tcfg.streamed_bytes() / world + kv_bytes + k * dcfg.streamed_bytes()
        line = {
            "metric": METRIC, "value": tokens / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(wl, world, {"mean_tokens_per_seq_step": tokens / (args.steps * B),
                                              "setup_seconds": round(setup_s, 1)}),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_tokens / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": host_state.numel() * 4,
                    "d2h_bytes_per_step": (host_tokens.numel() + host_acc.numel() + host_state.numel()) * 4,
                    "api": "SpecDecoder.step_host (pinned host buffers in/out, sync per step)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": roof_kernel,
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src,
                         # measured dram__bytes_read + dram__bytes_write of the committed `ncu --set full` capture of
                         # these launches (profiles/traffic.json names the capture and the commit), per forward
                         "traffic": ((layer_bytes + head_bytes) * traffic["dram_bytes_per_algorithmic_byte"]
                                     if (traffic and world == 1 and args.workload == "32b") else None),
                         "traffic_source": traffic["source"] if traffic else None,
                         "bytes_per_forward": layer_bytes + head_bytes, "ms_per_forward": roof_ms},
            "verify_step_us": verify_ms_timed * 1e3, "draft_step_us": draft_ms * 1e3,
            "verify_step_hbm_frac": (tcfg.streamed_bytes() / world + kv_bytes) / (verify_ms_timed * 1e-3) / 1e9 / peak,
            "verify_step_us_profiled": verify_ms * 1e3,
            "verify_breakdown_ms": {c: round(v[0] / ps, 4) for c, v in prof_t.items()},
            "draft_breakdown_ms": {c: round(v[0] / ps / max(k, 1), 4) for c, v in prof_d.items()},
            "step_hbm_frac": step_bytes / (ms_total / args.steps * 1e-3) / 1e9 / peak,
        }
    tpc = None
    if world > 1:
        with wd.lock:
            wd.line = line if rank == 0 else {}     # from here on a stuck stage still yields the measured bench line
        wd.stage("tp_check", 90)
        tpc = tp_check(torch, dist, rank, world, dev, dec, target, wl["target"])
        if rank == 0:
            line["tp_check"] = tpc
    # ---- the main engines are done: free them before the companion measurements
    dec = None
    target.close()
    draft.close()
    target = draft = None
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if world > 1 and not args.no_config5:
        import json as _json
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = _json.load(f)
        wd.stage("config5", 200)
        c5 = config5(torch, dist, rank, world, dev, args, peaks)
        if rank == 0:
            line["config5"] = c5
    wd.stage("companions / output", 0)
    if rank == 0:
        if world == 1 and not args.no_companions:
            line["companions"] = companions(torch, dev, peak)
            line["stage_generate"] = stage_generate_record(torch, wl, dev)
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            from oracle.cpu_baseline import spec_step_baseline
            oracle.build()
            r = spec_step_baseline(tcfg, dcfg, B, k, prefix, T, sample_layers=1, threads=os.cpu_count())
            line["cpu_baseline"] = {"value": r["tokens_per_step"] / r["step_seconds"], "unit": UNIT,
                                    "cores": r["threads"], "kind": "port", "sample": r["sample"], "extrapolated": True}
        line["stage_seconds"] = wd.seconds()
        print(json.dumps(line), flush=True)
    if world > 1:
        wd.printed_rc = 3 if (tpc is not None and rank == 0 and not tpc.get("ok", False)) else 0
        wd.stage("teardown", 60)
    if comm is not None:
        barrier()
        comm.destroy()
    if world > 1:
        dist.destroy_process_group()
    if tpc is not None and rank == 0 and not tpc.get("ok", False):
        sys.stderr.write(f"tensor-parallel parity check FAILED: {tpc}\n")
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="32b", choices=["32b", "72b"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--prefix", type=int, default=512)
    ap.add_argument("--temperature", type=float, default=0.7)
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-companions", action="store_true", help="skip the stop-rule / sampler-grid / config-1 / Stage.generate records")
    ap.add_argument("--config5-only", action="store_true", help="N > 1: only the 72B verify-only record")
    ap.add_argument("--no-config5", action="store_true", help="N > 1: skip the 72B verify-only record (BASELINE configs[4])")
    ap.add_argument("--no-fuse-norm", action="store_true", help="separate add+RMSNorm kernels instead of the fused epilogues")
    ap.add_argument("--nccl-only", action="store_true", help="TP boundaries through ncclAllReduce instead of the fused kernel")
    ap.add_argument("--opt", action="append", help="engine option name=value (e.g. pdl=0, attn_impl=0)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
