"""bench.py's static contract (no GPU): workload naming, parallelism strings, CLI defaults."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Args:
    workload, batch, k, prefix, temperature = "32b", 16, 5, 512, 0.7


def test_config_names_the_baseline_workload_and_the_exchange_path():
    b = _bench()
    wl = b.workload(_Args)
    assert wl["target"].name.lower().endswith("32b") and wl["draft"].name.lower().endswith("7b")
    c1 = b.config_dict(wl, 1)
    assert "workload" in c1 and "model" not in c1                      # tier contract: workload key, no model keys
    assert "k=5" in c1["workload"] and "batch 16" in c1["workload"] and "prefix 512" in c1["workload"]
    assert c1["parallelism"] == "single-gpu" and "L2" in c1["l2"] or "l2" in c1
    assert "fused into the row-parallel GEMM" in b.config_dict(wl, 2)["parallelism"]
    for n in (4, 8):
        assert "peer-memory all-reduce" in b.config_dict(wl, n)["parallelism"]
    assert b.METRIC == "accepted_tokens_per_second" and b.UNIT == "tok/s"


def test_cli_defaults_finish_in_minutes_and_enforce_three_warmups(monkeypatch):
    b = _bench()
    seen = {}
    monkeypatch.setattr(b, "run_ours", lambda a: seen.update(vars(a)))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--warmup", "1"])
    b.main()
    assert seen["gpus"] == 1 and seen["steps"] == 24 and seen["warmup"] == 3 and seen["impl"] == "ours"
    monkeypatch.setattr(b, "run_reference", lambda a: seen.update(impl_called="reference"))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--gpus", "2"])
    b.main()
    assert seen["impl_called"] == "reference"
