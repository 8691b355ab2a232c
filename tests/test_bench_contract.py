"""bench.py's static contract (no GPU): workload naming, parallelism strings, CLI defaults."""
import importlib.util
import json
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


def test_watchdog_prints_what_it_has_and_leaves():
    """N > 1: a stage that overruns its budget must not hang the job - rank 0 prints the finished bench line (or an error
    record when the timed region never finished) and the process exits (bench.Watchdog)."""
    import subprocess
    import sys
    code = """
import sys, time
sys.path.insert(0, %r)
import bench
wd = bench.Watchdog(0, 8)
wd.start()
if sys.argv[1] == "line":
    wd.line = {"metric": "m", "value": 1.0}
wd.stage(sys.argv[2], 1)
time.sleep(20)
print("NOT REACHED")
""" % ROOT
    for have, stage, rc, key in (("none", "setup + prefill", 5, "error"), ("line", "config5", 0, "config5"),
                                 ("line", "tp_check", 3, "tp_check")):
        r = subprocess.run([sys.executable, "-c", code, have, stage], capture_output=True, text=True, timeout=60)
        assert r.returncode == rc, (have, stage, r.returncode, r.stderr[-300:])
        assert "NOT REACHED" not in r.stdout
        rec = json.loads(r.stdout.strip().splitlines()[-1])
        assert key in rec and ("watchdog" in rec or "watchdog" in rec.get("error", ""))
