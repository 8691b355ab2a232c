"""Load pieces of the reference by file path (ORACLE tooling; container-only).

/root/reference cannot be imported as a package (``import src`` raises TypeError at
src/core/interfaces.py:466) and does not exist on the GPU box, so
this module is used only by ``oracle/gen_golden.py`` and by CPU tests that skip when
the reference tree is absent.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import time
import types

REF_ROOT = os.environ.get("ASD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "algorithms", "dp_solver.py"))


def load_by_path(name: str, relpath: str, package: str | None = None):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    if package is not None:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def dp_solver():
    return load_by_path("refsrc_dp_solver", "src/algorithms/dp_solver.py")


def optimal_stopping():
    return load_by_path("refsrc_optimal_stopping", "src/theory/optimal_stopping.py")


def pipeline_module(stage_module, predictor_module):
    """The reference's UNMODIFIED src/serving/pipeline.py, importable through a synthetic
    parent package whose ``models.stage`` / ``models.predictor`` are the given modules
    (SURVEY.md Appendix F).  ``start_time`` (pipeline.py:269 reads an undefined global)
    is injected so a request can complete."""
    for pkg in ("refsrc", "refsrc.models", "refsrc.algorithms", "refsrc.serving"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    load_by_path("refsrc.algorithms.dp_solver", "src/algorithms/dp_solver.py")
    load_by_path("refsrc.serving.cache_manager", "src/serving/cache_manager.py")
    sys.modules["refsrc.models.stage"] = stage_module
    sys.modules["refsrc.models.predictor"] = predictor_module
    mod = load_by_path("refsrc.serving.pipeline", "src/serving/pipeline.py", package="refsrc.serving")
    mod.start_time = time.time()
    return mod
