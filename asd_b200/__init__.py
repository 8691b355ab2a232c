"""Import shim: the package sources live in ``adaptive-speculative-decoding_b200/`` (a directory
name Python cannot import directly), so this package extends its search path to that directory.
``import asd_b200.serving.pipeline`` etc. resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "adaptive-speculative-decoding_b200")
__path__.insert(0, _real)

from ._lib import lib, library_path  # noqa: E402,F401

__all__ = ["lib", "library_path"]


def install_as_src():
    """Register this package under the reference's module names (``src.serving.pipeline``,
    ``src.models.stage``, ``src.models.predictor``, ``src.algorithms.dp_solver``, ...) so that code written
    against sa2shun/adaptive-speculative-decoding imports the B200 implementation unchanged."""
    import importlib
    import sys
    import types
    names = ["algorithms", "algorithms.dp_solver", "algorithms.optimizer", "models", "models.stage", "models.predictor", "serving",
             "serving.pipeline", "serving.cache_manager", "serving.real_model_pipeline", "serving.server", "theory",
             "theory.optimal_stopping"]
    root = sys.modules.setdefault("src", types.ModuleType("src"))
    root.__path__ = []
    for n in names:
        mod = importlib.import_module(f"asd_b200.{n}")
        sys.modules[f"src.{n}"] = mod
        parent, _, leaf = n.rpartition(".")
        setattr(sys.modules[f"src.{parent}"] if parent else root, leaf, mod)
    return root
