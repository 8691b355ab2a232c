"""Import shim: the package sources live in ``adaptive-speculative-decoding_b200/`` (a directory
name Python cannot import directly), so this package extends its search path to that directory.
``import asd_b200.serving.pipeline`` etc. resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "adaptive-speculative-decoding_b200")
__path__.insert(0, _real)

from ._lib import lib, library_path  # noqa: E402,F401

__all__ = ["lib", "library_path"]
