"""CPU: the C-ABI library loads and exports every symbol include/asd_b200.h declares."""
import ctypes

from asd_b200 import _lib


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/asd_b200.h but not exported"
    assert L.asd_abi_version() == 1


def test_last_error_is_a_string():
    L = _lib.lib()
    assert isinstance(L.asd_last_error(), bytes)
