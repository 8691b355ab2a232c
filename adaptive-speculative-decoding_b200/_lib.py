"""ctypes binding of libasd_b200.so (the C ABI declared in include/asd_b200.h).

There is no CPU fallback: if the shared library is missing this raises, and the CUDA entry
points raise ``AsdError`` when the device or the kernel launch fails."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class AsdError(RuntimeError):
    pass


class ModelConfigC(ctypes.Structure):
    """mirror of ``asd_model_config`` (include/asd_b200.h)"""
    _fields_ = [("hidden", ctypes.c_int), ("n_layers", ctypes.c_int), ("n_heads", ctypes.c_int),
                ("n_kv_heads", ctypes.c_int), ("head_dim", ctypes.c_int), ("ffn", ctypes.c_int),
                ("vocab", ctypes.c_int), ("rms_eps", ctypes.c_float), ("max_tokens", ctypes.c_int),
                ("page_size", ctypes.c_int), ("tp_rank", ctypes.c_int), ("tp_size", ctypes.c_int)]


def library_path() -> str:
    return os.path.join(_HERE, "libasd_b200.so")


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise AsdError(f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback for the CUDA path)")
    L = ctypes.CDLL(path)
    c = ctypes
    vp, i32, f32, f64, sz = c.c_void_p, c.c_int, c.c_float, c.c_double, c.c_size_t
    sigs = {
        "asd_abi_version": (i32, []),
        "asd_last_error": (c.c_char_p, []),
        "asd_launch_count": (c.c_longlong, []),
        "asd_reset_launch_count": (None, []),
        "asd_reject_sample_workspace_bytes": (sz, [i32, i32]),
        "asd_reject_sample_set_impl": (None, [i32]),
        "asd_reject_sample": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp, vp]),
        "asd_reject_sample_host": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, f32, vp, vp, vp, vp, vp]),
        "asd_stop_rule": (i32, [vp, vp, i32, i32, f64, i32, f64, f64, vp, vp, vp]),
        "asd_stop_rule_host": (i32, [vp, vp, i32, f64, i32, f64, f64, vp]),
        "asd_stop_rule_rows": (i32, [vp, vp, vp, i32, i32, i32, f64, f64, vp, vp, vp]),
        "asd_stop_rule_rows_host": (i32, [vp, vp, vp, i32, i32, i32, f64, f64, vp, vp]),
        "asd_cascade_decide": (i32, [vp, vp, i32, i32, vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, i32, f64, i32, f64, f64,
                                     f64, vp, vp, vp, vp]),
        "asd_engine_persist_trace": (i32, [vp, vp]),
        "asd_engine_peer_connect": (i32, [vp, i32]),
        "asd_bayesian_adjustment_host": (f64, [f64, f64, f64, f64]),
        "asd_linear_bf16": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
        "asd_linear_bf16_tc": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
        "asd_linear_plan": (i32, [i32, i32, i32, i32, vp, vp, vp]),
        "asd_engine_create": (vp, [vp]),
        "asd_engine_destroy": (None, [vp]),
        "asd_engine_set_layer": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp]),
        "asd_engine_set_globals": (i32, [vp, vp, vp, vp, vp]),
        "asd_engine_kv_pool_bytes": (sz, [vp, i32]),
        "asd_engine_set_kv": (i32, [vp, vp, i32, vp, i32, i32]),
        "asd_engine_set_allreduce": (i32, [vp, vp, vp]),
        "asd_engine_set_option": (i32, [vp, c.c_char_p, i32]),
        "asd_engine_ipc_export": (i32, [vp, vp]),
        "asd_engine_ipc_import": (i32, [vp, vp]),
        "asd_engine_tp_error": (i32, [vp]),
        "asd_engine_profile_read": (i32, [vp, vp, vp, i32]),
        "asd_debug_gemm_trace": (i32, [vp, i32]),
        "asd_debug_attn_trace": (i32, [vp, i32]),
        "asd_engine_forward": (i32, [vp, vp, vp, vp, i32, vp, vp, i32, i32, i32, vp, i32, vp, c.c_longlong, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().asd_last_error().decode("utf-8", "replace")
        raise AsdError(f"{what}: {msg}" if what else msg)


def declared_symbols() -> list[str]:
    """every asd_* function declared in include/asd_b200.h (used by the ABI test)"""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "asd_b200.h")
    with open(hdr) as f:
        text = f.read()
    return sorted(set(re.findall(r"ASD_API[^;(]*?\b(asd_[a-z0-9_]+)\s*\(", text)))
