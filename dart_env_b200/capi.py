"""ctypes binding of libdartb.so (include/dartb.h).  The library is the product path: if it
is missing or no CUDA device is present, calls fail loudly — there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build
from .cstructs import CModel, CTask

_LIB = None

_EXPORTS = [
    "dartb_create", "dartb_create_f64", "dartb_destroy", "dartb_set_option", "dartb_reset", "dartb_set_state",
    "dartb_get_state", "dartb_set_state_f64", "dartb_get_state_f64", "dartb_step", "dartb_substep",
    "dartb_substep_f64", "dartb_get_contacts", "dartb_get_truncated", "dartb_max_contacts", "dartb_num_worlds",
    "dartb_num_dofs", "dartb_is_f64", "dartb_launch_count", "dartb_kernel_name", "dartb_last_error",
    "dartb_version", "dartb_describe", "dartb_step_host", "dartb_step_host_gym", "dartb_seed", "dartb_seed_worlds", "dartb_register_host", "dartb_unregister_host", "dartb_set_aux", "dartb_get_aux", "dartb_set_obs_peers",
    "dartb_set_body_params", "dartb_get_body_table",
]


class DartbError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.SO


def load(build_if_missing: bool = True):
    """Load libdartb.so (building it with nvcc when absent/stale and nvcc is available)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    so = _build.SO
    if build_if_missing and not (os.environ.get("DARTB_NO_REBUILD") and os.path.exists(so)):
        try:
            if _build.is_stale():
                _build.build()
        except Exception as exc:  # no nvcc on this machine: use the shipped .so if there is one
            if not os.path.exists(so):
                raise DartbError("libdartb.so is not built and nvcc failed: %s" % exc)
    if not os.path.exists(so):
        raise DartbError("libdartb.so not found at %s (run python -m dart_env_b200.build)" % so)
    L = C.CDLL(so)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    sig = {
        "dartb_create": (C.c_int, [C.POINTER(CModel), C.POINTER(CTask), i32, i32, u64, i64, C.POINTER(vp)]),
        "dartb_create_f64": (C.c_int, [C.POINTER(CModel), C.POINTER(CTask), i32, i32, u64, i64, C.POINTER(vp)]),
        "dartb_destroy": (C.c_int, [vp]),
        "dartb_set_option": (C.c_int, [vp, i32, dbl]),
        "dartb_seed": (C.c_int, [vp, u64]),
        "dartb_seed_worlds": (C.c_int, [vp, vp]),
        "dartb_register_host": (C.c_int, [vp, vp, C.c_size_t]),
        "dartb_unregister_host": (C.c_int, [vp, vp]),
        "dartb_set_aux": (C.c_int, [vp, vp, vp]),
        "dartb_set_body_params": (C.c_int, [vp, vp, vp]),
        "dartb_get_body_table": (C.c_int, [vp, vp, vp]),
        "dartb_set_obs_peers": (C.c_int, [vp, C.POINTER(vp), i32, i64]),
        "dartb_get_aux": (C.c_int, [vp, vp, vp]),
        "dartb_reset": (C.c_int, [vp, vp, vp, vp]),
        "dartb_set_state": (C.c_int, [vp, vp, vp, vp]),
        "dartb_get_state": (C.c_int, [vp, vp, vp, vp]),
        "dartb_set_state_f64": (C.c_int, [vp, vp, vp, vp]),
        "dartb_get_state_f64": (C.c_int, [vp, vp, vp, vp]),
        "dartb_step": (C.c_int, [vp, vp, vp, vp, vp, i32, vp]),
        "dartb_substep": (C.c_int, [vp, vp, vp, vp]),
        "dartb_step_host": (C.c_int, [vp, vp, vp, vp, vp, i32, vp]),
        "dartb_step_host_gym": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, vp]),
        "dartb_substep_f64": (C.c_int, [vp, vp, vp, vp]),
        "dartb_get_contacts": (C.c_int, [vp, vp, vp, vp, vp]),
        "dartb_get_truncated": (C.c_int, [vp, vp, vp]),
        "dartb_max_contacts": (i32, [vp]),
        "dartb_num_worlds": (i32, [vp]),
        "dartb_num_dofs": (i32, [vp]),
        "dartb_is_f64": (i32, [vp]),
        "dartb_launch_count": (i64, [vp]),
        "dartb_kernel_name": (C.c_char_p, [vp]),
        "dartb_last_error": (C.c_char_p, []),
        "dartb_version": (C.c_char_p, []),
        "dartb_describe": (C.c_int, [C.POINTER(CModel), C.POINTER(CTask), C.c_char_p, i32]),
    }
    for name in _EXPORTS:
        fn = getattr(L, name)
        fn.restype, fn.argtypes = sig[name]
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise DartbError(load().dartb_last_error().decode("utf-8", "replace"))


def describe(model, task) -> str:
    """Which kernel specialisation a Model/Task lowers to (host-only, no GPU needed)."""
    from .cstructs import pack_model, pack_task
    L = load()
    cm, ct = pack_model(model), pack_task(task)
    buf = C.create_string_buffer(256)
    check(L.dartb_describe(C.byref(cm), C.byref(ct), buf, 256))
    return buf.value.decode()
