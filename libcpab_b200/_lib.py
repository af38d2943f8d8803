"""ctypes binding of libcpab_b200.so (the C ABI declared in include/libcpab_b200.h).

There is no CPU fallback: if the CUDA library is missing and cannot be built, or a call fails,
this module raises.  (The reference silently degrades to a pure-python path when its extension
fails to compile, libcpab/pytorch/transformer.py:39-69; north_star forbids that here.)
"""
from __future__ import annotations

import ctypes
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

CPAB_F32, CPAB_F64 = 0, 1
CPAB_FLAG_FAST_MATH = 1
CPAB_FLAG_FAST_GRAD = 2
ABI_VERSION = 2

_vp, _i, _l, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_size_t
_ip = ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); must list every symbol include/libcpab_b200.h declares
SIGNATURES = {
    "cpab_b200_abi_version": (_i, []),
    "cpab_b200_last_error": (ctypes.c_char_p, []),
    "cpab_b200_build_info": (ctypes.c_char_p, []),
    "cpab_b200_set_tuning": (_i, [ctypes.c_char_p, _i]),
    "cpab_b200_launch_count": (ctypes.c_longlong, []),
    "cpab_b200_profile_enable": (_i, [_i]),
    "cpab_b200_profile_read": (_i, [_i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]),
    "cpab_b200_fp32_fma_probe": (_i, [_i, _i, _vp, _vp]),
    "cpab_b200_findcellidx": (_i, [_i, _i, _ip, _vp, _l, _vp, _vp]),
    "cpab_b200_theta_to_trels": (_i, [_i, _i, _ip, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "cpab_b200_expm": (_i, [_i, _i, _l, _vp, _vp, _vp]),
    "cpab_b200_forward": (_i, [_i, _i, _i, _ip, _i, _i, _l, _i, _vp, _vp, _vp, _vp]),
    "cpab_b200_backward_jacobian": (_i, [_i, _i, _ip, _i, _i, _i, _l, _i, _vp, _vp, _vp, _vp, _vp]),
    "cpab_b200_backward_workspace_bytes": (_sz, [_i, _i, _ip, _i, _l]),
    "cpab_b200_backward_theta": (_i, [_i, _i, _i, _ip, _i, _i, _i, _l, _i, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _vp, _sz, _vp]),
    "cpab_b200_backward_theta_diag": (_i, [_i, _i, _i, _ip, _i, _i, _i, _l, _i, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _vp, _sz, _vp, _vp]),
    "cpab_b200_rk2_cell_trace": (_i, [_i, _ip, _i, _i, _l, _i, _i, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "cpab_b200_forward_closed_form": (_i, [_i, _i, _ip, _i, _l, _i, _vp, _vp, _vp, _vp]),
    "cpab_b200_backward_theta_closed_form": (_i, [_i, _i, _ip, _i, _i, _l, _i, _vp, _vp, _vp, _vp, _vp,
                                                  _vp, _vp, _sz, _vp]),
    "cpab_b200_backward_theta_closed_form_from": (_i, [_i, _i, _ip, _i, _i, _l, _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                                       _vp, _vp, _sz, _vp]),
    "cpab_b200_closed_form_lane_stats": (_i, [_i, _i, _ip, _i, _l, _i, _vp, _vp, _vp, _vp, _vp]),
    "cpab_b200_interpolate_forward": (_i, [_i, _i, _i, _i, _ip, _ip, _vp, _vp, _vp, _vp]),
    "cpab_b200_transform_data_forward": (_i, [_i, _i, _i, _ip, _i, _i, _i, _ip, _ip, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cpab_b200_transform_data_backward": (_i, [_i, _i, _i, _ip, _i, _i, _i, _i, _ip, _ip, _vp, _vp, _vp, _vp, _vp,
                                               _vp, _vp, _vp, _sz, _vp]),
    "cpab_b200_interpolate_backward": (_i, [_i, _i, _i, _i, _ip, _ip, _vp, _vp, _vp, _vp, _vp,
                                            _vp]),
}


class CpabError(RuntimeError):
    """A libcpab_b200 call returned a negative status."""


def library_path() -> str:
    return _build.LIB


def load() -> ctypes.CDLL:
    """Load (building first if the shared object is absent or stale and nvcc is available)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("LIBCPAB_B200_SO")      # development: an experimental build of the same ABI
        if not path:
            path = _build.LIB
            if _build.needs_build():
                _build.build()      # raises if nvcc is missing: no silent fallback
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = ABI mismatch, surfaced loudly
            fn.restype = res
            fn.argtypes = args
        got = lib.cpab_b200_abi_version()
        if got != ABI_VERSION:
            raise CpabError(f"libcpab_b200.so ABI {got} != expected {ABI_VERSION}; rebuild")
        _lib = lib
        return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().cpab_b200_last_error().decode()
        raise CpabError(f"{what} failed with status {status}: {msg}")


def nc_array(nc):
    return (ctypes.c_int * len(nc))(*[int(v) for v in nc])


def set_tuning(key: str, value: int) -> None:
    check(load().cpab_b200_set_tuning(key.encode(), int(value)), "set_tuning")


PROFILE_SLOTS = {"forward": 0, "backward": 1, "interp_fwd": 2, "interp_bwd": 3,
                 "theta_to_trels": 4, "epilogue": 5, "backward_redo": 6}


def launch_count() -> int:
    return int(load().cpab_b200_launch_count())


def profile_enable(on: bool) -> None:
    check(load().cpab_b200_profile_enable(1 if on else 0), "profile_enable")


def profile_read(slot: str):
    """(total_ms, launches) accumulated for one kernel slot since profile_enable(True)."""
    ms, n = ctypes.c_double(0.0), ctypes.c_longlong(0)
    check(load().cpab_b200_profile_read(PROFILE_SLOTS[slot], ctypes.byref(ms), ctypes.byref(n)),
          "profile_read")
    return ms.value, n.value
