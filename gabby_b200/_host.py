"""ctypes bindings of include/gabby_b200_host.h (the C view of the C++ host layer)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libgabby_host.so")

SYMBOLS = ["gb_last_error", "gb_rope_table"]

_LIB = None


class HostError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise HostError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        # libgabby_host.so links libb2l.so ($ORIGIN rpath)
        L = C.CDLL(LIB_PATH)
        L.gb_last_error.restype = C.c_char_p
        L.gb_rope_table.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p]
        _LIB = L
    return _LIB


def rope_table(arch, max_pos: int) -> np.ndarray:
    """(cos, sin) table from the host layer's RopeTable (gabby_b200/host/params.cc)."""
    out = np.empty((max_pos, arch.head_dim // 2, 2), dtype=np.float32)
    rc = lib().gb_rope_table(arch.rope_theta, 1, arch.rope_factor, arch.rope_low_freq_factor, arch.rope_high_freq_factor,
                             arch.rope_original_max_position, arch.head_dim, max_pos, out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise HostError(lib().gb_last_error().decode())
    return out
