"""float(f"{p:.3e}") for arrays through the library's record writer (tg_record.cpp): the decimal round trip improved-policy
values take through the SGF comment (sgf/selfplay_record.py:61 -> nn/feature.py:96)."""
import ctypes as C

import numpy as np

from . import _lib


def round_policy_like_sgf(policy):
    lib = _lib.load()
    a = np.ascontiguousarray(policy, dtype=np.float64).copy()
    lib.tg_round_policy(a.ctypes.data_as(C.POINTER(C.c_double)), a.size)
    return a
