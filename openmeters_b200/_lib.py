"""Loads the CUDA extension (libomb200.so, built in-tree by openmeters_b200.build).

There is no CPU fallback: if the shared library is missing this raises, and
every compute entry point inside the library fails with OMB_ERR_CUDA when no
sm_100 device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

from . import _capi as capi

HERE = os.path.dirname(os.path.abspath(__file__))
# OMB_LIB: measurement knob for A/B builds of the same sources (tools/ab_build.py); never set in production.
LIB_PATH = os.environ.get("OMB_LIB") or os.path.join(HERE, "libomb200.so")


class ExtensionMissing(ImportError):
    pass


@lru_cache(maxsize=1)
def cdll() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ExtensionMissing(
            f"{LIB_PATH} not found: build it with `python -m openmeters_b200.build` "
            "(nvcc, sm_100a). openmeters_b200 has no CPU fallback.")
    return C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)


@lru_cache(maxsize=1)
def api():
    return capi.bind(cdll(), "omb_")
