"""Loader of libainmf.so -- the only native library the product path uses.

There is no CPU path: if the library is missing or no sm_100 device is present, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libainmf.so")
_lock = threading.Lock()
_lib = None
_handles: dict[int, C.c_void_p] = {}


def lib() -> C.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} not found: build it with `python audio-inpainting_b200/build.py` "
                    "(nvcc, sm_100a). ainmf has no CPU fallback.")
            _lib = _capi.bind(C.CDLL(LIB_PATH))
        return _lib


def handle(device: int) -> C.c_void_p:
    """One ainmf handle per CUDA device, created on first use."""
    L = lib()
    with _lock:
        h = _handles.get(device)
        if h is None:
            h = C.c_void_p()
            rc = L.ainmf_create(C.byref(h), int(device))
            if rc != 0:
                raise _capi.AinmfError(rc, (L.ainmf_last_error(None) or b"").decode())
            _handles[device] = h
        return h


def check(rc: int, device: int) -> None:
    if rc != 0:
        raise _capi.AinmfError(rc, (lib().ainmf_last_error(_handles.get(device)) or b"").decode())
