"""TEST HARNESS: loader of audio-inpainting_b200/libainmf_diag.so (descriptor probe, sweep unit kernel, tcgen05
issue-rate microbenchmark; csrc/diag/).  Diagnostics only -- the product library does not contain or export them."""
from __future__ import annotations

import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIAG_PATH = os.path.join(ROOT, "audio-inpainting_b200", "libainmf_diag.so")
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(DIAG_PATH):
            raise RuntimeError(f"{DIAG_PATH} not found: python audio-inpainting_b200/build.py")
        _lib = C.CDLL(DIAG_PATH)
    return _lib
