"""Import the UNMODIFIED reference scripts -- TEST INFRASTRUCTURE ONLY, build container only.

``/root/reference`` does not exist on the GPU box, so nothing that runs there may import this
module; it is used by ``tests/golden/make_golden.py`` (fixture generation) and by the
``not gpu`` tests that are skipped when the reference tree is absent.

The three hot-path scripts import matplotlib at line 2 (absent here) and write under a
cwd-relative ``demo_assets/``; two of them run their pipeline at import time
(main4_NMF_gap.py:86-89, main4_NMF_mask.py:92-95).  They are therefore loaded with a stub
``matplotlib`` and a scratch working directory that holds copies of their inputs.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import shutil
import sys
import tempfile
import types

REF = os.environ.get("AINMF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "main4_NMF_gap.py"))


def _stub_matplotlib():
    if "matplotlib.pyplot" in sys.modules and not getattr(sys.modules["matplotlib.pyplot"], "_ainmf_stub", False):
        return
    plt = types.ModuleType("matplotlib.pyplot")
    plt._ainmf_stub = True
    for n in ("figure specgram axis tight_layout savefig close subplot plot axvspan legend title "
              "pcolormesh axvline ylabel xlabel show").split():
        setattr(plt, n, lambda *a, **k: None)
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


@contextlib.contextmanager
def scratch_cwd(files: dict[str, str] | None = None):
    """cd into a temp dir; `files` maps relative destination -> absolute source (copied)."""
    old = os.getcwd()
    d = tempfile.mkdtemp(prefix="ainmf_ref_")
    try:
        for rel, src in (files or {}).items():
            dst = os.path.join(d, rel)
            os.makedirs(os.path.dirname(dst) or d, exist_ok=True)
            shutil.copyfile(src, dst)
        os.chdir(d)
        yield d
    finally:
        os.chdir(old)
        shutil.rmtree(d, ignore_errors=True)


def load(name: str, quiet: bool = True):
    """exec a reference script as a module (import side effects included)."""
    _stub_matplotlib()
    spec = importlib.util.spec_from_file_location("ainmf_ref_" + name, os.path.join(REF, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    out = io.StringIO()
    with contextlib.redirect_stdout(out if quiet else sys.stdout):
        spec.loader.exec_module(m)
    return m
