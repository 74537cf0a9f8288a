"""ainmf -- B200-native NMF spectrogram inpainting (the main4_NMF / main4_NMF_gap / main4_NMF_mask hot path of
conniemessi/Audio-Inpainting), hand-written sm_100a CUDA behind a C ABI (include/ainmf.h, libainmf.so).

The directory is named `audio-inpainting_b200`; import it as `ainmf` (repo-root `ainmf.py` does the aliasing).
"""
from . import _capi, _lib  # noqa: F401
from ._capi import AinmfError  # noqa: F401

__all__ = ["ops", "inpainters", "NMFFairGapInpainter", "NMFFairInpainter", "SpectralInpainter", "AinmfError"]


def __getattr__(name):
    # torch is imported lazily so that `import ainmf` (e.g. to build) stays cheap
    if name in ("ops", "inpainters", "sharding"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    if name in ("NMFFairGapInpainter", "NMFFairInpainter", "SpectralInpainter"):
        from . import inpainters
        return getattr(inpainters, name)
    raise AttributeError(name)
