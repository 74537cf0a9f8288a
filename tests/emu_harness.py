"""TEST HARNESS: drives tests/_emu/libainmf_emu.so (the kernels of audio-inpainting_b200/csrc compiled for the
host against csrc/emu/cuda_emu.h) with numpy buffers, through the same C ABI the CUDA library exports.
It lets `pytest -m "not gpu"` check kernel indexing/logic against the oracle without a GPU.  It is never
imported by the product package."""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "audio-inpainting_b200")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


capi = _load("ainmf_capi_for_tests", os.path.join(PKG, "_capi.py"))
_build = _load("ainmf_build_for_tests", os.path.join(PKG, "build.py"))
_lib = None
_handle = None


def lib():
    global _lib, _handle
    if _lib is None:
        so = _build.build_emulator()
        _lib = capi.bind(C.CDLL(so))
        h = C.c_void_p()
        rc = _lib.ainmf_create(C.byref(h), 0)
        assert rc == 0, _lib.ainmf_last_error(None)
        _handle = h
    return _lib


def handle():
    lib()
    return _handle


def check(rc):
    if rc != 0:
        raise capi.AinmfError(rc, (lib().ainmf_last_error(handle()) or b"").decode())


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def stft(x, n_fft, hop):
    x = np.ascontiguousarray(x, np.float32)
    if x.ndim == 1:
        x = x[None]
    B, N = x.shape
    T, F, _ = capi.stft_geometry(lib(), N, n_fft, hop)
    mag = np.empty((B, F, T), np.float32)
    Z = np.empty((B, F, T), np.complex64)
    check(lib().ainmf_stft(handle(), ptr(x), B, N, n_fft, hop, ptr(mag), ptr(Z), None))
    return mag, Z


def istft(Z, n_fft, hop, N):
    Z = np.ascontiguousarray(Z, np.complex64)
    if Z.ndim == 2:
        Z = Z[None]
    B, F, T = Z.shape
    y = np.empty((B, N), np.float32)
    check(lib().ainmf_istft(handle(), ptr(Z), B, T, n_fft, hop, N, ptr(y), None))
    return y


def gap_mask(x, hop, T, thr, num, den):
    x = np.ascontiguousarray(x, np.float32)
    if x.ndim == 1:
        x = x[None]
    B, N = x.shape
    bad = np.zeros((B, T), np.uint8)
    idx = np.full((B, T), -1, np.int32)
    nb = np.zeros(B, np.int32)
    check(lib().ainmf_gap_mask(handle(), ptr(x), B, N, hop, T, thr, num, den, ptr(bad), ptr(idx), ptr(nb), None))
    return bad, idx, nb


def nmf_fit(X, K, max_iter=200, tol=1e-4, seed=0, W0=None, H0=None, solver=0):
    X = np.ascontiguousarray(X, np.float32)
    if X.ndim == 2:
        X = X[None]
    B, F, T = X.shape
    W = np.empty((B, F, K), np.float32)
    H = np.empty((B, K, T), np.float32)
    err = np.empty(B, np.float32)
    nit = np.empty(B, np.int32)
    if W0 is not None:
        W0 = np.ascontiguousarray(np.broadcast_to(W0, (B, F, K)), np.float32)
        H0 = np.ascontiguousarray(np.broadcast_to(H0, (B, K, T)), np.float32)
    check(lib().ainmf_nmf_fit(handle(), ptr(X), B, F, T, K, max_iter, tol, solver, seed, ptr(W0), ptr(H0),
                              ptr(W), ptr(H), ptr(err), ptr(nit), None))
    return W, H, err, nit


def inpaint(x, **kw):
    x = np.ascontiguousarray(x, np.float32)
    if x.ndim == 1:
        x = x[None]
    B, N = x.shape
    p = capi.default_params(lib(), batch=B, n_samples=N, **kw)
    T, F, _ = capi.stft_geometry(lib(), N, p.n_fft, p.hop)
    K = p.rank
    ws_bytes = lib().ainmf_workspace_bytes(handle(), C.byref(p))
    if ws_bytes == 0:
        check(-1)
    ws = np.zeros(ws_bytes + 256, np.uint8)
    off = (-ws.ctypes.data) % 256
    y = np.empty((B, N), np.float32)
    idx = np.full((B, T), -1, np.int32)
    nb = np.zeros(B, np.int32)
    W = np.zeros((B, F, K), np.float32)
    H = np.zeros((B, K, T), np.float32)
    err = np.zeros(B, np.float32)
    nit = np.zeros(B, np.int32)
    check(lib().ainmf_inpaint(handle(), C.byref(p), ptr(x), None, None, ptr(y), ptr(idx), ptr(nb), ptr(W), ptr(H),
                              ptr(err), ptr(nit), C.c_void_p(ws.ctypes.data + off), ws_bytes, None))
    return dict(y=y, bad_idx=idx, n_bad=nb, W=W, H=H, err=err, n_iter=nit)


# ---- callers / baselines either side of the NMF path (gaps.cu) -------------------------------------------------
def find_main_gap(x, thr):
    x = np.ascontiguousarray(x, np.float32)
    span = np.zeros((x.shape[0], 2), np.int64)
    check(lib().ainmf_find_main_gap(handle(), ptr(x), x.shape[0], x.shape[1], thr, ptr(span), None))
    return span


def find_gaps(x, thr, min_len, max_runs):
    x = np.ascontiguousarray(x, np.float32)
    runs = np.full((x.shape[0], max_runs, 2), -1, np.int64)
    n = np.zeros(x.shape[0], np.int32)
    check(lib().ainmf_find_gaps(handle(), ptr(x), x.shape[0], x.shape[1], thr, min_len, ptr(runs), max_runs, ptr(n), None))
    return runs, n


def linear_interp(x, thr):
    x = np.ascontiguousarray(x, np.float32)
    y = np.zeros_like(x)
    nd = np.zeros(x.shape[0], np.int64)
    check(lib().ainmf_linear_interp(handle(), ptr(x), x.shape[0], x.shape[1], thr, ptr(y), ptr(nd), None))
    return y, nd


def blend_boundaries(raw, restored, gs, ge, blend_len):
    raw = np.ascontiguousarray(raw, np.float32); restored = np.ascontiguousarray(restored, np.float32)
    out = np.zeros_like(raw)
    check(lib().ainmf_blend_boundaries(handle(), ptr(raw), ptr(restored), raw.size, gs, ge, blend_len, ptr(out), None))
    return out


def snr_db(ref, est, begin, end):
    ref = np.ascontiguousarray(ref, np.float32); est = np.ascontiguousarray(est, np.float32)
    out = C.c_double()
    check(lib().ainmf_snr_db(handle(), ptr(ref), ptr(est), begin, end, C.byref(out), None))
    return out.value


def apply_gaps(x, starts, lens):
    x = np.ascontiguousarray(x, np.float32).copy()
    starts = np.ascontiguousarray(starts, np.int64); lens = np.ascontiguousarray(lens, np.int64)
    check(lib().ainmf_apply_gaps(handle(), ptr(x), x.shape[0], x.shape[1], ptr(starts), ptr(lens), starts.shape[1], None))
    return x
