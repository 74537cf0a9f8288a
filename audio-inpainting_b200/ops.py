"""PyTorch custom ops `torch.ops.ainmf.*` over the C ABI of libainmf.so.

CUDA tensors only (no CPU kernels are registered: calling an op with CPU tensors raises NotImplementedError from
the dispatcher).  Work is enqueued on the current CUDA stream of the tensors' device; no autograd.

Shapes follow the reference's arrays: spectrogram-like tensors are (B, F, T) as returned by
`scipy.signal.stft` / consumed by `sklearn.decomposition.NMF` (main4_NMF_gap.py:47-64).
"""
from __future__ import annotations

import ctypes as C
import functools
import threading

import torch

from . import _capi, _lib

_LIBRARY = torch.library.Library("ainmf", "DEF")
_LIBRARY.define("stft(Tensor x, int n_fft, int hop) -> (Tensor, Tensor)")
_LIBRARY.define("gap_mask(Tensor x, int hop, int n_frames, float threshold, int frac_num, int frac_den) "
                "-> (Tensor, Tensor, Tensor)")
_LIBRARY.define("nmf_fit(Tensor X, int rank, int max_iter, float tol, int seed, Tensor? W0, Tensor? H0, "
                "str solver='cd') -> (Tensor, Tensor, Tensor, Tensor)")
_LIBRARY.define("istft(Tensor Z, int n_fft, int hop, int length) -> Tensor")
_LIBRARY.define("nmf_inpaint(Tensor x, int n_fft, int hop, int rank, int max_iter, float tol, int seed, "
                "float threshold, int frac_num, int frac_den, int col_start, int col_end, int n_outer, "
                "Tensor? W0, Tensor? H0, str solver='cd') -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)")
_LIBRARY.define("load_pcm16(Tensor pcm) -> (Tensor, Tensor)")
_LIBRARY.define("store_pcm16(Tensor y) -> Tensor")
_LIBRARY.define("find_main_gap(Tensor x, float threshold) -> Tensor")
_LIBRARY.define("find_gaps(Tensor x, float threshold, int min_len, int max_runs) -> (Tensor, Tensor)")
_LIBRARY.define("linear_interp(Tensor x, float threshold) -> (Tensor, Tensor)")
_LIBRARY.define("blend_boundaries(Tensor raw, Tensor restored, int gap_start, int gap_end, int blend_len) -> Tensor")


# One handle per device owns shared scratch, pinned staging and the cached tables, and ctypes releases the GIL during a call:
# two Python threads entering the library on the same device would share (and could free) that scratch.  Every op therefore
# takes the device's lock around its C call.  Work of one device is also stream-ordered on ONE stream at a time: callers that
# switch CUDA streams between ops must order those streams themselves (documented in include/ainmf.h and INTEGRATION.md).
_locks: dict[int, threading.RLock] = {}
_locks_guard = threading.Lock()


def _lock(dev: int) -> threading.RLock:
    with _locks_guard:
        lk = _locks.get(dev)
        if lk is None:
            lk = _locks[dev] = threading.RLock()
        return lk


def _serialised(fn):
    """Run `fn` holding the lock of the device of its first tensor argument."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        t = next((a for a in args if isinstance(a, torch.Tensor)), None)
        if t is None or not t.is_cuda:
            return fn(*args, **kw)
        dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
        with _lock(dev):
            return fn(*args, **kw)
    return wrapper


def _dev(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise RuntimeError("ainmf ops take CUDA tensors only (there is no CPU implementation)")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(dev: int) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _stft(x, n_fft, hop):
    x = _f32(x, "x")
    if x.dim() != 2:
        raise RuntimeError("x must be [B, N]")
    dev = _dev(x)
    B, N = x.shape
    L = _lib.lib()
    with torch.cuda.device(dev):
        T, F, _ = _capi.stft_geometry(L, N, n_fft, hop)
        mag = torch.empty((B, F, T), dtype=torch.float32, device=x.device)
        Z = torch.empty((B, F, T), dtype=torch.complex64, device=x.device)
        _lib.check(L.ainmf_stft(_lib.handle(dev), _p(x), B, N, n_fft, hop, _p(mag), _p(Z), _stream(dev)), dev)
    return mag, Z


def _gap_mask(x, hop, n_frames, threshold, frac_num, frac_den):
    x = _f32(x, "x")
    if x.dim() != 2:
        raise RuntimeError("x must be [B, N]")
    dev = _dev(x)
    B, N = x.shape
    L = _lib.lib()
    with torch.cuda.device(dev):
        bad = torch.empty((B, n_frames), dtype=torch.uint8, device=x.device)
        idx = torch.full((B, n_frames), -1, dtype=torch.int32, device=x.device)
        nb = torch.empty((B,), dtype=torch.int32, device=x.device)
        _lib.check(L.ainmf_gap_mask(_lib.handle(dev), _p(x), B, N, hop, n_frames, threshold, frac_num, frac_den,
                                    _p(bad), _p(idx), _p(nb), _stream(dev)), dev)
    return bad, idx, nb


def _solver_id(solver: str) -> int:
    try:
        return {"cd": _capi.SOLVER_CD, "mu": _capi.SOLVER_MU, "mu-kl": _capi.SOLVER_MU_KL}[solver]
    except KeyError:
        raise RuntimeError(f"unknown solver {solver!r}: 'cd' (sklearn coordinate descent, what the reference runs) or "
                           "'mu' (multiplicative update, Frobenius), 'mu-kl' (multiplicative update, Kullback-Leibler)") from None


def _nmf_fit(X, rank, max_iter, tol, seed, W0, H0, solver="cd"):
    X = _f32(X, "X")
    if X.dim() != 3:
        raise RuntimeError("X must be [B, F, T]")
    dev = _dev(X)
    B, F, T = X.shape
    if (W0 is None) != (H0 is None):
        raise RuntimeError("W0 and H0 must be given together")
    if W0 is not None:
        W0 = _f32(W0, "W0").expand(B, F, rank).contiguous()
        H0 = _f32(H0, "H0").expand(B, rank, T).contiguous()
    L = _lib.lib()
    with torch.cuda.device(dev):
        W = torch.empty((B, F, rank), dtype=torch.float32, device=X.device)
        H = torch.empty((B, rank, T), dtype=torch.float32, device=X.device)
        err = torch.empty((B,), dtype=torch.float32, device=X.device)
        nit = torch.empty((B,), dtype=torch.int32, device=X.device)
        _lib.check(L.ainmf_nmf_fit(_lib.handle(dev), _p(X), B, F, T, rank, max_iter, tol, _solver_id(solver),
                                   seed & 0xFFFFFFFF, _p(W0), _p(H0), _p(W), _p(H), _p(err), _p(nit),
                                   _stream(dev)), dev)
    return W, H, err, nit


def _istft(Z, n_fft, hop, length):
    if Z.dtype != torch.complex64 or Z.dim() != 3:
        raise RuntimeError("Z must be complex64 [B, F, T]")
    Z = Z.contiguous()
    dev = _dev(Z)
    B, F, T = Z.shape
    if F != n_fft // 2 + 1:
        raise RuntimeError(f"Z has {F} bins, n_fft={n_fft} needs {n_fft // 2 + 1}")
    L = _lib.lib()
    with torch.cuda.device(dev):
        y = torch.empty((B, length), dtype=torch.float32, device=Z.device)
        _lib.check(L.ainmf_istft(_lib.handle(dev), _p(Z), B, T, n_fft, hop, length, _p(y), _stream(dev)), dev)
    return y


_workspaces: dict[int, torch.Tensor] = {}


def _workspace(dev: int, nbytes: int) -> torch.Tensor:
    ws = _workspaces.get(dev)
    if ws is None or ws.numel() < nbytes:
        _workspaces.pop(dev, None)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=torch.device("cuda", dev))
        _workspaces[dev] = ws
    return ws


def _nmf_inpaint(x, n_fft, hop, rank, max_iter, tol, seed, threshold, frac_num, frac_den, col_start, col_end,
                 n_outer, W0, H0, solver="cd"):
    x = _f32(x, "x")
    if x.dim() != 2:
        raise RuntimeError("x must be [B, N]")
    dev = _dev(x)
    B, N = x.shape
    L = _lib.lib()
    with torch.cuda.device(dev):
        h = _lib.handle(dev)
        p = _capi.default_params(L, batch=B, n_samples=N, n_fft=n_fft, hop=hop, rank=rank, max_iter=max_iter, tol=tol,
                                 seed=seed & 0xFFFFFFFF, threshold=threshold, frac_num=frac_num, frac_den=frac_den,
                                 col_start=col_start, col_end=col_end, n_outer=n_outer, solver=_solver_id(solver))
        nbytes = L.ainmf_workspace_bytes(h, C.byref(p))
        if nbytes == 0:
            _lib.check(_capi.ERR_INVALID, dev)
        T, F, _ = _capi.stft_geometry(L, N, n_fft, hop)
        if (W0 is None) != (H0 is None):
            raise RuntimeError("W0 and H0 must be given together")
        if W0 is not None:
            W0 = _f32(W0, "W0").expand(B, F, rank).contiguous()
            H0 = _f32(H0, "H0").expand(B, rank, T).contiguous()
        ws = _workspace(dev, nbytes)
        y = torch.empty((B, N), dtype=torch.float32, device=x.device)
        idx = torch.full((B, T), -1, dtype=torch.int32, device=x.device)
        nb = torch.zeros((B,), dtype=torch.int32, device=x.device)
        W = torch.zeros((B, F, rank), dtype=torch.float32, device=x.device)
        H = torch.zeros((B, rank, T), dtype=torch.float32, device=x.device)
        err = torch.zeros((B,), dtype=torch.float32, device=x.device)
        nit = torch.zeros((B,), dtype=torch.int32, device=x.device)
        _lib.check(L.ainmf_inpaint(h, C.byref(p), _p(x), _p(W0), _p(H0), _p(y), _p(idx), _p(nb), _p(W), _p(H),
                                   _p(err), _p(nit), _p(ws), ws.numel(), _stream(dev)), dev)
    return y, idx, nb, W, H, err, nit


def _load_pcm16(pcm):
    if pcm.dtype != torch.int16 or pcm.dim() not in (2, 3):
        raise RuntimeError("pcm must be int16 [B, N] or [B, N, C]")
    pcm = pcm.contiguous()
    dev = _dev(pcm)
    B, N = pcm.shape[:2]
    ch = pcm.shape[2] if pcm.dim() == 3 else 1
    L = _lib.lib()
    with torch.cuda.device(dev):
        x = torch.empty((B, N), dtype=torch.float32, device=pcm.device)
        peak = torch.empty((B,), dtype=torch.float32, device=pcm.device)
        _lib.check(L.ainmf_load_pcm16(_lib.handle(dev), _p(pcm), B, N, ch, _p(x), _p(peak), _stream(dev)), dev)
    return x, peak


def _store_pcm16(y):
    y = _f32(y, "y")
    dev = _dev(y)
    L = _lib.lib()
    with torch.cuda.device(dev):
        out = torch.empty(y.shape, dtype=torch.int16, device=y.device)
        _lib.check(L.ainmf_store_pcm16(_lib.handle(dev), _p(y), y.numel(), _p(out), _stream(dev)), dev)
    return out


def _find_main_gap(x, threshold):
    """main3_AR_text_gap.py:34-49 -> [B, 2] int64 {start, end} ({-1, -1}: no gap)."""
    x = _f32(x, "x")
    if x.dim() != 2:
        raise RuntimeError("x must be [B, N]")
    dev = _dev(x)
    B, N = x.shape
    L = _lib.lib()
    with torch.cuda.device(dev):
        span = torch.empty((B, 2), dtype=torch.int64, device=x.device)
        _lib.check(L.ainmf_find_main_gap(_lib.handle(dev), _p(x), B, N, threshold, _p(span), _stream(dev)), dev)
    return span


def _find_gaps(x, threshold, min_len, max_runs):
    """main3_AR_text_mask.py:30-52 -> (runs [B, max_runs, 2] int64 (-1 padded), n_runs [B] int32)."""
    x = _f32(x, "x")
    if x.dim() != 2:
        raise RuntimeError("x must be [B, N]")
    dev = _dev(x)
    B, N = x.shape
    L = _lib.lib()
    with torch.cuda.device(dev):
        runs = torch.full((B, max_runs, 2), -1, dtype=torch.int64, device=x.device)
        n = torch.empty((B,), dtype=torch.int32, device=x.device)
        _lib.check(L.ainmf_find_gaps(_lib.handle(dev), _p(x), B, N, threshold, min_len, _p(runs), max_runs, _p(n), _stream(dev)), dev)
    return runs, n


def _linear_interp(x, threshold):
    """linear_interp_part1.py:52-75 -> (y [B, N], n_damaged [B] int64)."""
    x = _f32(x, "x")
    if x.dim() != 2:
        raise RuntimeError("x must be [B, N]")
    dev = _dev(x)
    B, N = x.shape
    L = _lib.lib()
    with torch.cuda.device(dev):
        y = torch.empty_like(x)
        nd = torch.empty((B,), dtype=torch.int64, device=x.device)
        _lib.check(L.ainmf_linear_interp(_lib.handle(dev), _p(x), B, N, threshold, _p(y), _p(nd), _stream(dev)), dev)
    return y, nd


def _blend_boundaries(raw, restored, gap_start, gap_end, blend_len):
    """main4_NMF.py:114-126 on 1-D signals."""
    raw, restored = _f32(raw, "raw"), _f32(restored, "restored")
    if raw.dim() != 1 or raw.shape != restored.shape:
        raise RuntimeError("raw and restored must be 1-D and of equal length")
    dev = _dev(raw)
    L = _lib.lib()
    with torch.cuda.device(dev):
        out = torch.empty_like(raw)
        _lib.check(L.ainmf_blend_boundaries(_lib.handle(dev), _p(raw), _p(restored), raw.numel(), gap_start, gap_end, blend_len,
                                            _p(out), _stream(dev)), dev)
    return out


@_serialised
def apply_gaps_(x, starts, lens):
    """In place: x[b, s:s+l] = 0 for the gap lists starts/lens [B, G] (int64, device) -- the zeroing step of the fixture
    producers (generate_part1_data.py:44-46, generate_part2_data.py:36-43)."""
    x = _f32(x, "x")
    if x.dim() != 2 or starts.shape != lens.shape or starts.dim() != 2 or starts.shape[0] != x.shape[0]:
        raise RuntimeError("x must be [B, N] and starts/lens [B, G]")
    if starts.dtype != torch.int64 or lens.dtype != torch.int64:
        raise RuntimeError("starts and lens must be int64")
    dev = _dev(x)
    starts, lens = starts.contiguous(), lens.contiguous()
    L = _lib.lib()
    with torch.cuda.device(dev):
        _lib.check(L.ainmf_apply_gaps(_lib.handle(dev), _p(x), x.shape[0], x.shape[1], _p(starts), _p(lens), starts.shape[1],
                                      _stream(dev)), dev)
    return x


def standard_normal(seed, n, device=None):
    """float32(numpy.random.RandomState(seed).standard_normal(n)) generated on the device (ainmf_standard_normal): the stream
    sklearn's init='random' draws H0 then W0 from."""
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    L = _lib.lib()
    with torch.cuda.device(dev):
        out = torch.empty((int(n),), dtype=torch.float32, device=torch.device("cuda", dev))
        _lib.check(L.ainmf_standard_normal(_lib.handle(dev), int(seed) & 0xFFFFFFFF, int(n), _p(out), _stream(dev)), dev)
    return out


def set_window(n_fft, window=None, device=None):
    """The `window` argument of scipy.signal.stft / istft for every later call with this n_fft on `device` (default: the
    current CUDA device): an array-like of n_fft values, or None for scipy's default periodic Hann window (what the
    reference runs, main4_NMF_gap.py:47,71).  ainmf_set_window."""
    import numpy as np
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    L = _lib.lib()
    ptr = None
    if window is not None:
        w = np.ascontiguousarray(np.asarray(window, dtype=np.float32))
        if w.shape != (n_fft,):
            raise RuntimeError(f"window must hold n_fft = {n_fft} values, got shape {w.shape}")
        ptr = w.ctypes.data_as(C.c_void_p)
    with torch.cuda.device(dev):
        _lib.check(L.ainmf_set_window(_lib.handle(dev), int(n_fft), ptr), dev)


@_serialised
def snr_db(ref, est, begin=0, end=None):
    """10 log10(sum ref^2 / (sum (ref - est)^2 + 1e-10)) over [begin, end) (main4_NMF.py:99-110); returns a Python float."""
    ref, est = _f32(ref, "ref"), _f32(est, "est")
    if ref.dim() != 1 or ref.shape != est.shape:
        raise RuntimeError("ref and est must be 1-D and of equal length")
    dev = _dev(ref)
    end = ref.numel() if end is None else end
    out = C.c_double()
    L = _lib.lib()
    with torch.cuda.device(dev):
        _lib.check(L.ainmf_snr_db(_lib.handle(dev), _p(ref), _p(est), begin, end, C.byref(out), _stream(dev)), dev)
    return out.value


for _name, _fn in (("find_main_gap", _find_main_gap), ("find_gaps", _find_gaps), ("linear_interp", _linear_interp),
                   ("blend_boundaries", _blend_boundaries)):
    _LIBRARY.impl(_name, _serialised(_fn), "CUDA")

for _name, _fn in (("stft", _stft), ("gap_mask", _gap_mask), ("nmf_fit", _nmf_fit), ("istft", _istft),
                   ("nmf_inpaint", _nmf_inpaint), ("load_pcm16", _load_pcm16), ("store_pcm16", _store_pcm16)):
    _LIBRARY.impl(_name, _serialised(_fn), "CUDA")

stft = torch.ops.ainmf.stft
gap_mask = torch.ops.ainmf.gap_mask
nmf_fit = torch.ops.ainmf.nmf_fit
istft = torch.ops.ainmf.istft
nmf_inpaint = torch.ops.ainmf.nmf_inpaint
load_pcm16 = torch.ops.ainmf.load_pcm16
store_pcm16 = torch.ops.ainmf.store_pcm16
find_main_gap = torch.ops.ainmf.find_main_gap
find_gaps = torch.ops.ainmf.find_gaps
linear_interp = torch.ops.ainmf.linear_interp
blend_boundaries = torch.ops.ainmf.blend_boundaries
