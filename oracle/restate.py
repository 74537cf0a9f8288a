"""First-principles CPU restatement of the hot path -- TEST INFRASTRUCTURE ONLY.

What scipy.signal / sklearn / numpy compute on the reference's path, written out stage by stage
(numpy for array work, ``oracle/c/oracle_c.c`` for the two order-sensitive native loops), each
stage citing what it follows.  ``$SP`` = site-packages of scipy 1.18.1 / scikit-learn 1.9.0 /
numpy 2.3.5 (the versions the reference runs on in this image; it pins none, README.md:72).

The layouts here are the product's: spectrogram arrays are frame-major ``[T, F]``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _clib():
    """Load (building if needed) oracle/_build/liboracle_c.so."""
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "liboracle_c.so")
        src = os.path.join(_HERE, "c", "oracle_c.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        lib = ctypes.CDLL(so)
        lib.oracle_cd_sweep_f32.restype = ctypes.c_float
        lib.oracle_cd_sweep_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_ssize_t, ctypes.c_ssize_t]
        lib.oracle_standard_normal.restype = None
        lib.oracle_standard_normal.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t]
        _LIB = lib
    return _LIB


# ---- window / framing ---------------------------------------------------------------------------

def hann_periodic(n: int) -> np.ndarray:
    """get_window('hann_periodic', n) -> general_cosine(n, [0.5, 0.5], sym=False)
    ($SP/scipy/signal/windows/_windows.py:56-66): fac = linspace(-pi, pi, n+1)[:n];
    w = 0.5 + 0.5 cos(fac), float64."""
    fac = np.linspace(-np.pi, np.pi, n + 1)[:n]
    return 0.5 * np.cos(0.0 * fac) + 0.5 * np.cos(1.0 * fac)


def stft_geometry(N: int, n_fft: int, hop: int):
    """Frame count and tail pad of scipy.signal.stft(boundary='zeros', padded=True)
    ($SP/scipy/signal/_spectral_py.py:2240,2247): returns (T, nadd)."""
    ext = N + 2 * (n_fft // 2)
    nadd = (-(ext - n_fft) % hop) % n_fft
    T = (ext + nadd - n_fft) // hop + 1
    return T, nadd


def stft(x: np.ndarray, n_fft: int, hop: int, window: np.ndarray | None = None,
         dtype=np.float32) -> np.ndarray:
    """Z[T, F] complex64 = scipy.signal.stft(x, nperseg=n_fft, noverlap=n_fft-hop)[2].T
    ($SP/scipy/signal/_spectral_py.py:2215-2326, 2376-2395).  `dtype` is the arithmetic type of
    the FFT (scipy ends up in float64 because np.zeros pads promote x; the result is cast to
    complex64 at :2329)."""
    N = len(x)
    w = hann_periodic(n_fft) if window is None else np.asarray(window, dtype=np.float64)
    w32 = w.astype(np.float32)                      # win.astype(complex64), :2272
    T, nadd = stft_geometry(N, n_fft, hop)
    ext = np.zeros(N + n_fft + nadd, dtype=dtype)
    ext[n_fft // 2: n_fft // 2 + N] = x
    idx = np.arange(T)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = ext[idx] * w32.astype(dtype)[None, :]
    Z = np.fft.rfft(frames.astype(np.float64 if dtype == np.float64 else np.float32), axis=1)
    scale = 1.0 / float(w32.sum(dtype=np.float32))  # sqrt(1/sum(w)^2), :2277-2282
    return (Z * scale).astype(np.complex64)


def istft(Z: np.ndarray, n_fft: int, hop: int, length: int | None = None,
          window: np.ndarray | None = None) -> np.ndarray:
    """scipy.signal.istft(Z.T, nperseg=n_fft, noverlap=n_fft-hop)[1] in float32
    ($SP/scipy/signal/_spectral_py.py:1872-1910); Z is [T, F]."""
    T = Z.shape[0]
    w = (hann_periodic(n_fft) if window is None else np.asarray(window, np.float64)).astype(np.float32)
    xs = np.fft.irfft(Z.astype(np.complex64), n=n_fft, axis=1).astype(np.float32)
    xs *= w.sum(dtype=np.float32)
    out_len = n_fft + (T - 1) * hop
    y = np.zeros(out_len, np.float32)
    norm = np.zeros(out_len, np.float32)
    w2 = w * w
    for c in range(T):
        y[c * hop: c * hop + n_fft] += xs[c] * w
        norm[c * hop: c * hop + n_fft] += w2
    y = y[n_fft // 2: out_len - n_fft // 2]
    norm = norm[n_fft // 2: out_len - n_fft // 2]
    y = y / np.where(norm > 1e-10, norm, np.float32(1.0))
    return y if length is None else y[:length]


# ---- column mask (bit-exact contract) -------------------------------------------------------------

def frac_to_ratio(frac: float):
    """cnt/len > frac  <=>  den*cnt > num*len for the two fractions the reference uses
    (0.9 -> 9/10, main4_NMF_gap.py:38; 0.8 -> 4/5, main4_NMF_mask.py:42).  Verified exhaustively
    for len <= 4096 in tests/test_oracle.py."""
    from fractions import Fraction
    fr = Fraction(frac).limit_denominator(1000)
    return fr.numerator, fr.denominator


def column_mask(x: np.ndarray, n_frames: int, hop: int, threshold: float, num: int, den: int):
    """Integer form of get_gap_mask / get_mask_from_signal (main4_NMF_gap.py:28-40,
    main4_NMF_mask.py:28-45): g[i] = |x[i]| < float32(threshold); column c is bad iff
    len > 0 and den*cnt > num*len over [max(0, c*hop - hop//2), min(N, c*hop + hop//2))."""
    thr32 = np.float32(threshold)
    g = (np.abs(x.astype(np.float32)) < thr32).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(g)])
    N = len(x)
    c = np.arange(n_frames, dtype=np.int64) * hop
    ws = np.maximum(0, c - hop // 2)
    we = np.minimum(N, c + hop // 2)
    ln = we - ws
    ok = ln > 0
    cnt = np.where(ok, cs[np.clip(we, 0, N)] - cs[np.clip(ws, 0, N)], 0)
    bad = ok & (den * cnt > num * ln)
    return np.nonzero(bad)[0].astype(np.int64)


# ---- imputation / initial factors -----------------------------------------------------------------

def impute(V: np.ndarray, bad_cols: np.ndarray):
    """V is [T, F].  fill[f] = mean over good frames (main4_NMF_gap.py:56-58); X = V with bad
    frames <- fill (:59).  Returns (X, fill)."""
    good = np.ones(V.shape[0], bool)
    good[bad_cols] = False
    fill = np.mean(V[good].T, axis=1).astype(np.float32)   # same pairwise order as mean(axis=1) on (F,Tg)
    X = V.copy()
    X[bad_cols] = fill[None, :]
    return X, fill


def standard_normal(seed: int, n: int) -> np.ndarray:
    """RandomState(seed).standard_normal(n) via the C restatement (float64)."""
    out = np.empty(n, np.float64)
    _clib().oracle_standard_normal(ctypes.c_uint32(seed & 0xFFFFFFFF), out.ctypes.data, n)
    return out


def init_factors(mean_X: float, F: int, T: int, K: int, seed: int):
    """_initialize_nmf(init='random') ($SP/sklearn/decomposition/_nmf.py:296-307):
    avg = sqrt(mean(X)/K) in float32; H (K,T) drawn FIRST then W (F,K); float64 normals cast to
    float32, scaled, abs.  Returns (W0 [F,K], Ht0 [T,K]) -- Ht0 is H0 transposed."""
    avg = np.sqrt(np.float32(mean_X) / np.float32(K)).astype(np.float32)
    z = standard_normal(seed, K * T + F * K)
    H0 = np.abs(avg * z[:K * T].astype(np.float32).reshape(K, T))
    W0 = np.abs(avg * z[K * T:].astype(np.float32).reshape(F, K))
    return np.ascontiguousarray(W0), np.ascontiguousarray(H0.T)


# ---- coordinate-descent NMF -----------------------------------------------------------------------

def cd_sweep(A: np.ndarray, G: np.ndarray, B: np.ndarray) -> float:
    """In-place _update_cdnmf_fast (float32) on C-contiguous A [rows,K]."""
    assert A.dtype == np.float32 and A.flags.c_contiguous
    G = np.ascontiguousarray(G, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    return float(_clib().oracle_cd_sweep_f32(A.ctypes.data, G.ctypes.data, B.ctypes.data,
                                             A.shape[0], A.shape[1]))


def nmf_cd(Xt: np.ndarray, W0: np.ndarray, Ht0: np.ndarray, max_iter=200, tol=1e-4,
           trace: list | None = None):
    """_fit_coordinate_descent ($SP/sklearn/decomposition/_nmf.py:399-518) on X = Xt.T.
    Xt [T,F], W0 [F,K], Ht0 [T,K]; returns (W, Ht, n_iter, err) with
    err = ||X - W H||_F (_beta_divergence, :118-127 then sqrt(2*.), :1623)."""
    X = np.ascontiguousarray(Xt.T)            # (F,T) C-order like the reference's current_mag
    W = np.array(W0, np.float32, order="C", copy=True)
    Ht = np.array(Ht0, np.float32, order="C", copy=True)
    v1 = 0.0
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        v = cd_sweep(W, np.dot(Ht.T, Ht), np.dot(X, Ht))
        v += cd_sweep(Ht, np.dot(W.T, W), np.dot(X.T, W))
        if trace is not None:
            trace.append(v)
        if n_iter == 1:
            v1 = v
        if v1 == 0:
            break
        if v / v1 <= tol:
            break
    R = X - W @ Ht.T
    err = float(np.sqrt(np.dot(R.ravel(), R.ravel())))
    return W, Ht, n_iter, err


def nmf_mu_fro(Xt, W0, Ht0, n_iter):
    """Fixed-count Frobenius multiplicative update ($SP/sklearn/decomposition/_nmf.py:536-549,
    615-624, 633-635, 701-721) without the every-10-iterations stop test."""
    EPS = np.finfo(np.float32).eps
    X = np.ascontiguousarray(Xt.T)
    W = np.array(W0, np.float32, copy=True)
    H = np.array(Ht0.T, np.float32, order="C", copy=True)
    for _ in range(n_iter):
        num = X @ H.T
        den = W @ (H @ H.T)
        den[den == 0] = EPS
        W *= num / den
        num = W.T @ X
        den = (W.T @ W) @ H
        den[den == 0] = EPS
        H *= num / den
    R = X - W @ H
    return W, np.ascontiguousarray(H.T), float(np.sqrt(np.sum(R.astype(np.float64) ** 2)))


# ---- recombination / full pipeline ----------------------------------------------------------------

def recombine(Z: np.ndarray, V: np.ndarray, W: np.ndarray, Ht: np.ndarray, bad_cols: np.ndarray):
    """final_mag = V with bad frames <- (W H)[:, bad]; Z' = final_mag * exp(1j*angle(Z))
    (main4_NMF_gap.py:65-70).  All [T, F]."""
    M = V.copy()
    M[bad_cols] = (Ht[bad_cols] @ W.T).astype(np.float32)
    phase = np.angle(Z)
    return (M * np.exp(1j * phase)).astype(np.complex64)


def restore_columns(x, *, n_fft=1024, hop=256, threshold=1e-4, frac=0.9, K=40, seed=42,
                    max_iter=200, tol=1e-4, W0=None, Ht0=None):
    """Stage-wise NMFFairGapInpainter.restore (main4_NMF_gap.py:42-72); returns (y, stages)."""
    x = np.asarray(x, np.float32)
    Z = stft(x, n_fft, hop)
    V = np.abs(Z)
    T, F = V.shape
    num, den = frac_to_ratio(frac)
    bad = column_mask(x, T, hop, threshold, num, den)
    st = dict(Z=Z, V=V, bad=bad)
    if len(bad) == 0:
        return x, st
    X, fill = impute(V, bad)
    if W0 is None:
        W0, Ht0 = init_factors(X.T.mean(), F, T, K, seed)   # mean over the (F,T) C-order array
    W, Ht, n_iter, err = nmf_cd(X, W0, Ht0, max_iter, tol)
    Zr = recombine(Z, V, W, Ht, bad)
    y = istft(Zr, n_fft, hop, len(x))
    st.update(X=X, fill=fill, W0=W0, Ht0=Ht0, W=W, Ht=Ht, n_iter=n_iter, err=err, Zr=Zr)
    return y, st
